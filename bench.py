#!/usr/bin/env python
"""bench.py — BASELINE.json metric on the B200 engine (and, with --impl reference, on the CPU restatement).

Workload (BASELINE.json configs[1]): Qwen3-TTS-12Hz-0.6B, MLX 4-bit g64, `generateStream` semantics (stream sampler variant,
temperature 0.85, chunk 12, codec windows 18 / 8+18), `--batch` independent utterances per GPU (default 64) with 8-40 text
ids from seed 1 and different speakers, `--frames` frames each (default 36 = two decode windows).  Synthetic seeded weights
(no network).  One *step* = one such batch: prompt assembly + prefill + 36 frame steps + windowed codec decode to PCM.

  value  = audio seconds produced by all GPUs / device time of the step (CUDA events on the engine's stream: talker span +
           codec passes; ids are tiny so "inputs resident" only excludes the PCM read-back);
  e2e    = the same through the public C-ABI call (`q3tts_generate_pcm_batch`) with host buffers, wall clock, including the
           H2D of ids/codes and the D2H of every PCM sample;
  roofline = the dequant-fused linear kernel (dominant: >97 % of bytes), measured live with CUDA events over the launches of
           one talker decode step (q3tts_profile_linear), against MEASURED_PEAKS.json hbm_gbs.

Multi-GPU: request-parallel replicas — one process per GPU (torchrun), each with its own shard of
utterances (weak scaling); NCCL is used only for the barrier, the max-over-ranks time and the gather of per-rank counts.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))

import numpy as np  # noqa: E402

METRIC = "audio-sec generated per sec (RTFx)"
UNIT = "audio-sec/s"
SPEAKERS = [3066, 3065, 3010, 3061, 2861, 2873, 2864, 2875, 2878]


def make_requests(q, n, frames, seed, rank=0):
    rng = np.random.default_rng(seed * 1000 + rank)
    reqs = []
    for i in range(n):
        n_ids = int(rng.integers(8, 41)) + 9  # 8-40 text ids + the 9 template ids
        ids = rng.integers(0, 150000, size=n_ids).tolist()
        reqs.append(q.GenRequest(text_ids=ids, speaker_id=SPEAKERS[i % len(SPEAKERS)], temperature=0.85, max_tokens=frames,
                                 seed=seed * 100000 + rank * 1000 + i, stream_variant=True))
    return reqs


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.p, self.first = index, [], None, 0

    def mark(self):
        """samples before this point (warm-up) are not reported"""
        self.first = len(self.rows)

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.p = None

    def _read(self):
        for line in self.p.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.p:
            self.p.terminate()
        rows = self.rows[self.first:] or self.rows[-1:]
        sm = [float(r[0]) for r in rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return j.get("hbm_gbs", 6650.0), j.get("bf16_tflops_sustained", 1400.0), "measured"
    return 6650.0, 1590.0, "fallback"


# --------------------------------------------------------------------------------------------------- CPU restatement arm
def cpu_sample(ckpt_dir, frames, reps, warmup):
    """Times the CPU restatement of the reference graph (oracle/, torch-CPU fp32, all host threads) on a bounded sample of
    the same workload: one utterance, `frames` frames of the stream-variant loop + one codec decode of those frames.
    Warm-up runs are 3 frames long (they only page the weights in), timed runs `frames` long."""
    import torch

    from oracle import codec as ocodec, pipeline as opipe, talker as otalker

    torch.set_num_threads(os.cpu_count() or 1)
    orc = otalker.TalkerOracle(ckpt_dir)
    cdc = ocodec.load_codec(ckpt_dir)
    rng = np.random.default_rng(1)
    times = []
    for it in range(warmup + reps):
        ids = rng.integers(0, 150000, size=20).tolist()
        t0 = time.perf_counter()
        fr = orc.generate_codes(otalker.Request(text_ids=ids, speaker_id=2861, temperature=0.85, max_tokens=frames if it >= warmup else 3, seed=it,
                                                stream_variant=True), filter_invalid=False)
        valid = [f for f in fr if 0 <= f[0] < 2048] or [[0] * 16]
        opipe.decode_whole(cdc, valid)
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append((len(fr) * 0.08, dt))
    audio = sum(a for a, _ in times)
    secs = sum(t for _, t in times)
    return audio / secs, secs / max(1, len(times))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=36)
    ap.add_argument("--bits", type=int, default=4)
    ap.add_argument("--model", default="0.6b")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-frames", type=int, default=60)
    a = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    assert world == max(1, a.gpus) or world == 1, f"WORLD_SIZE {world} != --gpus {a.gpus}"
    wfmt = f"{a.bits}-bit g64" if a.bits else "bf16"
    config = {"workload": f"Qwen3-TTS-12Hz-{a.model} {wfmt} generateStream: {a.batch} utterances/GPU x {a.frames} frames, stream windows 18/8+18",
              "model": a.model, "bits": a.bits, "batch_per_gpu": a.batch, "frames": a.frames, "parallelism": f"request-parallel x{world}",
              "cache": "batched steps stream 3.3 GB of fp16 weight copies per frame-step (0.6B) >> L2 126 MB; no explicit flush"}

    from oracle import checkpoint

    ckpt_dir = f"/tmp/q3tts_bench_{a.model}_{a.bits}"

    if a.impl == "reference":
        # The reference's own implementation of the path cannot run here (Swift + MLX): this arm times the CPU
        # restatement (oracle/) on the host cores, rank 0 only.
        if rank != 0:
            return
        checkpoint.write_checkpoint(ckpt_dir, a.model, bits=a.bits, dtype="bf16", seed=0)
        v, sec = cpu_sample(ckpt_dir, a.cpu_frames, max(1, a.steps), min(a.warmup, 1))
        cores = os.cpu_count() or 1
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1),
                "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": config,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"1 utterance x {a.cpu_frames} frames (stream loop, batch 1 = the reference's only mode) + 1 whole-sequence codec decode per step; CPU restatement of the reference graph (torch fp32, all host threads), not MLX"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return

    # stdout carries exactly one JSON line: NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION in this image) would precede it
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"
    import torch
    import torch.distributed as dist

    import qwen3tts_b200 as q

    if world > 1:
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    if rank == 0:
        checkpoint.write_checkpoint(ckpt_dir, a.model, bits=a.bits, dtype="bf16", seed=0)
    if world > 1:
        dist.barrier()

    # torch's own CUDA state (lazy init on first use) is brought up BEFORE the warm-up steps: initialising it at the
    # synchronize() that opens the timed region stalled kernel submission inside the first timed step on some boxes.
    torch.cuda.set_device(local_rank)
    torch.zeros(8, device="cuda").sum().item()
    torch.cuda.synchronize()
    eng = q.Engine(ckpt_dir, device=local_rank, max_batch=a.batch, max_frames=max(64, a.frames), kv_capacity=512)
    up = eng.info.codec_total_upsample
    out_bufs = [np.zeros(a.frames * up, dtype=np.float32) for _ in range(a.batch)]

    def step(i):
        reqs = make_requests(q, a.batch, a.frames, i, rank)
        t0 = time.perf_counter()
        pcm, frames = eng.generate_pcm_batch(reqs, q.DECODE_STREAM, out_buffers=out_bufs)
        wall = time.perf_counter() - t0
        tm = eng.timing()
        samples = int(sum(p.size for p in pcm))
        return {"wall": wall, "dev": (tm.talker_ms + tm.decode_ms) * 1e-3, "talker": tm.talker_ms * 1e-3, "decode": tm.decode_ms * 1e-3,
                "samples": samples, "frames": int(tm.frames), "launches": int(tm.kernel_launches), "h2d": int(tm.h2d_bytes), "d2h": int(tm.d2h_bytes),
                "codec_flops": int(tm.codec_flops), "bytes_frame": int(tm.weight_bytes_per_frame)}

    # nvidia-smi is started BEFORE the warm-up steps: its start-up (NVML init takes driver locks for tens to hundreds of ms on a
    # box without persistence mode) would otherwise stall kernel submission inside the first timed step; it keeps sampling
    # through the timed region, and only the samples taken after the warm-up are reported.
    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(a.warmup):
        step(1000 + i)
    sampler.mark()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t_begin = time.perf_counter()
    res = [step(i) for i in range(a.steps)]
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t_total = time.perf_counter() - t_begin
    clocks = sampler.stop()

    dev = sum(r["dev"] for r in res)
    wall = sum(r["wall"] for r in res)
    samples = sum(r["samples"] for r in res)
    stats = torch.tensor([dev, wall, t_total], dtype=torch.float64, device="cuda")
    sums = torch.tensor([samples, sum(r["launches"] for r in res), sum(r["h2d"] for r in res), sum(r["d2h"] for r in res)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)  # time = max over ranks
        dist.all_reduce(sums, op=dist.ReduceOp.SUM)   # work = sum over ranks
    dev_max, wall_max, total_max = [float(x) for x in stats.tolist()]
    samples_all, launches_all, h2d_all, d2h_all = [float(x) for x in sums.tolist()]
    audio_s = samples_all / 24000.0

    # roofline of the dominant kernel (dequant-fused linear), live, on rank 0
    hbm, tf, src = peaks()
    roof = lat = codec4 = None
    cpu_base = None
    if rank == 0:
        iters = 20
        ms, n, nbytes = eng.profile_linear(0, a.batch, iters)
        ach = nbytes * iters / (ms * 1e-3) / 1e9
        kname = ("tc_skinny_kernel (tcgen05 / TMEM split-K cluster GEMM over fp16 dense weight copies), the 113 linears of one talker decode step at m = batch rows, replayed as a CUDA graph"
                 if 16 <= a.batch <= 128 else "tc_gemm_kernel" if a.batch > 128 else "linear_kernel (dequant-fused GEMV), the linears of one talker decode step at m = batch rows")
        traffic = None
        tp = os.path.join(ROOT, "profiles", "traffic.json")  # dram__bytes_read+write per launch from the committed ncu --set full capture
        if os.path.exists(tp):
            tj = json.load(open(tp))
            traffic = tj.get("tc_skinny_m64" if a.batch >= 16 else "linear_m1", {}).get("dram_bytes_per_launch")
        roof = {"bound": "hbm", "kernel": kname,
                "achieved": ach, "peak": hbm, "unit": "GB/s", "frac": ach / hbm, "traffic": traffic, "peak_source": src,
                "launches_timed": n, "algorithmic_bytes_per_launch": nbytes / (n / iters), "avg_launch_us": ms * 1e3 / n}
        ms1, n1, nb1 = eng.profile_linear(0, 1, iters)
        msc, nc, nbc = eng.profile_linear(1, 1, iters)
        # batch-1 latency view of the same path (the reference's only mode)
        e1 = q.Engine(ckpt_dir, device=local_rank, max_batch=1, max_frames=64, load_codec=False)
        r1 = q.GenRequest(text_ids=list(range(1000, 1024)), speaker_id=2861, temperature=0.85, max_tokens=36, stream_variant=True, keep_invalid_frames=True)
        e1.generate_codes(r1)
        e1.generate_codes(r1)
        t1 = e1.timing()
        msf = (t1.talker_ms - t1.prefill_ms) / max(1, t1.frames)
        lat = {"batch1_path": "persistent frame kernel (1 cooperative launch per 8 frames)" if t1.persistent_launches else "CUDA graph of per-op kernels",
               "batch1_ms_per_frame": msf, "batch1_rtfx": 80.0 / msf, "batch1_prefill_ms": t1.prefill_ms,
               "batch1_frame_roofline_frac": (t1.weight_bytes_per_frame / (msf * 1e-3) / 1e9) / hbm,
               "batch1_linear_gbs_talker_step": nb1 * iters / (ms1 * 1e-3) / 1e9, "batch1_linear_gbs_cp_pass_L2": nbc * iters / (msc * 1e-3) / 1e9,
               "time_to_first_chunk_ms": t1.prefill_ms + 18 * msf}
        e1.close()
        # BASELINE.json configs[3] (codec decode only, 16-codebook 12.5 Hz codes -> 24 kHz, 60 s clips, chunkedDecode(100, 10)) on a
        # bounded sample of its 128 clips: 16 clips x 750 frames -> 128 chunks of 110 frames, device time of the passes
        clips = np.random.default_rng(3).integers(0, 2048, size=(16, 750, 16)).astype(np.int32)
        eng.decode_chunked(clips)  # warm-up at the measured shape (workspace growth, first-use kernels)
        eng.decode_chunked(clips)
        tc = eng.timing()
        codec4 = {"workload": "configs[3] sample: 16 of 128 clips x 750 frames, chunkedDecode(100, 10) = 128 chunks x 110 frames",
                  "samples_per_s": 16 * 750 * up / max(1e-9, tc.decode_ms * 1e-3), "decode_ms": tc.decode_ms,
                  "tflops": tc.codec_flops / max(1e-9, tc.decode_ms * 1e-3) / 1e12, "peak_tflops": tf}
        if world == 1 and not a.no_cpu_baseline:
            v, sec = cpu_sample(ckpt_dir, a.cpu_frames, 1, 1)
            cpu_base = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"1 utterance x {a.cpu_frames} frames (stream loop, batch 1 = the reference's only mode) + 1 whole-sequence codec decode, 1 timed run of {sec:.1f} s after a 3-frame warm-up; CPU restatement of the reference graph (torch fp32, all host threads), not MLX"}

    if rank == 0:
        codec_sps = samples / max(1e-9, sum(r["decode"] for r in res))
        line = {"metric": METRIC, "value": audio_s / dev_max, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": dev_max / a.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16 operands (dequantised once at load) x f32 accumulate, f32 residual stream" if a.batch >= 16 else "f32 activations, u4 g64 weights",
                "data": "synthetic", "config": config,
                "e2e": {"value": audio_s / wall_max, "unit": UNIT, "h2d_bytes_per_step": h2d_all / world / a.steps, "d2h_bytes_per_step": d2h_all / world / a.steps,
                        "ms_per_step": wall_max / a.steps * 1e3},
                "gpu_launches": int(launches_all), "clocks": clocks, "roofline": roof, "cpu_baseline": cpu_base, "latency": lat,
                "codec": {"samples_per_s_rank0": codec_sps, "tflops_rank0": sum(r["codec_flops"] for r in res) / max(1e-9, sum(r["decode"] for r in res)) / 1e12,
                          "peak_tflops": tf, "share_of_step": sum(r["decode"] for r in res) / max(1e-9, dev)},
                "codec_decode_only": codec4,
                "talker": {"ms_per_frame_batch": sum(r["talker"] for r in res) / a.steps / a.frames * 1e3, "share_of_step": sum(r["talker"] for r in res) / max(1e-9, dev)},
                "wall_s_timed_region": total_max}
        print(json.dumps(line), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
