"""Probe: the headline workload (64 utterances x 36 frames, stream windows, host buffers) through ONE call on handles of
max_batch x lanes = 64 x 1, 32 x 2, 16 x 4 -- does splitting one batch over lanes (talker of one lane under the codec of another) pay?"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
import qwen3tts_b200 as q
from bench import make_requests, ckpt_path, INIT
from oracle import checkpoint

d = ckpt_path("0.6b", 4)
checkpoint.write_checkpoint(d, "0.6b", bits=4, dtype="bf16", seed=0, init=INIT)
frames, total = 36, 64
for mb, lanes in ((64, 1), (32, 2), (16, 4)):
    eng = q.Engine(d, max_batch=mb, max_frames=64, lanes=lanes)
    bufs = [np.zeros(frames * 1920, dtype=np.float32) for _ in range(total)]
    def run(seed):
        reqs = make_requests(q, total, frames, seed)
        t0 = time.perf_counter()
        pcm, _ = eng.generate_pcm_batch(reqs, q.DECODE_STREAM, out_buffers=bufs)
        return time.perf_counter() - t0, sum(p.size for p in pcm)
    for i in range(3): run(1000 + i)
    res = [run(i) for i in range(5)]
    best = min(r[0] for r in res)
    print(f"max_batch {mb} x lanes {lanes}: {best*1e3:.1f} ms per step -> {res[0][1]/24000/best:.0f} audio-s/s e2e", flush=True)
    eng.close()
