"""On-box: where does the wall time of one bench step (64 utterances x 36 frames, stream windows) go?  Not the bench."""
import os, sys, time, json, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q
import bench

d = checkpoint.write_checkpoint("/tmp/q3tts_bench_0.6b_4", "0.6b", bits=4, dtype="bf16", seed=0)
B, F = 64, 36
eng = q.Engine(d, max_batch=B, max_frames=64, kv_capacity=512)
up = eng.info.codec_total_upsample
outs = [np.zeros(F * up, dtype=np.float32) for _ in range(B)]
for i in range(4):
    reqs = bench.make_requests(q, B, F, i)
    t0 = time.perf_counter()
    pcm, fr = eng.generate_pcm_batch(reqs, q.DECODE_STREAM, out_buffers=outs)
    wall = time.perf_counter() - t0
    tm = eng.timing()
    print(json.dumps({"wall_ms": wall * 1e3, "device_ms": tm.device_ms, "talker_ms": tm.talker_ms, "prefill_ms": tm.prefill_ms, "decode_ms": tm.decode_ms,
                      "launches": tm.kernel_launches, "replays": tm.graph_replays}), flush=True)
reqs = bench.make_requests(q, B, F, 9)
pr = cProfile.Profile(); pr.enable()
eng.generate_pcm_batch(reqs, q.DECODE_STREAM, out_buffers=outs)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(8)
for i in range(2):
    t0 = time.perf_counter(); fr = eng.generate_codes_batch(reqs); wall = time.perf_counter() - t0
    tm = eng.timing()
    print(json.dumps({"codes_only_wall_ms": wall * 1e3, "device_ms": tm.device_ms, "talker_ms": tm.talker_ms, "prefill_ms": tm.prefill_ms}), flush=True)
eng.close()
