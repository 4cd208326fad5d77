"""On-box: average launch time of the talker-step / CP-pass linears at m rows (q3tts_profile_linear).  Not the bench."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
from oracle import checkpoint
import qwen3tts_b200 as q
d = checkpoint.write_checkpoint("/tmp/q3tts_bench_0.6b_4", "0.6b", bits=4, dtype="bf16", seed=0)
eng = q.Engine(d, max_batch=64, max_frames=64, load_codec=False)
for which in (0, 1):
    for m in (16, 64, 128):
        ms, n, nb = eng.profile_linear(which, m, 20)
        print(json.dumps({"which": which, "m": m, "us_per_launch": ms * 1e3 / n, "GBs": nb * 20 / (ms * 1e-3) / 1e9}))
eng.close()
