"""ncu driver: the linear launches of one talker decode step at m rows (q3tts_profile_linear).  usage: prof_linear.py <m> [bits]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
from oracle import checkpoint
import qwen3tts_b200 as q
m = int(sys.argv[1]); bits = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d = checkpoint.write_checkpoint(f"/tmp/q3tts_bench_0.6b_{bits}", "0.6b", bits=bits, dtype="bf16", seed=0)
eng = q.Engine(d, max_batch=max(m, 1), max_frames=64, load_codec=False)
ms, n, b = eng.profile_linear(0, m, 2)
print(f"m={m}: {ms/2*1e3/(n/2):.2f} us/launch, {b*2/(ms*1e-3)/1e9:.0f} GB/s over {n} launches")
