"""On-box: codec decode throughput (BASELINE config 4 shapes, reduced) and a launch list under ncu.  Not the bench."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
T = int(sys.argv[2]) if len(sys.argv) > 2 else 26
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
d = checkpoint.write_checkpoint("/tmp/q3tts_codecfull", "codecfull", bits=8, dtype="bf16", seed=0)
eng = q.Engine(d, load_talker=False, codec_max_frames=2400)
codes = np.random.default_rng(3).integers(0, 2048, size=(B, T, 16)).astype(np.int32)
for rep in range(reps):
    t0 = time.time(); pcm = eng.decode(codes); wall = time.time() - t0
    tm = eng.timing()
    print(json.dumps({"B": B, "T": T, "wall_ms": wall * 1e3, "decode_ms": tm.decode_ms, "samples_per_s": B * T * 1920 / (tm.decode_ms * 1e-3),
                      "tflops": tm.codec_flops / (tm.decode_ms * 1e-3) / 1e12, "launches": tm.kernel_launches}), flush=True)
eng.close()
