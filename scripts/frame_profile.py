"""Small driver for ncu: one batched generation (eager launches, no CUDA graph) so every kernel of a frame step is visible.
usage: frame_profile.py <batch> <frames> [bits]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q

batch, frames = int(sys.argv[1]), int(sys.argv[2])
bits = int(sys.argv[3]) if len(sys.argv) > 3 else 4
d = checkpoint.write_checkpoint(f"/tmp/q3tts_bench_0.6b_{bits}", "0.6b", bits=bits, dtype="bf16", seed=0)
eng = q.Engine(d, max_batch=batch, max_frames=64, use_cuda_graph=False, load_codec=False)
rng = np.random.default_rng(0)
reqs = [q.GenRequest(text_ids=rng.integers(0, 150000, size=int(rng.integers(17, 50))).tolist(), speaker_id=2861, temperature=0.85,
                     max_tokens=frames, seed=i, stream_variant=True, keep_invalid_frames=True) for i in range(batch)]
t0 = time.time()
out = eng.generate_codes_batch(reqs) if batch > 1 else [eng.generate_codes(reqs[0])]
tm = eng.timing()
print(f"batch {batch} frames {frames}: wall {time.time()-t0:.3f}s talker_ms {tm.talker_ms:.2f} prefill_ms {tm.prefill_ms:.2f} launches {tm.kernel_launches}")
