"""On-box: ms per batched frame-step (CUDA-graph path) for a few batch sizes.  Not the bench.
usage: batch_probe.py [frames] [batch ...]"""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 24
batches = [int(a) for a in sys.argv[2:]] or [64]
d = checkpoint.write_checkpoint("/tmp/q3tts_bench_0.6b_4", "0.6b", bits=4, dtype="bf16", seed=0)
for batch in batches:
    eng = q.Engine(d, max_batch=batch, max_frames=64, load_codec=False)
    rng = np.random.default_rng(0)
    reqs = [q.GenRequest(text_ids=rng.integers(0, 150000, size=int(rng.integers(17, 50))).tolist(), speaker_id=2861, temperature=0.85,
                         max_tokens=frames, seed=i, stream_variant=True, keep_invalid_frames=True) for i in range(batch)]
    for rep in range(3):
        eng.generate_codes_batch(reqs)
        tm = eng.timing()
        if rep:
            print(json.dumps({"batch": batch, "frames": int(tm.frames), "talker_ms": tm.talker_ms, "prefill_ms": tm.prefill_ms,
                              "ms_per_frame_step": (tm.talker_ms - tm.prefill_ms) / frames, "launches": int(tm.kernel_launches)}), flush=True)
    eng.close()
