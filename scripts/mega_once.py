"""On-box: one short batch-1 generation on the 0.6B 4-bit synthetic checkpoint (persistent frame kernel); used under ncu."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
from oracle import checkpoint
import qwen3tts_b200 as q

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
frames = int(sys.argv[2]) if len(sys.argv) > 2 else 8
d = checkpoint.write_checkpoint(f"/tmp/q3tts_06b_{bits}", "0.6b", bits=bits, dtype="bf16", seed=0)
eng = q.Engine(d, max_frames=64, load_codec=False)
for rep in range(2):
    fr = eng.generate_codes(q.GenRequest(text_ids=list(range(1000, 1024)), speaker_id=2861, temperature=0.0, max_tokens=frames, keep_invalid_frames=True))
    tm = eng.timing()
    print(json.dumps({"frames": len(fr), "device_ms": tm.device_ms, "prefill_ms": tm.prefill_ms, "ms_per_frame": (tm.device_ms - tm.prefill_ms) / max(1, len(fr)),
                      "persistent_launches": tm.persistent_launches}))
eng.close()
