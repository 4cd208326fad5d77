"""Quick on-box probe: 0.6B synthetic checkpoint -> ms/frame (graph vs eager), codec decode time.  Not the bench."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
t0 = time.time()
d = checkpoint.write_checkpoint(f"/tmp/q3tts_06b_{bits}", "0.6b", bits=bits, dtype="bf16", seed=0)
print(f"checkpoint ready in {time.time()-t0:.1f}s", flush=True)
t0 = time.time()
eng = q.Engine(d, max_frames=512)
print(f"engine load {time.time()-t0:.1f}s, device bytes {eng.info.device_bytes/1e9:.2f} GB", flush=True)
ids = list(range(1000, 1024))
for graph in (True, False):
    e = eng if graph else q.Engine(d, max_frames=512, use_cuda_graph=False)
    for rep in range(2):
        t0 = time.time()
        fr = e.generate_codes(q.GenRequest(text_ids=ids, speaker_id=2861, temperature=0.0, max_tokens=64, keep_invalid_frames=True))
        wall = time.time() - t0
        tm = e.timing()
        print(json.dumps({"graph": graph, "rep": rep, "frames": len(fr), "wall_ms": wall*1e3, "device_ms": tm.device_ms, "prefill_ms": tm.prefill_ms,
                          "ms_per_frame": (tm.device_ms - tm.prefill_ms)/max(1,len(fr)), "launches": tm.kernel_launches,
                          "bytes_per_frame": tm.weight_bytes_per_frame}), flush=True)
codes = np.random.default_rng(0).integers(0, 2048, size=(1, 26, 16)).astype(np.int32)
for rep in range(3):
    t0 = time.time(); pcm = eng.decode(codes); wall = time.time() - t0
    tm = eng.timing()
    print(json.dumps({"codec_T": 26, "wall_ms": wall*1e3, "device_ms": tm.device_ms, "decode_ms": tm.decode_ms, "launches": tm.kernel_launches}), flush=True)
