"""On-box debug: where do batched (tensor-core) results leave the single-utterance results?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
os.environ["Q3TTS_TC_MIN_ROWS"] = "1"
import numpy as np
from oracle import checkpoint
import qwen3tts_b200 as q
d = checkpoint.write_checkpoint("/tmp/q3tts_dbg_tiny8", "tiny", bits=8)
e1 = q.Engine(d, max_frames=256, load_codec=False)
eb = q.Engine(d, max_frames=256, load_codec=False, max_batch=24)
reqs = [q.GenRequest(text_ids=[11, 21, 22] + list(range(60 + i, 60 + i + 8 + (i % 5))), speaker_id=[2861, 3066, -1, 2873][i % 4],
                     temperature=0.0 if i % 2 else 0.8, top_k=0 if i % 3 else 40, seed=i, max_tokens=8 + (i % 7), keep_invalid_frames=True)
        for i in range(40)]
singles = [e1.generate_codes(r).tolist() for r in reqs]
batch = [b.tolist() for b in eb.generate_codes_batch(reqs)]
for i, (a, b) in enumerate(zip(singles, batch)):
    if a != b:
        f = next((k for k in range(min(len(a), len(b))) if a[k] != b[k]), -1)
        g = next((k for k in range(16) if f >= 0 and a[f][k] != b[f][k]), -1)
        print(f"req {i}: temp {reqs[i].temperature} top_k {reqs[i].top_k} len {len(a)}/{len(b)} first diff frame {f} group {g}")
print("mismatching:", sum(a != b for a, b in zip(singles, batch)), "of", len(reqs))
# the same with every request alone in the batch engine (prefill rows <= 128 -> same kernels as the single engine)
alone = [eb.generate_codes_batch([r])[0].tolist() for r in reqs]
print("batch engine, one request at a time, mismatching:", sum(a != b for a, b in zip(singles, alone)))
# minimal reproduction: request 2 together with k other requests
for k in (1, 3, 7, 15, 23):
    sub = [reqs[2]] + [reqs[i] for i in range(3, 3 + k)]
    out = eb.generate_codes_batch(sub)[0].tolist()
    print(f"req 2 with {k} others: {'same' if out == singles[2] else 'DIFFERENT'}")
# and with longer generations (more chances for a real M-dependence to show)
long = [q.GenRequest(text_ids=r.text_ids, speaker_id=r.speaker_id, temperature=0.0, max_tokens=40, keep_invalid_frames=True) for r in reqs[:24]]
s_long = [e1.generate_codes(r).tolist() for r in long]
b_long = [b.tolist() for b in eb.generate_codes_batch(long)]
print("greedy 40 frames x 24: mismatching", sum(a != b for a, b in zip(s_long, b_long)))
