"""Probe: does running K independent launch chains concurrently (K handles, K host threads, 64 / K utterances each) beat one chain of 64 rows?
The batched frame step is a chain of ~570 dependent launches at 6-7 us each with DRAM / tensor pipes > 90 % idle."""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
import qwen3tts_b200 as q
from bench import make_requests, ckpt_path, INIT
from oracle import checkpoint

d = ckpt_path("0.6b", 4)
checkpoint.write_checkpoint(d, "0.6b", bits=4, dtype="bf16", seed=0, init=INIT)
frames = 36
total = int(sys.argv[1]) if len(sys.argv) > 1 else 64
KS = tuple(int(k) for k in sys.argv[2].split(',')) if len(sys.argv) > 2 else (1, 2, 4)
for K in KS:
    per = total // K
    engs = [q.Engine(d, max_batch=per, max_frames=64, load_codec=False) for _ in range(K)]
    reqs = make_requests(q, total, frames, 1)
    parts = [reqs[i * per:(i + 1) * per] for i in range(K)]
    def work(i):
        engs[i].generate_codes_batch(parts[i])
    def run():
        th = [threading.Thread(target=work, args=(i,)) for i in range(K)]
        t0 = time.perf_counter()
        [t.start() for t in th]; [t.join() for t in th]
        return time.perf_counter() - t0
    for _ in range(2): run()
    ts = [run() for _ in range(8)]
    best = min(ts)
    print(f"K={K} chains x {per} rows: {best*1e3:.1f} ms per {frames}-frame batch -> {best*1e3/frames:.3f} ms per frame-step of {total} utterances "
          f"({total*frames*0.08/best:.0f} talker audio-s/s)", flush=True)
    for e in engs: e.close()
