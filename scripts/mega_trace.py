"""On-box diagnostics of the persistent frame kernel: per-phase cycle stamps (Q3TTS_MEGA_TRACE) of one 8-frame launch on the
0.6B 4-bit synthetic checkpoint, summarised per phase kind.  Not the bench."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
from oracle import checkpoint

bits = int(sys.argv[1]) if len(sys.argv) > 1 else 4
path = "/tmp/q3tts_mega_trace.bin"
os.environ["Q3TTS_MEGA_TRACE"] = path
import qwen3tts_b200 as q

d = checkpoint.write_checkpoint(f"/tmp/q3tts_06b_{bits}", "0.6b", bits=bits, dtype="bf16", seed=0)
eng = q.Engine(d, max_frames=64, load_codec=False)
fr = eng.generate_codes(q.GenRequest(text_ids=list(range(1000, 1024)), speaker_id=2861, temperature=0.0, max_tokens=16, keep_invalid_frames=True))
tm = eng.timing()
print(json.dumps({"frames": len(fr), "device_ms": tm.device_ms, "prefill_ms": tm.prefill_ms, "ms_per_frame": (tm.device_ms - tm.prefill_ms) / max(1, len(fr))}))
t = np.fromfile(path, dtype=np.int64).reshape(2, -1, 8)
names = {0: "mtp", 1: "qkv", 2: "o", 3: "gate_up", 4: "down", 5: "head", 10: "attention", 20: "sample"}
us = lambda x: x / 1.965e3
for cta in range(2):
    a = t[cta]
    n = int((a[:, 5] != 0).sum())
    a = a[:n]
    kind = a[:, 7] & 0xFF
    poll = a[:, 7] >> 8
    print(f"--- CTA {'0' if cta == 0 else 'grid/2'}: {n} phases, total {us(a[-1,5]-a[0,0]):.1f} us at 1.965 GHz (thread 0 = warp 0)")
    gap = np.zeros(n); gap[1:] = a[1:, 0] - a[:-1, 5]
    for k, name in names.items():
        m = kind == k
        if not m.any():
            continue
        f = lambda x: f"{us(x[m].mean()):6.2f}"
        lin = k < 10
        line = f"{name:10s} n={int(m.sum()):5d}  gap {f(gap)}  stage {f(a[:,1]-a[:,0])} (poll {f(poll)})  body {f(a[:,2]-a[:,1])}"
        if lin:
            mm = m & (a[:, 3] != 0) & (a[:, 4] != 0)
            g = lambda x: f"{us(x[mm].mean()):6.2f}" if mm.any() else "   n/a"
            line += f" [pre {g(a[:,3]-a[:,1])} (mbar {f(a[:,6])}) math {g(a[:,4]-a[:,3])} epilogue {g(a[:,2]-a[:,4])}]"
        line += f"  tail {f(a[:,5]-a[:,2])}  total {f(a[:,5]-a[:,0]+gap)} us"
        print(line)
eng.close()
