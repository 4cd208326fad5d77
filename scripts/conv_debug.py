"""On-box debug: where does a large conv_probe case differ from numpy?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import qwen3tts_b200 as q
from test_gpu_gemm_tc import ref_conv
B, T, cin, N, ntap, dil = [int(a) for a in sys.argv[1:7]]
rng = np.random.default_rng(B * 7 + T + cin + N + ntap)
x = rng.standard_normal((B, T, cin)).astype(np.float32)
w = (rng.standard_normal((ntap, N, cin)) / np.sqrt(cin * ntap)).astype(np.float32)
bias = rng.standard_normal(N).astype(np.float32) * 0.1
y32, y16 = q.conv_probe(x, w, bias, ntap=ntap, dil=dil)
want = ref_conv(x, w, bias, ntap, dil)
err = np.abs(y32 - want).max(axis=2)  # [B, T]
bad = np.argwhere(err > 1e-3)
print("bad rows:", len(bad), "of", B * T)
if len(bad):
    bs = sorted(set(int(b) for b, _ in bad))
    print("batches with bad rows:", bs[:20])
    for b in bs[:3]:
        ts = [int(t) for bb, t in bad if bb == b]
        print(f"  batch {b}: {len(ts)} bad rows, t in [{min(ts)}, {max(ts)}], tiles {sorted(set(t // 128 for t in ts))[:20]}")
    b, t = bad[0]
    print("first bad row", b, t, "got", y32[b, t, :4], "want", want[b, t, :4])
