"""On-box debug: is a row's result independent of how many rows share the launch (m_pad 32 / 64 / 128)?"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
import numpy as np
import qwen3tts_b200 as q
rng = np.random.default_rng(0)
for (K, N) in [(128, 256), (256, 1024), (1024, 2048), (2048, 1024)]:
    x = rng.standard_normal((1, 128, K)).astype(np.float32)
    w = (rng.standard_normal((1, N, K)) / np.sqrt(K)).astype(np.float32)
    base, _ = q.conv_probe(x[:, :1], w, None)
    for M in (2, 24, 32, 33, 48, 64, 65, 128):
        y, _ = q.conv_probe(x[:, :M], w, None)
        d = np.abs(y[0, 0] - base[0, 0]).max()
        print(f"K {K} N {N} M {M}: row 0 max|diff| vs M=1: {d:.3e}")
