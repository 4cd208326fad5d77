"""Regenerates tests/golden/tiny_golden.npz from the CPU oracle (the reference itself cannot run here: no Swift / MLX;
its tests hold no numeric vectors — parity is unpinned, these fixtures pin the ORACLE against drift and give the GPU tests
a box-independent target).  Run:  python tests/golden/make_golden.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import TEXT_IDS, ckpt  # noqa: E402
from oracle import codec as ocodec, mlx_quant, talker as otalker  # noqa: E402

out = {}
rng = np.random.default_rng(42)
w = (rng.standard_normal((8, 128)) * 0.05).astype(np.float32)
for bits in (4, 8):
    p, s, b = mlx_quant.quantize(w, 64, bits, "bf16")
    out[f"q{bits}_packed"], out[f"q{bits}_scales"], out[f"q{bits}_biases"] = p, s, b
    out[f"q{bits}_deq_f32"] = mlx_quant.dequantize(p, s, b, 64, bits, "f32")
    out[f"q{bits}_deq_f16"] = mlx_quant.dequantize(p, s, b, 64, bits, "f16")
out["q_w"] = w
for name, bits in (("tiny8", 8), ("tiny4", 4)):
    d = ckpt("tiny", bits)
    orc = otalker.TalkerOracle(d)
    rec = {}
    orc.generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=8), record=rec, filter_invalid=False)
    out[f"{name}_frames"] = np.asarray(rec["raw_frames"], np.int32)
    out[f"{name}_code0_logits"] = rec["code0_logits"][:2]
    out[f"{name}_cp_logits"] = rec["cp_logits"][:1, :3]
    out[f"{name}_margins"] = rec["margins"]
cd = ocodec.load_codec(ckpt("tiny", 8))
codes = np.random.default_rng(7).integers(0, 2048, size=(1, 4, 16)).astype(np.int32)
tc = torch.as_tensor(codes).transpose(1, 2).contiguous()
first, rest = cd.rvq_embed(tc)
out["codec_codes"] = codes
out["codec_first"], out["codec_rest"] = first.numpy(), rest.numpy()
out["codec_pcm"] = cd.decode(tc).reshape(-1).numpy()
out["text_ids"] = np.asarray(TEXT_IDS, np.int32)
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tiny_golden.npz")
np.savez_compressed(path, **out)
print("wrote", path, os.path.getsize(path), "bytes")
