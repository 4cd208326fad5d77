"""MLX group-affine quantisation layout, restated in numpy (oracle; test infrastructure only).

The reference never spells this arithmetic out: it calls `MLXNN.QuantizedLinear`
(`Model/QuantizedLayerFactory.swift:49-66`) and `dequantized(...)` (`Model/Qwen3Talker.swift:156`)
from the un-vendored dependency ml-explore/mlx-swift 0.30.3.  What follows restates MLX's
published `quantize` / `dequantize` semantics (mode "affine", bits in {4, 8}):

* weight `[out, in]`, grouped along `in` in groups of `group_size`;
* each uint32 packs `32 / bits` consecutive elements, element j at bit offset `j * bits`
  (least-significant first);
* `scales`, `biases` are `[out, in / group_size]` in the model float dtype;
* `w ~= scale * q + bias`.

Bit-exact dequantisation contract used by the parity tests:
`deq32 = fp32(scale) * fp32(q) + fp32(bias)` as two separately rounded fp32 operations
(mul, then add), and `deqT = round_to_nearest_even_T(deq32)` for T in {fp16, bf16}.
"""
from __future__ import annotations

import numpy as np


def _to_f32(a) -> np.ndarray:
    """Accept numpy fp32/fp16 arrays or torch tensors (bf16 included) and return fp32 numpy."""
    try:
        import torch

        if isinstance(a, torch.Tensor):
            return a.to(torch.float32).cpu().numpy()
    except ImportError:  # pragma: no cover
        pass
    return np.asarray(a, dtype=np.float32)


def round_to_dtype(a: np.ndarray, dtype: str) -> np.ndarray:
    """Round an fp32 array to `dtype` ('f32' | 'f16' | 'bf16') and return it widened back to fp32."""
    a = np.ascontiguousarray(a, dtype=np.float32)
    if dtype == "f32":
        return a
    if dtype == "f16":
        return a.astype(np.float16).astype(np.float32)
    if dtype == "bf16":
        u = a.view(np.uint32).astype(np.uint64)
        # round-to-nearest-even on the 16 dropped bits; NaN/Inf do not occur in our weights
        rounded = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
        return rounded.astype(np.uint32).view(np.float32)
    raise ValueError(dtype)


def quantize(w: np.ndarray, group_size: int = 64, bits: int = 4, scale_dtype: str = "bf16"):
    """MLX `quantize(w, group_size, bits)` (affine).  Returns (packed uint32, scales f32, biases f32).

    Scales and biases are rounded to `scale_dtype` BEFORE the integer codes are chosen, so the
    triple on disk is self-consistent; they are returned as fp32 holding those rounded values.
    """
    assert bits in (4, 8) and 32 % bits == 0
    q, s, b = quantize_codes(w, group_size, bits, scale_dtype)
    return pack(q, bits), s, b


def quantize_codes(w: np.ndarray, group_size: int = 64, bits: int = 4, scale_dtype: str = "bf16"):
    """The quantiser itself for any bit width (4, 6, 8): integer codes [out, in] (uint32), scales, biases (fp32 holding `scale_dtype` values).
    `pack(codes, 8)` is the 8-bit container the engine keeps runtime-quantised 4/6-bit leaves in (Qwen3TTSPipeline.swift:961-980)."""
    w = np.ascontiguousarray(w, dtype=np.float32)
    out, inn = w.shape
    assert inn % group_size == 0
    g = w.reshape(out, inn // group_size, group_size)
    n_bins = np.float32((1 << bits) - 1)
    w_max = g.max(axis=-1)
    w_min = g.min(axis=-1)
    side = np.abs(w_min) > np.abs(w_max)
    s = np.maximum((w_max - w_min) / n_bins, np.float32(1e-7)).astype(np.float32)
    s = np.where(side, s, -s).astype(np.float32)
    edge = np.where(side, w_min, w_max).astype(np.float32)
    q0 = np.rint(edge / s)
    s = np.where(q0 != 0, edge / np.where(q0 != 0, q0, 1), s).astype(np.float32)
    b = np.where(q0 == 0, np.float32(0), edge).astype(np.float32)
    s = round_to_dtype(s, scale_dtype)
    b = round_to_dtype(b, scale_dtype)
    s_safe = np.where(s == 0, np.float32(1e-7), s)
    q = np.clip(np.rint((g - b[..., None]) / s_safe[..., None]), 0, n_bins).astype(np.uint32)
    return q.reshape(out, inn), s, b


def runtime_bits(path: str) -> int:
    """applyMixedQuantization's rule (Qwen3TTSPipeline.swift:966-980): 6 bits for embeddings, q/k/v projections and the heads, else 4."""
    p = path.lower()
    return 6 if any(k in p for k in ("embed", "qproj", "kproj", "vproj", "q_proj", "k_proj", "v_proj", "lm_head", "codec_head")) else 4


def fake_quantize(w: np.ndarray, bits: int, scale_dtype: str = "bf16", out_dtype: str = "f32") -> np.ndarray:
    """dequantize(quantize(w)) for group 64 and any bit width: what a runtime-quantised leaf multiplies with."""
    q, s, b = quantize_codes(w, 64, bits, scale_dtype)
    return dequantize(pack(q, 8), s, b, 64, 8, out_dtype)


def pack(q: np.ndarray, bits: int) -> np.ndarray:
    """[out, in] integer codes -> [out, in*bits/32] uint32, element j at bit offset j*bits."""
    per = 32 // bits
    out, inn = q.shape
    q = q.astype(np.uint32).reshape(out, inn // per, per)
    shifts = (np.arange(per, dtype=np.uint32) * np.uint32(bits))[None, None, :]
    return np.bitwise_or.reduce(q << shifts, axis=-1).astype(np.uint32)


def unpack(packed: np.ndarray, bits: int) -> np.ndarray:
    """Inverse of `pack`: [out, in*bits/32] uint32 -> [out, in] uint32 codes."""
    per = 32 // bits
    packed = np.ascontiguousarray(packed).view(np.uint32)
    out, words = packed.shape
    shifts = (np.arange(per, dtype=np.uint32) * np.uint32(bits))[None, None, :]
    mask = np.uint32((1 << bits) - 1)
    return ((packed[:, :, None] >> shifts) & mask).reshape(out, words * per)


def dequantize(packed, scales, biases, group_size: int = 64, bits: int = 4, out_dtype: str = "f32") -> np.ndarray:
    """MLX `dequantized(w, scales, biases, group_size, bits)` -> fp32 array holding values of `out_dtype`.

    deq32 = fp32(scale) * q  (rounded)  + fp32(bias)  (rounded); deqT = round(deq32).
    """
    q = unpack(np.asarray(packed), bits).astype(np.float32)
    s = _to_f32(scales)
    b = _to_f32(biases)
    out, inn = q.shape
    g = q.reshape(out, inn // group_size, group_size)
    prod = (g * s[..., None]).astype(np.float32)
    deq = (prod + b[..., None]).astype(np.float32).reshape(out, inn)
    return round_to_dtype(deq, out_dtype)


def quantized_matmul(x: np.ndarray, packed, scales, biases, group_size=64, bits=4) -> np.ndarray:
    """y = x @ dequant(W)^T with fp32 accumulation (MLX `quantized_matmul(transpose=true)`)."""
    w = dequantize(packed, scales, biases, group_size, bits, "f32")
    return np.asarray(x, dtype=np.float32) @ w.T
