"""CPU: the MLX affine quantisation restatement (oracle/mlx_quant.py) against independent formulations."""
import numpy as np
import pytest

from oracle import mlx_quant


@pytest.mark.parametrize("bits", [4, 8])
def test_pack_layout_lsb_first(bits):
    per = 32 // bits
    q = np.arange(per, dtype=np.uint32)[None, :] % (1 << bits)
    w = mlx_quant.pack(q, bits)
    assert w.shape == (1, 1)
    # element j sits at bit offset j*bits
    want = 0
    for j in range(per):
        want |= int(q[0, j]) << (j * bits)
    assert int(w[0, 0]) == want
    assert np.array_equal(mlx_quant.unpack(w, bits), q)


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("group", [32, 64, 128])
def test_pack_unpack_roundtrip(bits, group):
    rng = np.random.default_rng(bits + group)
    q = rng.integers(0, 1 << bits, size=(17, 256)).astype(np.uint32)
    assert np.array_equal(mlx_quant.unpack(mlx_quant.pack(q, bits), bits), q)


@pytest.mark.parametrize("bits", [4, 8])
@pytest.mark.parametrize("sdt", ["f32", "bf16", "f16"])
def test_quantize_error_bound(bits, sdt):
    """Round-trip error.  Interior points land within |scale|/2; MLX re-fits the scale so the dominant edge is exact
    (s = edge / round(edge / s)), which can leave the opposite end of the group up to ~one step outside the code range
    (measured max 0.9995 |scale|), so the bound that holds for every element is one step."""
    rng = np.random.default_rng(0)
    w = (rng.standard_normal((64, 512)) * 0.02).astype(np.float32)
    packed, s, b = mlx_quant.quantize(w, 64, bits, sdt)
    deq = mlx_quant.dequantize(packed, s, b, 64, bits, "f32")
    step = np.repeat(np.abs(s), 64, axis=1)
    slack = {"f32": 1e-7, "bf16": 2 ** -8, "f16": 2 ** -11}[sdt] * np.abs(w).max() * 4
    assert np.all(np.abs(deq - w) <= step * 1.0 + slack + 1e-7)
    assert np.median(np.abs(deq - w) / step) <= 0.30


def test_quantize_edges_are_exact():
    """MLX picks the edge (w_min or w_max by magnitude) so that it is represented exactly: q0 = round(edge/s), s = edge/q0."""
    rng = np.random.default_rng(1)
    w = (rng.standard_normal((8, 128)) * 0.05).astype(np.float32)
    packed, s, b = mlx_quant.quantize(w, 64, 4, "f32")
    deq = mlx_quant.dequantize(packed, s, b, 64, 4, "f32").reshape(8, 2, 64)
    g = w.reshape(8, 2, 64)
    edge = np.where(np.abs(g.min(-1)) > np.abs(g.max(-1)), g.min(-1), g.max(-1))
    # the bias IS the edge value (when q0 != 0), and code 0 decodes to it
    assert np.allclose(b, edge, rtol=0, atol=0)
    idx = np.abs(g - edge[..., None]).argmin(-1)
    assert np.allclose(np.take_along_axis(deq, idx[..., None], -1)[..., 0], edge, rtol=1e-6, atol=1e-8)


def test_zero_group():
    w = np.zeros((2, 64), np.float32)
    packed, s, b = mlx_quant.quantize(w, 64, 4, "bf16")
    assert np.all(mlx_quant.dequantize(packed, s, b, 64, 4, "f32") == 0)


def test_dequant_two_roundings_contract():
    """deq32 = round(scale*q) then round(+bias), NOT a fused multiply-add: pick a case where they differ."""
    s = np.array([[np.float32(1.0000001)]], np.float32)
    b = np.array([[np.float32(-3.0)]], np.float32)
    packed = mlx_quant.pack(np.full((1, 64), 3, np.uint32), 8)
    got = mlx_quant.dequantize(packed, s, b, 64, 8, "f32")[0, 0]
    prod = np.float32(np.float32(1.0000001) * np.float32(3.0))
    assert got == np.float32(prod + np.float32(-3.0))
    assert got != np.float32(np.float64(np.float32(1.0000001)) * 3.0 - 3.0)  # what an FMA would return


def test_round_to_bf16_matches_torch():
    import torch

    rng = np.random.default_rng(2)
    a = rng.standard_normal(10000).astype(np.float32) * 3
    want = torch.from_numpy(a).to(torch.bfloat16).to(torch.float32).numpy()
    assert np.array_equal(mlx_quant.round_to_dtype(a, "bf16"), want)


def test_quantized_matmul_equals_dense():
    rng = np.random.default_rng(3)
    w = (rng.standard_normal((32, 128)) * 0.02).astype(np.float32)
    x = rng.standard_normal((3, 128)).astype(np.float32)
    packed, s, b = mlx_quant.quantize(w, 64, 8, "bf16")
    y = mlx_quant.quantized_matmul(x, packed, s, b, 64, 8)
    assert np.allclose(y, x @ mlx_quant.dequantize(packed, s, b, 64, 8).T)
    assert np.abs(y - x @ w.T).max() < 5e-3


# ------------------------------------------------------------------------------------------------ runtime mixed 4/6-bit quantisation (f4)
@pytest.mark.parametrize("bits", [4, 6, 8])
@pytest.mark.parametrize("dtype", ["bf16", "f16", "f32"])
def test_quantize_codes_any_width_is_self_consistent(bits, dtype):
    """`quantize_codes` (the quantiser behind applyMixedQuantization, Qwen3TTSPipeline.swift:961-980): codes stay inside [0, 2^bits), the
    error is at most one step (+ the rounding of scale / bias to the weight dtype), and
    for 4 / 8 bits it is the packed quantiser the golden vectors pin."""
    from oracle import mlx_quant as m

    rng = np.random.default_rng(bits * 7 + len(dtype))
    w = m.round_to_dtype((rng.standard_normal((24, 256)) * 0.05).astype(np.float32), dtype)
    q, s, b = m.quantize_codes(w, 64, bits, dtype)
    assert q.min() >= 0 and q.max() <= (1 << bits) - 1 and q.shape == w.shape
    deq = m.dequantize(m.pack(q, 8), s, b, 64, 8, "f32")
    step = np.abs(s)[..., None].repeat(64, -1).reshape(w.shape)
    ulp = {"bf16": 2.0 ** -8, "f16": 2.0 ** -11, "f32": 2.0 ** -23}[dtype]  # relative rounding of scale and bias to the weight dtype
    bias = np.abs(b)[..., None].repeat(64, -1).reshape(w.shape)
    # half a step inside the grid; up to one step at the far edge (re-deriving the scale from the kept edge can shorten the grid)
    tol = 1.0 * step + ((1 << bits) - 1) * step * ulp + bias * ulp
    assert np.all(np.abs(deq - w) <= tol * 1.001 + 1e-7)
    assert np.array_equal(deq, m.fake_quantize(w, bits, dtype))
    if bits in (4, 8):
        p2, s2, b2 = m.quantize(w, 64, bits, dtype)
        assert np.array_equal(m.unpack(p2, bits), q) and np.array_equal(s2, s) and np.array_equal(b2, b)


def test_runtime_bits_rule():
    from oracle import mlx_quant as m

    six = ["text_embedding", "codec_embedding", "code_predictor.codec_embedding.3", "layers.0.self_attn.q_proj", "layers.5.self_attn.k_proj",
           "code_predictor.layers.1.self_attn.v_proj", "code_predictor.lm_head.7", "codec_head"]
    four = ["layers.0.self_attn.o_proj", "layers.0.mlp.gate_proj", "layers.0.mlp.up_proj", "layers.0.mlp.down_proj", "text_projection.linear_fc1",
            "code_predictor.small_to_mtp_projection"]
    assert all(m.runtime_bits(p) == 6 for p in six) and all(m.runtime_bits(p) == 4 for p in four)


@pytest.mark.parametrize("bits,group", [(4, 64), (8, 64), (4, 128), (8, 32)])
def test_dequantize_layout_equals_hf_metal_affine(bits, group):
    """Third-party statement of the MLX affine LAYOUT (Appendix C of SURVEY.md: uint32 words, element j of a word at bit offset j * bits, LSB
    first; scales / biases per group along the input dimension; w = scale * q + bias): HuggingFace transformers' Metal integration dequantises
    `quantization-mlx` checkpoints with `_affine_dequantize_tensor`.  Same packed words, scales and biases in -> same weights out, and its
    packer produces the words `oracle.mlx_quant.pack` produces."""
    import numpy as np
    import torch

    mq_hf = pytest.importorskip("transformers.integrations.metal_quantization")
    from oracle import mlx_quant as mq

    rng = np.random.default_rng(bits * 100 + group)
    w = (rng.standard_normal((24, 256)) * 0.05).astype(np.float32)
    packed, scales, biases = mq.quantize(w, group, bits, "f32")
    mine = mq.dequantize(packed, scales, biases, group, bits, "f32")
    s32 = torch.from_numpy(np.asarray(scales, dtype=np.float32)); b32 = torch.from_numpy(np.asarray(biases, dtype=np.float32))
    theirs = mq_hf._affine_dequantize_tensor(torch.from_numpy(packed.view(np.int32)), s32, b32, group, bits).numpy()
    assert np.array_equal(mine, theirs)
    # the packer: HF quantises with a plain min/max rule (not MLX's edge-preserving one), so compare the PACKING of given codes
    q = rng.integers(0, 1 << bits, size=(24, 256)).astype(np.uint32)
    per = 32 // bits
    hf_words = torch.zeros(24, 256 // per, dtype=torch.int32)
    qi = torch.from_numpy(q.astype(np.int32))
    for i in range(per):  # the loop of _affine_quantize_tensor
        hf_words |= qi[:, i::per] << (bits * i)
    assert np.array_equal(mq.pack(q, bits).view(np.int32), hf_words.numpy())
