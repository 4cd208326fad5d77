"""GPU: the tcgen05/TMEM/TMA implicit-GEMM kernel against a float64 numpy restatement of the same causal conv
(operands rounded to fp16 exactly as the kernel sees them, so the only difference is fp32 accumulation order)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def ref_conv(x, w, bias, ntap, dil, act=0, swiglu=False, res=None, scale=None):
    x = x.astype(np.float16).astype(np.float64)
    w = w.astype(np.float16).astype(np.float64)
    B, T, cin = x.shape
    N = w.shape[1]
    y = np.zeros((B, T, N))
    for tap in range(ntap):
        sh = (ntap - 1 - tap) * dil
        if sh >= T:
            continue
        xs = np.zeros_like(x)
        xs[:, sh:] = x[:, : T - sh]
        y += xs @ w[tap].T
    if bias is not None:
        y += bias
    if act == 1:
        from math import erf
        y = 0.5 * y * (1 + np.vectorize(erf)(y / np.sqrt(2)))
    elif act == 2:
        y = y / (1 + np.exp(-y))
    if swiglu:
        g, u = y[..., 0::2], y[..., 1::2]
        y = g / (1 + np.exp(-g)) * u
    if res is not None:
        y = res + (scale if scale is not None else 1.0) * y
    return y


CASES = [
    # B, T, cin, N, ntap, dil
    (1, 128, 64, 64, 1, 1),      # exactly one tile, one k-block
    (1, 5, 64, 32, 1, 1),        # rows beyond T are zero-filled and masked
    (2, 200, 128, 96, 1, 1),     # N = 96 (3 x 32 columns), two M tiles per batch item, K = 2 blocks
    (1, 300, 96, 96, 7, 1),      # cin = 96: second k-block half zero-filled; 7 causal taps
    (2, 257, 192, 192, 7, 3),    # dilation 3, batch boundary must not leak across items
    (1, 140, 96, 96, 7, 9),      # dilation 9
    (1, 64, 512, 1024, 3, 1),    # pre_conv shape, N > 256 -> 4 N tiles of 256
    (3, 50, 1536, 768, 2, 1),    # polyphase transposed conv (2 taps), long K
    (1, 40, 1024, 4096, 1, 1),   # pwconv1
    (1, 130, 1024, 384, 1, 1),   # N = 384 -> 2 tiles of 192
    (64, 1, 1024, 3072, 1, 1),   # batched decode: 64 utterances x 1 row each (T = 1 per batch item)
    (1, 64, 1024, 2048, 1, 1),   # the same as one 64-row matrix
    # <= 128 rows and one tap route to the split-K cluster kernel (gemm_skinny.cu)
    (1, 17, 2048, 1024, 1, 1),   # o_proj shape: 8 weight tiles x 8 K slices, m_pad 32 (mc = 4)
    (1, 64, 3072, 1024, 1, 1),   # down_proj shape: 48 k-blocks over 8 CTAs, ring wraps
    (1, 128, 1024, 2048, 1, 1),  # m_pad 128 (code-predictor pass 0: two rows per utterance)
    (1, 100, 1024, 6144, 1, 1),  # gate|up shape: 48 tiles x 2 slices of 8 k-blocks
    (2, 33, 96, 96, 1, 1),       # N = 96 < one weight tile, cin = 96: second k-block half zero-filled; 66 rows -> m_pad 128
    (1, 40, 64, 160, 1, 1),      # one k-block: no split possible
    (1, 3, 320, 32, 1, 1),       # 5 k-blocks over 4 slices: uneven split
    # many tiles: the persistent schedule (ring running across tile boundaries, two TMEM accumulators)
    (16, 9600, 96, 96, 7, 3),    # 1200 tiles over 296 CTAs (2 per SM), thin-channel vocoder shape
    (5, 16000, 192, 192, 1, 1),  # 625 tiles over 148 CTAs (256-column accumulators: one CTA per SM)
]


@pytest.mark.parametrize("B,T,cin,N,ntap,dil", CASES)
def test_tc_conv_matches_numpy(B, T, cin, N, ntap, dil):
    import qwen3tts_b200 as q

    rng = np.random.default_rng(B * 7 + T + cin + N + ntap)
    x = rng.standard_normal((B, T, cin)).astype(np.float32)
    w = (rng.standard_normal((ntap, N, cin)) / np.sqrt(cin * ntap)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    y32, y16 = q.conv_probe(x, w, bias, ntap=ntap, dil=dil)
    want = ref_conv(x, w, bias, ntap, dil)
    err = np.abs(y32 - want).max()
    print(f"B{B} T{T} cin{cin} N{N} taps{ntap} dil{dil}: max err {err:.2e}")
    assert err < 2e-4 * max(1.0, np.abs(want).max())
    assert np.abs(y16 - want).max() < 2e-3 * max(1.0, np.abs(want).max())


def test_tc_persistent_epilogue_residual_in_place():
    """Persistent schedule with the residual epilogue (per-channel scale) over many tiles; rows past T in the last tile of each
    batch item are masked."""
    import qwen3tts_b200 as q

    rng = np.random.default_rng(11)
    B, T, cin, N = 7, 22000, 96, 96   # 172 tiles per item (the last one 112 rows), 1204 tiles
    x = rng.standard_normal((B, T, cin)).astype(np.float32)
    w = (rng.standard_normal((1, N, cin)) / np.sqrt(cin)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    res = rng.standard_normal((B, T, N)).astype(np.float32)
    scale = rng.uniform(0.3, 0.7, N).astype(np.float32)
    y32, y16 = q.conv_probe(x, w, bias, res=res, scale=scale)
    want = ref_conv(x, w, bias, 1, 1, res=res, scale=scale)
    assert np.abs(y32 - want).max() < 2e-4 * max(1.0, np.abs(want).max())
    assert np.abs(y16 - want).max() < 4e-3 * max(1.0, np.abs(want).max())


def test_tc_epilogues():
    import qwen3tts_b200 as q

    rng = np.random.default_rng(0)
    B, T, cin, N = 2, 70, 128, 256
    x = rng.standard_normal((B, T, cin)).astype(np.float32)
    w = (rng.standard_normal((1, N, cin)) / np.sqrt(cin)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    # GELU
    y32, _ = q.conv_probe(x, w, bias, act=1)
    assert np.abs(y32 - ref_conv(x, w, bias, 1, 1, act=1)).max() < 2e-4
    # residual + per-channel scale
    res = rng.standard_normal((B, T, N)).astype(np.float32)
    scale = rng.uniform(0.3, 0.7, N).astype(np.float32)
    y32, _ = q.conv_probe(x, w, bias, res=res, scale=scale)
    assert np.abs(y32 - ref_conv(x, w, bias, 1, 1, res=res, scale=scale)).max() < 2e-4
    # SwiGLU with interleaved (gate, up) columns
    y32, y16 = q.conv_probe(x, w, None, swiglu=True)
    want = ref_conv(x, w, None, 1, 1, swiglu=True)
    assert y32.shape == (B, T, N // 2) and np.abs(y32 - want).max() < 2e-4
    # SnakeBeta fused into the fp16 copy only
    ea = np.exp(rng.standard_normal(64) * 0.3).astype(np.float32)
    ieb = (1 / (np.exp(rng.standard_normal(64) * 0.3) + 1e-9)).astype(np.float32)
    y32, y16 = q.conv_probe(x, w, bias, snake=(ea, ieb))
    base = ref_conv(x, w, bias, 1, 1)
    ch = np.arange(N) % 64
    assert np.abs(y32 - base).max() < 2e-4
    assert np.abs(y16 - (base + ieb[ch] * np.sin(base * ea[ch]) ** 2)).max() < 4e-3


@pytest.mark.parametrize("M", [20, 64, 128])
def test_tc_skinny_epilogues(M):
    """Same epilogues through the <= 128-row cluster kernel (rows owned by different CTAs of the cluster)."""
    import qwen3tts_b200 as q

    rng = np.random.default_rng(M)
    cin, N = 1024, 512
    x = rng.standard_normal((1, M, cin)).astype(np.float32)
    w = (rng.standard_normal((1, N, cin)) / np.sqrt(cin)).astype(np.float32)
    bias = rng.standard_normal(N).astype(np.float32) * 0.1
    y32, y16 = q.conv_probe(x, w, bias, act=1)
    want = ref_conv(x, w, bias, 1, 1, act=1)
    assert np.abs(y32 - want).max() < 2e-4 and np.abs(y16 - want).max() < 4e-3
    y32, _ = q.conv_probe(x, w, bias, act=2)
    assert np.abs(y32 - ref_conv(x, w, bias, 1, 1, act=2)).max() < 2e-4
    res = rng.standard_normal((1, M, N)).astype(np.float32)
    scale = rng.uniform(0.3, 0.7, N).astype(np.float32)
    y32, _ = q.conv_probe(x, w, bias, res=res, scale=scale)
    assert np.abs(y32 - ref_conv(x, w, bias, 1, 1, res=res, scale=scale)).max() < 2e-4
    y32, _ = q.conv_probe(x, w, None, res=res)
    assert np.abs(y32 - ref_conv(x, w, None, 1, 1, res=res)).max() < 2e-4
    y32, y16 = q.conv_probe(x, w, None, swiglu=True)
    want = ref_conv(x, w, None, 1, 1, swiglu=True)
    assert y32.shape == (1, M, N // 2) and np.abs(y32 - want).max() < 2e-4 and np.abs(y16 - want).max() < 4e-3


def test_tc_skinny_is_deterministic():
    import qwen3tts_b200 as q

    rng = np.random.default_rng(5)
    x = rng.standard_normal((1, 64, 2048)).astype(np.float32)
    w = (rng.standard_normal((1, 1024, 2048)) / 45).astype(np.float32)
    a, _ = q.conv_probe(x, w, None)
    b, _ = q.conv_probe(x, w, None)
    assert np.array_equal(a, b)


def test_simt_conv_matches_numpy():
    import qwen3tts_b200 as q

    rng = np.random.default_rng(1)
    x = rng.standard_normal((2, 37, 12)).astype(np.float32)
    w = (rng.standard_normal((7, 24, 12)) / 9).astype(np.float32)
    bias = rng.standard_normal(24).astype(np.float32)
    y32, _ = q.conv_probe(x, w, bias, ntap=7, dil=3, use_tensor_cores=False)
    xs, ws = x.astype(np.float64), w.astype(np.float64)
    want = np.zeros((2, 37, 24))
    for tap in range(7):
        sh = (6 - tap) * 3
        if sh < 37:
            z = np.zeros_like(xs)
            z[:, sh:] = xs[:, : 37 - sh]
            want += z @ ws[tap].T
    assert np.abs(y32 - (want + bias)).max() < 1e-4


# ------------------------------------------------------------------------------------------------ dequant-fused GEMM (a9)
@pytest.mark.parametrize("bits,sdt,group", [(4, "bf16", 64), (8, "bf16", 64), (4, "f16", 32), (8, "f32", 128), (4, "f32", 128), (8, "f16", 32)])
@pytest.mark.parametrize("shape", [(1, 1024, 1024), (3, 4096, 1024), (24, 1024, 2048), (64, 1024, 3072), (64, 3072, 1024), (100, 2048, 1024), (128, 256, 512),
                                   (7, 128, 256), (33, 2048, 6144), (5, 160, 256)])
def test_dequant_fused_tensor_core_gemm(bits, sdt, group, shape):
    """MLX-packed weights streamed and dequantised INSIDE the tcgen05 GEMM (csrc/gemm_skinny_q.cu) == fp16(x) . fp16(dequant(W))^T:
    the operand bits are those of q3tts_dequantize(out_dtype=f16), accumulation is fp32."""
    import qwen3tts_b200 as q
    from oracle import mlx_quant

    m, out_f, in_f = shape
    rng = np.random.default_rng(m * 7 + out_f + bits + group)
    w = (rng.standard_normal((out_f, in_f)) * 0.02).astype(np.float32)
    w[5] = 0.0
    x = rng.standard_normal((m, in_f)).astype(np.float32)
    packed, s, b = mlx_quant.quantize(w, group, bits, sdt)
    wd = mlx_quant.dequantize(packed, s, b, group, bits, "f16").astype(np.float64)
    x16 = x.astype(np.float16).astype(np.float64)
    got = q.quantized_matmul_tc(x, packed, s, b, group, bits, sdt)
    want = x16 @ wd.T
    bound = 4e-6 * (np.abs(x16) @ np.abs(wd).T) + 1e-6  # fp32 accumulation of exact fp16 x fp16 products
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= bound), f"max err {np.abs(got - want).max():.3e} (bound {bound.max():.3e})"


@pytest.mark.parametrize("bits", [4, 8])
def test_dequant_fused_gemm_fold_swiglu_residual(bits):
    """The three fused features the decode step uses: RMSNorm weight folded into the dequantised columns (one fp16 rounding of
    deq32 * fold), SwiGLU over a [gate ; up] packed matrix, residual add."""
    import qwen3tts_b200 as q
    from oracle import mlx_quant

    rng = np.random.default_rng(bits)
    m, inter, in_f = 48, 1536, 1024
    w = (rng.standard_normal((2 * inter, in_f)) * 0.03).astype(np.float32)
    fold = rng.uniform(0.8, 1.2, in_f).astype(np.float32)
    x = rng.standard_normal((m, in_f)).astype(np.float32)
    packed, s, b = mlx_quant.quantize(w, 64, bits, "bf16")
    wd = (mlx_quant.dequantize(packed, s, b, 64, bits, "f32") * fold[None, :]).astype(np.float32).astype(np.float16).astype(np.float64)
    x16 = x.astype(np.float16).astype(np.float64)
    z = x16 @ wd.T
    gate, up = z[:, :inter], z[:, inter:]
    want = gate / (1 + np.exp(-gate)) * up
    got = q.quantized_matmul_tc(x, packed, s, b, 64, bits, "bf16", fold=fold, swiglu_halves=True)
    assert np.abs(got - want).max() <= 2e-5 * max(1.0, np.abs(want).max())
    # plain rows + residual
    res = rng.standard_normal((m, 2 * inter)).astype(np.float32)
    got2 = q.quantized_matmul_tc(x, packed, s, b, 64, bits, "bf16", fold=fold, residual=res)
    assert np.abs(got2 - (z + res)).max() <= 2e-5 * max(1.0, np.abs(z).max())
