"""CPU: codec-decoder restatement (oracle/codec.py) against independent formulations (no upstream golden vectors exist)."""
import numpy as np
import pytest
import torch

from conftest import ckpt
from oracle import codec as ocodec, pipeline as opipe


def test_causal_conv_is_left_padded_only():
    torch.manual_seed(0)
    x = torch.randn(2, 5, 20)
    w = torch.randn(7, 5, 3)
    b = torch.randn(7)
    for d in (1, 3):
        y = ocodec.causal_conv1d(x, w, b, dilation=d)
        assert y.shape == (2, 7, 20)
        # direct definition: y[t] = b + sum_k w[:, :, k] x[t - (K-1-k)*d]
        ref = torch.zeros_like(y)
        for t in range(20):
            acc = b.clone()[None].repeat(2, 1)
            for k in range(3):
                ts = t - (2 - k) * d
                if ts >= 0:
                    acc = acc + x[:, :, ts] @ w[:, :, k].T
            ref[:, :, t] = acc
        assert torch.allclose(y, ref, atol=1e-5)
        # causality: changing the future does not change the past
        x2 = x.clone()
        x2[:, :, 12:] += 1.0
        assert torch.allclose(ocodec.causal_conv1d(x2, w, b, dilation=d)[:, :, :12], y[:, :, :12])


@pytest.mark.parametrize("stride,k", [(2, 2), (3, 6), (5, 10), (8, 16)])
def test_transposed_conv_polyphase_identity(stride, k):
    """Output sample t*s + j = x[t] . w[:,:,j] (+ x[t-1] . w[:,:,j+s] when k = 2s) — the form the CUDA engine uses."""
    torch.manual_seed(1)
    x = torch.randn(2, 4, 9)
    w = torch.randn(4, 6, k)
    b = torch.randn(6)
    y = ocodec.causal_conv_transpose1d(x, w, b, stride)
    assert y.shape == (2, 6, 9 * stride)
    ref = torch.zeros_like(y)
    for t in range(9):
        for j in range(stride):
            acc = b[None] + x[:, :, t] @ w[:, :, j]
            if k == 2 * stride and t > 0:
                acc = acc + x[:, :, t - 1] @ w[:, :, j + stride]
            ref[:, :, t * stride + j] = acc
    assert torch.allclose(y, ref, atol=1e-5)


def test_snake_beta():
    x = torch.linspace(-3, 3, 50).reshape(1, 2, 25)
    a, b = torch.tensor([0.3, -0.2]), torch.tensor([0.1, 0.4])
    y = ocodec.snake_beta(x, a, b)
    ref = x + torch.sin(x * torch.exp(a)[None, :, None]) ** 2 / (torch.exp(b)[None, :, None] + 1e-9)
    assert torch.allclose(y, ref)
    assert torch.allclose(ocodec.snake_beta(x, torch.zeros(2), torch.zeros(2)), x + torch.sin(x) ** 2 / (1 + 1e-9))


@pytest.fixture(scope="module")
def codec():
    return ocodec.load_codec(ckpt("tiny", 8))


def test_codebook_is_embedding_sum_over_clipped_usage(codec):
    w = codec.w
    p = "decoder.quantizer.rvq_rest.vq.layers.3._codebook"
    want = w[p + ".embedding_sum"] / w[p + ".cluster_usage"].clamp(min=1e-5)[:, None]
    assert torch.equal(codec.codebooks[1 + 3], want)
    assert (w[p + ".cluster_usage"] == 0).any()  # the synthetic checkpoint exercises the clip


def test_rvq_embed_is_plain_indexing(codec):
    codes = torch.randint(0, 2048, (2, 16, 5), dtype=torch.int32)
    first, rest = codec.rvq_embed(codes)
    assert torch.equal(first, codec.codebooks[0][codes[:, 0].long()])
    acc = torch.zeros_like(rest)
    for q in range(1, 16):
        acc = acc + codec.codebooks[q][codes[:, q].long()]
    assert torch.equal(rest, acc)


def test_transformer_attention_matches_sdpa(codec):
    """The hand-rolled softmax(QK^T + causal mask)V equals torch SDPA with is_causal=True."""
    torch.manual_seed(2)
    c = codec.c
    x = torch.randn(2, 9, c.latent_dim)
    y = codec.pre_transformer(x)
    assert y.shape == x.shape and torch.isfinite(y).all()
    # causality of the whole transformer
    x2 = x.clone()
    x2[:, 6:] += 1.0
    assert torch.allclose(codec.pre_transformer(x2)[:, :6], y[:, :6], atol=1e-5)
    # T == 1: no mask path
    assert torch.isfinite(codec.pre_transformer(x[:, :1])).all()


def test_decode_shapes_clip_and_causality(codec):
    codes = torch.randint(0, 2048, (2, 16, 6), dtype=torch.int32)
    wav = codec.decode(codes)
    assert wav.shape == (2, 1, 6 * 1920) and wav.abs().max() <= 1.0
    # the whole decoder is causal in frames: a later code cannot change earlier samples
    codes2 = codes.clone()
    codes2[:, :, 4:] = (codes2[:, :, 4:] + 7) % 2048
    assert torch.allclose(codec.decode(codes2)[:, :, : 4 * 1920], wav[:, :, : 4 * 1920], atol=1e-5)
    assert codec.decode(codes[:, :15]).shape == (2, 1, 0)  # wrong quantizer count -> empty (SpeechTokenizer.swift:918-920)
    rms = float(codec.decode(codes, clip=False).pow(2).mean().sqrt())
    assert 0.02 < rms < 0.6, "synthetic codec output should sit inside the clip range"


def test_chunked_decode_equals_manual_windows(codec):
    B, T, chunk, left = 2, 23, 10, 3
    codes = torch.randint(0, 2048, (B, 16, T), dtype=torch.int32)
    got = codec.chunked_decode(codes, chunk, left)
    assert got.shape == (B, 1, T * 1920)
    for b in range(B):
        parts = []
        for ci in range((T + chunk - 1) // chunk):
            win = torch.zeros(1, 16, chunk + left, dtype=torch.int32)  # left context of the first chunk = code 0 (quirk 8)
            for p in range(chunk + left):
                src = ci * chunk + p - left
                if 0 <= src < T:
                    win[0, :, p] = codes[b, :, src]
            parts.append(codec.decode(win)[0, 0, left * 1920:])
        want = torch.cat(parts)[: T * 1920]
        assert torch.allclose(got[b, 0], want, atol=2e-5)  # batched vs single-window conv: fp32 summation order only


def test_windowed_schedules(codec):
    frames = torch.randint(0, 2048, (41, 16)).tolist()
    for chunk in (16, 18, 24):
        wins = opipe.decode_windowed(codec, frames, chunk, 8)
        assert [r for _, r in wins] == [(p, min(p + chunk, 41)) for p in range(0, 41, chunk)]
        assert sum(w.size for w, _ in wins) == 41 * 1920
    # stream consumer: 12-frame code chunks -> 18-frame decode windows, flush, trailing empty final (quirk 9)
    out = opipe.stream_chunks(codec, [frames[i:i + 12] for i in range(0, 41, 12)])
    assert [(c["token_range"], c["is_final"]) for c in out] == [((0, 18), False), ((18, 36), False), ((36, 41), True), ((41, 41), True)]
    assert out[-1]["samples"].size == 0
    # frames with code0 outside [0, 2048) are dropped before decoding (Qwen3TTSPipeline.swift:576-579)
    bad = [list(f) for f in frames[:20]]
    bad[3][0] = 2148
    bad[7][0] = 3000
    out = opipe.stream_chunks(codec, [bad])
    assert out[0]["token_range"] == (0, 18) and out[-1]["token_range"] == (18, 18)


def test_crossfade():
    a, b = np.ones(1000, np.float32), np.zeros(1000, np.float32)
    y = opipe.crossfade_concat([a, b], 480)
    assert y.size == 1000 + 1000 - 480
    assert y[519] == 1.0 and y[520] == 1.0 and abs(y[520 + 240] - 0.5) < 1e-6 and y[-1] == 0.0
