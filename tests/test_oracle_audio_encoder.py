"""CPU self-checks of the ICL reference-audio encoder restatement (oracle/audio_encoder.py) against independent formulations:
MimiConv1d padding arithmetic vs brute force, the polyphase identity the CUDA path relies on for strided convs, nearest-neighbour
search vs explicit distances, frame-count formula."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ckpt
from oracle import audio_encoder as ae


@pytest.mark.parametrize("k,stride,dil", [(7, 1, 1), (3, 1, 1), (3, 1, 2), (8, 4, 1), (10, 5, 1), (12, 6, 1), (16, 8, 1), (4, 2, 1)])
@pytest.mark.parametrize("L", [1, 5, 16, 37, 960, 1001])
def test_mimi_conv_padding_and_length(k, stride, dil, L):
    """paddingLeft = (k-1)d + 1 - stride, extra right padding completes the last frame (Qwen3TTSAudioEncoder.swift:42-62): the output has
    ceil(L / stride) frames and frame t only sees samples <= t*stride + stride - 1 (causal)."""
    g = torch.Generator().manual_seed(k * 100 + L)
    x = torch.randn(1, 3, L, generator=g)
    w = torch.randn(5, 3, k, generator=g)
    y = ae.mimi_conv1d(x, w, None, stride, dil)
    assert y.shape[2] == -(-L // stride)
    # causality: changing samples after frame t's last input leaves frames <= t unchanged
    t = y.shape[2] // 2
    cut = min(L, t * stride + stride)
    x2 = x.clone()
    x2[:, :, cut:] += 1.0
    y2 = ae.mimi_conv1d(x2, w, None, stride, dil)
    assert torch.equal(y[:, :, : t + 1], y2[:, :, : t + 1])


@pytest.mark.parametrize("r,L", [(4, 37), (5, 40), (8, 9), (2, 7)])
def test_strided_conv_is_a_two_tap_conv_on_the_folded_input(r, L):
    """The CUDA path runs every stride-r, kernel-2r causal conv as a 2-tap causal conv over the [ceil(L / r), r * C] view of the
    zero-padded input (csrc/audio_encoder.cu load_conv): same numbers."""
    g = torch.Generator().manual_seed(r * 10 + L)
    C, O = 3, 4
    x = torch.randn(1, C, L, generator=g, dtype=torch.float64)
    w = torch.randn(O, C, 2 * r, generator=g, dtype=torch.float64)
    want = ae.mimi_conv1d(x, w, None, r, 1)
    To = -(-L // r)
    xp = F.pad(x, (0, To * r - L))[0].T.reshape(To, r * C)  # channels-last rows folded r at a time: index j * C + ci
    w0 = w[:, :, :r].permute(0, 2, 1).reshape(O, r * C)      # tap 0: kernel indices [0, r) against the previous folded row
    w1 = w[:, :, r:].permute(0, 2, 1).reshape(O, r * C)      # tap 1: [r, 2r) against the current row
    prev = torch.cat([torch.zeros(1, r * C, dtype=torch.float64), xp[:-1]], 0)
    got = (prev @ w0.T + xp @ w1.T).T[None]
    assert torch.allclose(got, want, atol=1e-12)


def test_encoder_oracle_end_to_end_shapes_and_search():
    d = ckpt("tiny", 8, encoder="tiny")
    orc = ae.AudioEncoderOracle(d + "/speech_tokenizer")
    rng = np.random.default_rng(0)
    for L in (960, 1921, 24000 + 517):
        a = (rng.standard_normal(L) * 0.1).astype(np.float32)
        rec = {}
        codes = orc.encode(a, rec)
        T = -(-(-(-L // 960)) // 2)
        assert codes.shape == (1, 16, T) and codes.dtype == np.int32
        assert codes.min() >= 0 and codes.max() < orc.cfg.codebook_size
        assert rec["latent"].shape == (1, T, orc.cfg.hidden_size)
    # the search is a true arg-min of explicit squared distances (first quantiser of each chain)
    lat = torch.from_numpy(rec["latent"])[0]
    for name, row in (("semantic", 0), ("acoustic", orc.cfg.num_semantic_quantizers)):
        proj = lat @ orc.w[f"quantizer.{name}_residual_vector_quantizer.input_proj.weight"][:, :, 0].T
        d2 = ((proj[:, None, :] - orc.codebooks[name][0][None]) ** 2).sum(-1)
        near = rec["margins"][0, row] > 1e-3
        assert np.array_equal(torch.argmin(d2, -1).numpy()[near], codes[0, row][near])
    assert len(np.unique(codes)) > 40  # the synthetic latent spreads over the codebooks


def test_encoder_checkpoint_key_scheme():
    """Keys follow the module tree `Qwen3TTSAudioEncoder.sanitizeEncoderWeights` unflattens (Qwen3TTSAudioEncoder.swift:589-648)."""
    from safetensors.torch import load_file

    d = ckpt("tiny", 8, encoder="tiny")
    w = load_file(d + "/speech_tokenizer/model.safetensors")
    for k in ("encoder.encoder.layers.0.conv.weight", "encoder.encoder.layers.1.block.1.conv.weight", "encoder.encoder.layers.3.conv.weight",
              "encoder.encoder.layers.14.conv.bias", "encoder.encoder_transformer.layers.0.self_attn.q_proj.weight",
              "encoder.encoder_transformer.layers.1.mlp.fc2.bias", "encoder.encoder_transformer.layers.0.self_attn_layer_scale.scale",
              "encoder.downsample.conv.conv.weight", "encoder.quantizer.semantic_residual_vector_quantizer.input_proj.weight",
              "encoder.quantizer.acoustic_residual_vector_quantizer.layers.30._codebook.embedding_sum",
              "encoder.quantizer.acoustic_residual_vector_quantizer.layers.0._codebook.cluster_usage"):
        assert k in w, k
    assert w["encoder.encoder.layers.3.conv.weight"].shape == (16, 8, 8)  # [out, in, 2 * ratio], first ratio = 4 (reversed list)
    assert any(k.startswith("decoder.") for k in w)  # the decoder's tensors share the file
