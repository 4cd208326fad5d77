"""Third-party anchor for the oracle's vocoder blocks (the reference pins nothing, so independent implementations of the same
blocks are the strongest pins available): HuggingFace `transformers` ships Qwen3-Omni's code2wav, the architecture family the reference's
`Vocoder/SpeechTokenizer.swift` restates -- SnakeBeta (:92-110), causal (dilated / grouped) convolution (:114-170), ConvNeXt block
(:208-236), DecoderResidualUnit (:696-718).  Same weights in, same numbers out.  The transposed convolution is the one block where the two
differ BY DESIGN of the reference (it trims only the right side, :174-204; HF's upstream trims both): the test states that relation."""
import numpy as np
import pytest
import torch

from conftest import ckpt
from oracle import codec as oc

pytest.importorskip("transformers")
hf = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def test_snake_beta():
    C = 24
    m = hf.SnakeBeta(C)
    with torch.no_grad():
        m.alpha.copy_(_rand(C, seed=1, scale=0.5)); m.beta.copy_(_rand(C, seed=2, scale=0.5))
        x = _rand(2, C, 50, seed=3, scale=2.0)
        assert torch.allclose(oc.snake_beta(x, m.alpha, m.beta), m(x), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("k,dil,groups", [(7, 1, 1), (7, 3, 1), (7, 9, 1), (1, 1, 1), (3, 1, 1), (7, 1, 16)])
def test_causal_conv(k, dil, groups):
    cin = cout = 16
    m = hf.Qwen3OmniMoeCausalConvNet(cin, cout, k, dilation=dil, groups=groups)
    with torch.no_grad():
        x = _rand(2, cin, 41, seed=k * 10 + dil)
        want = m(x)
        got = oc.causal_conv1d(x, m.conv.weight, m.conv.bias, dilation=dil, groups=groups)
        assert got.shape == want.shape == x.shape
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_transposed_conv_differs_from_upstream_only_by_the_left_trim():
    """Reference: conv_transpose, drop k - s samples on the RIGHT -> T * s samples (SpeechTokenizer.swift:174-204).  HF's
    Qwen3OmniMoeCausalTransConvNet drops k - s on both sides.  Same transposed convolution underneath: HF's output is the oracle's without its
    first k - s samples."""
    cin, cout, s = 12, 6, 4
    k = 2 * s
    m = hf.Qwen3OmniMoeCausalTransConvNet(cin, cout, k, s)
    with torch.no_grad():
        x = _rand(2, cin, 9, seed=5)
        got = oc.causal_conv_transpose1d(x, m.conv.weight, m.conv.bias, s)
        want = m(x)
        assert got.shape[-1] == 9 * s and want.shape[-1] == 9 * s - (k - s)
        assert torch.allclose(got[..., k - s:], want, rtol=1e-5, atol=1e-6)
    m1 = hf.Qwen3OmniMoeCausalTransConvNet(cin, cout, s, s)  # k == stride (the two upsample stages): nothing to trim on either side
    with torch.no_grad():
        assert torch.allclose(oc.causal_conv_transpose1d(x, m1.conv.weight, m1.conv.bias, s), m1(x), rtol=1e-5, atol=1e-6)


def test_convnext_block():
    d = ckpt("tiny", 8)
    dec = oc.load_codec(d)
    p = "decoder.upsample.0.1"
    C = dec.w[p + ".gamma"].shape[0]
    m = hf.Qwen3OmniMoeConvNeXtBlock(C)
    with torch.no_grad():
        m.dwconv.conv.weight.copy_(dec.w[p + ".dwconv.conv.weight"]); m.dwconv.conv.bias.copy_(dec.w[p + ".dwconv.conv.bias"])
        m.norm.weight.copy_(dec.w[p + ".norm.weight"]); m.norm.bias.copy_(dec.w[p + ".norm.bias"])
        m.pwconv1.weight.copy_(dec.w[p + ".pwconv1.weight"]); m.pwconv1.bias.copy_(dec.w[p + ".pwconv1.bias"])
        m.pwconv2.weight.copy_(dec.w[p + ".pwconv2.weight"]); m.pwconv2.bias.copy_(dec.w[p + ".pwconv2.bias"])
        m.gamma.copy_(dec.w[p + ".gamma"])
        x = _rand(2, C, 33, seed=6)
        got, want = dec.convnext(x, p), m(x)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


@pytest.mark.parametrize("j,dil", [(2, 1), (3, 3), (4, 9)])
def test_decoder_residual_unit(j, dil):
    d = ckpt("tiny", 8)
    dec = oc.load_codec(d)
    p = f"decoder.decoder.1.block.{j}"
    C = dec.w[p + ".act1.alpha"].shape[0]
    m = hf.Qwen3OmniMoeCode2WavDecoderResidualUnit(C, dil)
    with torch.no_grad():
        m.act1.alpha.copy_(dec.w[p + ".act1.alpha"]); m.act1.beta.copy_(dec.w[p + ".act1.beta"])
        m.act2.alpha.copy_(dec.w[p + ".act2.alpha"]); m.act2.beta.copy_(dec.w[p + ".act2.beta"])
        m.conv1.conv.weight.copy_(dec.w[p + ".conv1.conv.weight"]); m.conv1.conv.bias.copy_(dec.w[p + ".conv1.conv.bias"])
        m.conv2.conv.weight.copy_(dec.w[p + ".conv2.conv.weight"]); m.conv2.conv.bias.copy_(dec.w[p + ".conv2.conv.bias"])
        x = _rand(2, C, 57, seed=7 + j)
        u = oc.snake_beta(x, dec.w[p + ".act1.alpha"], dec.w[p + ".act1.beta"])
        u = oc.causal_conv1d(u, dec.w[p + ".conv1.conv.weight"], dec.w[p + ".conv1.conv.bias"], dilation=dil)
        u = oc.snake_beta(u, dec.w[p + ".act2.alpha"], dec.w[p + ".act2.beta"])
        got = oc.causal_conv1d(u, dec.w[p + ".conv2.conv.weight"], dec.w[p + ".conv2.conv.bias"]) + x
        want = m(x)
        assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


def _hf_transformer(dec, sliding_window):
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeCode2WavConfig

    c = dec.c
    cfg = Qwen3OmniMoeCode2WavConfig(hidden_size=c.hidden_size, num_hidden_layers=c.num_hidden_layers, num_attention_heads=c.num_attention_heads,
                                     num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, intermediate_size=c.intermediate_size,
                                     rms_norm_eps=c.rms_norm_eps, rope_theta=c.rope_theta, sliding_window=sliding_window, attention_bias=False,
                                     attn_implementation="eager")
    m = hf.Qwen3OmniMoeCode2WavTransformerModel(cfg).to(torch.float32).eval()
    sd = m.state_dict()
    with torch.no_grad():
        for k in sd:
            assert "decoder.pre_transformer." + k in dec.w, f"no oracle tensor for {k}"
            sd[k].copy_(dec.w["decoder.pre_transformer." + k])
    return m


def _through_hf(dec, m, x):
    p = "decoder.pre_transformer."
    e = x @ dec.w[p + "input_proj.weight"].T + dec.w[p + "input_proj.bias"]
    h = m(inputs_embeds=e).last_hidden_state
    return h @ dec.w[p + "output_proj.weight"].T + dec.w[p + "output_proj.bias"]


@pytest.mark.parametrize("T", [1, 2, 18, 26, 72])
def test_codec_transformer_equals_hf_code2wav_transformer(T):
    """DecoderTransformer (SpeechTokenizer.swift:240-488: RMSNorm eps 1e-5, RoPE theta 1e4, LayerScale, SwiGLU, causal attention) against
    Qwen3OmniMoeCode2WavTransformerModel on the same weights, for windows up to the upstream sliding window of 72 frames."""
    dec = oc.load_codec(ckpt("tiny", 8))
    m = _hf_transformer(dec, 72)
    x = _rand(2, T, dec.w["decoder.pre_transformer.input_proj.weight"].shape[1], seed=T)
    with torch.no_grad():
        got, want = dec.pre_transformer(x), _through_hf(dec, m, x)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


def test_codec_transformer_ignores_the_sliding_window_like_the_reference():
    """Quirk 7 of SURVEY.md §8a: the reference builds a FULL causal mask and never applies `sliding_window` (SpeechTokenizer.swift:469-477).
    Beyond 72 frames the oracle therefore equals upstream with the window switched off, and differs from upstream with its window on."""
    dec = oc.load_codec(ckpt("tiny", 8))
    x = _rand(1, 110, dec.w["decoder.pre_transformer.input_proj.weight"].shape[1], seed=11)
    with torch.no_grad():
        got = dec.pre_transformer(x)
        full = _through_hf(dec, _hf_transformer(dec, 4096), x)
        windowed = _through_hf(dec, _hf_transformer(dec, 72), x)
    assert torch.allclose(got, full, rtol=1e-5, atol=1e-5 * float(full.abs().max()))
    assert not torch.allclose(got[:, 80:], windowed[:, 80:], rtol=1e-3, atol=1e-3 * float(full.abs().max()))
    assert torch.allclose(got[:, :72], windowed[:, :72], rtol=1e-5, atol=1e-5 * float(full.abs().max()))  # the first 72 frames see no window


def test_rvq_decode_equals_hf_mimi_split_quantizer():
    """a16 / a17: codebook = embedding_sum / clip(cluster_usage, 1e-5) (AudioDecoder.swift:285-302) and SplitResidualVectorQuantizer.decode
    (SpeechTokenizer.swift:492-692: gather-sum per chain in codebook order, 1x1 output projection without bias, first + rest) against HF Mimi's
    split residual vector quantiser -- the gather-sums bit for bit, the projected sum to fp32 rounding."""
    from transformers import MimiConfig
    import transformers.models.mimi.modeling_mimi as hm

    dec = oc.load_codec(ckpt("tiny", 8))
    c = dec.c
    D = dec.codebooks[0].shape[1]
    out_dim = dec.w["decoder.quantizer.rvq_first.output_proj.weight"].shape[0]
    n_sem = c.num_semantic_quantizers
    cfg = MimiConfig(hidden_size=out_dim, codebook_size=c.codebook_size, codebook_dim=D, vector_quantization_hidden_dimension=D,
                     num_quantizers=c.num_quantizers, num_semantic_quantizers=n_sem)
    q = hm.MimiSplitResidualVectorQuantizer(cfg).eval()
    with torch.no_grad():
        for name, rvq, n in (("rvq_first", q.semantic_residual_vector_quantizer, n_sem), ("rvq_rest", q.acoustic_residual_vector_quantizer, c.num_quantizers - n_sem)):
            for i in range(n):
                p = f"decoder.quantizer.{name}.vq.layers.{i}._codebook"
                cb = rvq.layers[i].codebook
                cb.embed_sum.copy_(dec.w[p + ".embedding_sum"]); cb.cluster_usage.copy_(dec.w[p + ".cluster_usage"]); cb.initialized.fill_(1.0)
                cb._embed = None
            rvq.output_proj.weight.copy_(dec.w[f"decoder.quantizer.{name}.output_proj.weight"])
        codes = torch.randint(0, c.codebook_size, (2, c.num_quantizers, 13), generator=torch.Generator().manual_seed(3))
        # a dead codebook entry (usage 0 -> clipped to 1e-5) must behave the same on both sides
        first, rest = dec.rvq_embed(codes)
        hf_first = sum(q.semantic_residual_vector_quantizer.layers[i].codebook.decode(codes[:, i]) for i in range(n_sem))
        hf_rest = torch.zeros_like(rest)
        for i in range(c.num_quantizers - n_sem):
            hf_rest = hf_rest + q.acoustic_residual_vector_quantizer.layers[i].codebook.decode(codes[:, n_sem + i])
        assert torch.equal(first, hf_first + torch.zeros_like(first)) and torch.equal(rest, hf_rest)
        want = q.decode(codes)
        got = dec.quantizer_decode(codes)
    assert got.shape == want.shape
    assert torch.allclose(got, want, rtol=1e-6, atol=1e-6 * float(want.abs().max()))
