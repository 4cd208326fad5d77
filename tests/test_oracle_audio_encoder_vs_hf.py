"""Third-party anchor for the ICL reference-audio encoder restatement (oracle/audio_encoder.py): the reference's
`Qwen3TTSAudioEncoder` (Vocoder/Qwen3TTSAudioEncoder.swift) is a Mimi encoder, and HuggingFace `transformers` ships Mimi.  Same weights in:
the SEANet CNN, the stride-2 downsampling convolution and the split residual vector quantiser's nearest-neighbour search must agree.
Where the REFERENCE deviates from upstream Mimi the oracle follows the reference, and the tests say so: the downsampling conv pads with
zeros (MimiConv1d's default `constant` mode, :340-358; upstream Mimi builds it with `replicate`), and the encoder transformer is
bidirectional (:331; upstream's is causal with a 250-frame window) -- so the transformer is anchored elsewhere
(tests/test_oracle_audio_encoder.py)."""
import numpy as np
import pytest
import torch

from conftest import ckpt
from oracle import audio_encoder as ae

pytest.importorskip("transformers")
hm = pytest.importorskip("transformers.models.mimi.modeling_mimi")


@pytest.fixture(scope="module")
def pair():
    from transformers import MimiConfig

    orc = ae.AudioEncoderOracle(ckpt("tiny", 8, encoder="tiny") + "/speech_tokenizer")
    c = orc.cfg
    cfg = MimiConfig(audio_channels=c.audio_channels, hidden_size=c.hidden_size, num_filters=c.num_filters, num_residual_layers=c.num_residual_layers,
                     upsampling_ratios=list(c.upsampling_ratios), kernel_size=c.kernel_size, last_kernel_size=c.last_kernel_size,
                     residual_kernel_size=c.residual_kernel_size, dilation_growth_rate=c.dilation_growth_rate, use_causal_conv=True,
                     compress=c.compress, codebook_size=c.codebook_size, codebook_dim=c.codebook_dim, num_quantizers=c.num_quantizers,
                     vector_quantization_hidden_dimension=c.vector_quantization_hidden_dimension, num_semantic_quantizers=c.num_semantic_quantizers,
                     num_hidden_layers=c.num_hidden_layers, intermediate_size=c.intermediate_size, num_attention_heads=c.num_attention_heads,
                     num_key_value_heads=c.num_key_value_heads, head_dim=c.head_dim, norm_eps=c.norm_eps)
    return orc, cfg


@pytest.mark.parametrize("L", [960, 1921, 2437, 24000])
def test_seanet_equals_hf_mimi_encoder(pair, L):
    orc, cfg = pair
    enc = hm.MimiEncoder(cfg).to(torch.float32).eval()
    sd = enc.state_dict()
    with torch.no_grad():
        for k in sd:
            assert "encoder." + k in orc.w, f"no oracle tensor for {k}"
            sd[k].copy_(orc.w["encoder." + k])
        x = torch.randn(2, 1, L, generator=torch.Generator().manual_seed(L)) * 0.1
        want, got = enc(x), orc.seanet(x)
    assert got.shape == want.shape and got.shape[2] == -(-L // 960)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-5 * float(want.abs().max()))


@pytest.mark.parametrize("T", [1, 2, 7, 26])
def test_downsample_equals_hf_mimi_conv_with_zero_padding(pair, T):
    orc, cfg = pair
    c = orc.cfg
    has_bias = "downsample.conv.conv.bias" in orc.w
    ds = hm.MimiConv1d(cfg, c.hidden_size, c.hidden_size, kernel_size=2 * c.compress, stride=c.compress, bias=has_bias, pad_mode="constant").eval()
    with torch.no_grad():
        ds.conv.weight.copy_(orc.w["downsample.conv.conv.weight"])
        if has_bias:
            ds.conv.bias.copy_(orc.w["downsample.conv.conv.bias"])
        x = torch.randn(2, c.hidden_size, T, generator=torch.Generator().manual_seed(T))
        want, got = ds(x), orc._conv("downsample.conv.conv", x, stride=c.compress)
    assert got.shape == want.shape and got.shape[2] == -(-T // c.compress)
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-6)


def test_split_rvq_search_equals_hf_mimi_quantizer(pair):
    """EuclideanCodebook.encode (SpeechTokenizer.swift:511-519) and the residual chains of EncoderSplitResidualVectorQuantizer (:424-460):
    identical code ids for all 16 output layers."""
    orc, cfg = pair
    c = orc.cfg
    q = hm.MimiSplitResidualVectorQuantizer(cfg).eval()
    sd = q.state_dict()
    with torch.no_grad():
        for k in sd:
            if k.endswith("initialized"):
                sd[k].fill_(1.0)
                continue
            if k.endswith("output_proj.weight"):
                continue  # decode side: not part of the encoder path, not in the checkpoint
            src = "quantizer." + k.replace(".codebook.embed_sum", "._codebook.embedding_sum").replace(".codebook.cluster_usage", "._codebook.cluster_usage")
            assert src in orc.w, f"no oracle tensor for {k}"
            sd[k].copy_(orc.w[src].reshape(sd[k].shape))
        lat = torch.randn(3, c.hidden_size, 25, generator=torch.Generator().manual_seed(4))
        want = q.encode(lat, num_quantizers=16)  # [Q, B, T]
        n_sem = c.num_semantic_quantizers
        mine = torch.stack(orc._rvq("semantic", lat.transpose(1, 2), n_sem) + orc._rvq("acoustic", lat.transpose(1, 2), 16 - n_sem), 0)
    assert mine.shape == want.shape
    assert torch.equal(mine.long(), want.long())
