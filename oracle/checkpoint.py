"""Synthetic Qwen3-TTS checkpoints in the REFERENCE's on-disk format (oracle; test infrastructure only).

There is no network, so every test / bench runs on seeded random-init weights of the named
architecture.  This writer emits exactly the files `Qwen3TTSPipeline.init` reads
(`Qwen3TTSPipeline.swift:118-232`):

    <dir>/config.json                       (`Model/Qwen3Config.swift:208-253`: nested `talker_config`,
                                             top-level `quantization`, `tts_*_token_id`)
    <dir>/model.safetensors                 (key scheme = what `Qwen3Talker.load` strips/remaps,
                                             `Model/Qwen3Talker.swift:114-270`)
    <dir>/speech_tokenizer/config.json      (`Vocoder/AudioDecoder.swift:7-102`)
    <dir>/speech_tokenizer/model.safetensors (PyTorch conv layouts, permuted at load by
                                             `AudioDecoder.sanitize`, `Vocoder/AudioDecoder.swift:196-305`)

Quantised leaves carry `weight` (uint32, MLX affine packing), `scales`, `biases`.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field, asdict

import numpy as np
import torch
from safetensors.torch import save_file

from . import mlx_quant

SPK_ID = {"serena": 3066, "vivian": 3065, "uncle_fu": 3010, "ryan": 3061, "aiden": 2861,
          "ono_anna": 2873, "sohee": 2864, "eric": 2875, "dylan": 2878}  # Qwen3Config.swift:126


@dataclass
class CodePredictorDims:  # Qwen3Config.swift:21-33
    hidden_size: int = 1024
    num_hidden_layers: int = 5
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 3072
    rms_norm_eps: float = 1e-6
    max_position_embeddings: int = 65536
    rope_theta: float = 1000000.0
    vocab_size: int = 2048
    num_code_groups: int = 16


@dataclass
class TalkerDims:  # Qwen3Config.swift:104-128 (`.standard`)
    hidden_size: int = 1024
    num_hidden_layers: int = 28
    vocab_size: int = 3072
    text_vocab_size: int = 151936
    text_hidden_size: int = 2048
    num_attention_heads: int = 16
    num_key_value_heads: int = 8
    head_dim: int = 128
    intermediate_size: int = 3072
    rms_norm_eps: float = 1e-6
    max_position_embeddings: int = 32768
    rope_theta: float = 1000000.0
    tts_bos_token_id: int = 151672
    tts_eos_token_id: int = 151673
    tts_pad_token_id: int = 151671
    codec_bos_id: int = 2149
    codec_eos_token_id: int = 2150
    codec_pad_id: int = 2148
    codec_nothink_id: int = 2155
    codec_think_bos_id: int = 2156
    codec_think_eos_id: int = 2157
    code_predictor: CodePredictorDims = field(default_factory=CodePredictorDims)
    mrope_section: list | None = None
    tts_model_type: str | None = None


@dataclass
class CodecDims:  # SpeechTokenizer.swift:42-74
    latent_dim: int = 1024
    codebook_dim: int = 512
    codebook_size: int = 2048
    decoder_dim: int = 1536
    hidden_size: int = 512
    intermediate_size: int = 1024
    layer_scale_initial_scale: float = 0.01
    max_position_embeddings: int = 8000
    head_dim: int = 64
    num_attention_heads: int = 16
    num_hidden_layers: int = 8
    num_key_value_heads: int = 16
    num_quantizers: int = 16
    num_semantic_quantizers: int = 1
    rms_norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    sliding_window: int = 72
    upsample_rates: list = field(default_factory=lambda: [8, 5, 4, 3])
    upsampling_ratios: list = field(default_factory=lambda: [2, 2])
    attention_bias: bool = False


def preset(name: str) -> tuple[TalkerDims, CodecDims]:
    """'0.6b' = `Qwen3TTSConfig.standard`; '1.7b' = assumed upstream dims (H 2048 / MLP 6144,
    same heads/layers, 1024-wide code predictor => `small_to_mtp_projection`); 'tiny' = CPU-test size."""
    if name == "0.6b":
        return TalkerDims(), CodecDims()
    if name == "1.7b":
        return TalkerDims(hidden_size=2048, intermediate_size=6144), CodecDims()
    if name == "1.7b-cv":  # BASELINE.json configs[4]: the 1.7B CustomVoice variant (config.json `tts_model_type`, Qwen3Config.swift:230-233)
        return TalkerDims(hidden_size=2048, intermediate_size=6144, tts_model_type="custom_voice"), CodecDims()
    if name == "tiny":
        t = TalkerDims(hidden_size=256, num_hidden_layers=2, text_vocab_size=640, text_hidden_size=128,
                       num_attention_heads=4, num_key_value_heads=2, head_dim=128, intermediate_size=512,
                       tts_bos_token_id=601, tts_eos_token_id=602, tts_pad_token_id=600,
                       code_predictor=CodePredictorDims(hidden_size=128, num_hidden_layers=2, num_attention_heads=2,
                                                        num_key_value_heads=1, head_dim=128, intermediate_size=256))
        c = CodecDims(latent_dim=128, codebook_dim=64, decoder_dim=192, hidden_size=128, intermediate_size=256,
                      num_attention_heads=2, num_key_value_heads=2, num_hidden_layers=2)
        return t, c
    if name == "tcsmall":  # tiny talker + a codec whose every contraction fits the tcgen05 path (cin % 8 == 0, N % 32 == 0)
        t, _ = preset("tiny")
        c = CodecDims(latent_dim=128, codebook_dim=64, decoder_dim=512, hidden_size=128, intermediate_size=256,
                      num_attention_heads=2, num_key_value_heads=2, num_hidden_layers=2)
        return t, c
    if name == "codecfull":  # tiny talker + the full-size codec decoder (SpeechTokenizer.swift:42-74 defaults)
        t, _ = preset("tiny")
        return t, CodecDims()
    if name == "tiny-cp":  # code predictor as wide as the talker (like the 0.6B model): no small_to_mtp_projection
        t, c = preset("tiny")
        t.code_predictor = CodePredictorDims(hidden_size=256, num_hidden_layers=2, num_attention_heads=2, num_key_value_heads=1, head_dim=128,
                                             intermediate_size=256)
        return t, c
    if name == "tiny-mrope":
        t, c = preset("tiny")
        t.mrope_section = [24, 20, 20]
        return t, c
    raise ValueError(name)


_TDT = {"bf16": torch.bfloat16, "f16": torch.float16, "f32": torch.float32}


class _Rng:
    def __init__(self, seed: int):
        self.g = torch.Generator().manual_seed(seed)

    def normal(self, shape, std):
        return torch.randn(shape, generator=self.g, dtype=torch.float32) * std

    def uniform(self, shape, lo, hi):
        return torch.rand(shape, generator=self.g, dtype=torch.float32) * (hi - lo) + lo


def _u32(a: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).view(torch.uint32)


def _put_linear(out: dict, key: str, w: torch.Tensor, bias: torch.Tensor | None, bits: int, group: int, dt: str):
    """Write one `QuantizedLayerFactory.linear` leaf (QuantizedLayerFactory.swift:49-66)."""
    if bits:
        packed, s, b = mlx_quant.quantize(w.numpy(), group, bits, dt)
        out[key + ".weight"] = _u32(packed)
        out[key + ".scales"] = torch.from_numpy(s).to(_TDT[dt])
        out[key + ".biases"] = torch.from_numpy(b).to(_TDT[dt])
    else:
        out[key + ".weight"] = w.to(_TDT[dt])
    if bias is not None:
        out[key + ".bias"] = bias.to(_TDT[dt])


def talker_tensors(t: TalkerDims, bits: int, dtype: str, seed: int, group: int = 64, head_std: float = 0.25,
                   init: str = "stress") -> dict:
    """Seeded random-init talker + code-predictor weights under the reference's checkpoint keys.

    init = "stress" (default of the unit tests): linears N(0, 0.02^2); `codec_head` / `lm_head` N(0, head_std^2) so greedy
    margins are far above fp noise (logit rms ~8 at full size); embeddings N(0, 0.5^2); norm weights U(0.8, 1.2) so they are
    numerically visible.
    init = "baseline" (BASELINE.md §3 / SURVEY.md §8d config 1, the init the north-star tolerances are quoted on): EVERY
    matrix, head, embedding and bias N(0, 0.02^2), norm weights = 1 (logit rms ~0.6 at full size).
    """
    assert init in ("stress", "baseline")
    base = init == "baseline"
    r = _Rng(seed)
    o: dict = {}
    dt = _TDT[dtype]
    H, cp = t.hidden_size, t.code_predictor
    emb_std = 0.02 if base else 0.5
    if base:
        head_std = 0.02

    def norm_w(n):
        return (torch.ones(n) if base else r.uniform((n,), 0.8, 1.2)).to(dt)

    o["talker.model.text_embedding.weight"] = r.normal((t.text_vocab_size, t.text_hidden_size), emb_std).to(dt)
    o["talker.model.codec_embedding.weight"] = r.normal((t.vocab_size, H), emb_std).to(dt)
    _put_linear(o, "talker.text_projection.linear_fc1", r.normal((t.text_hidden_size, t.text_hidden_size), 0.02),
                r.normal((t.text_hidden_size,), 0.02), bits, group, dtype)
    _put_linear(o, "talker.text_projection.linear_fc2", r.normal((H, t.text_hidden_size), 0.02 if base else 0.04),
                r.normal((H,), 0.02), bits, group, dtype)

    def layer(prefix, hid, nh, nkv, hd, inter):
        o[prefix + ".input_layernorm.weight"] = norm_w(hid)
        o[prefix + ".post_attention_layernorm.weight"] = norm_w(hid)
        _put_linear(o, prefix + ".self_attn.q_proj", r.normal((nh * hd, hid), 0.02), None, bits, group, dtype)
        _put_linear(o, prefix + ".self_attn.k_proj", r.normal((nkv * hd, hid), 0.02), None, bits, group, dtype)
        _put_linear(o, prefix + ".self_attn.v_proj", r.normal((nkv * hd, hid), 0.02), None, bits, group, dtype)
        _put_linear(o, prefix + ".self_attn.o_proj", r.normal((hid, nh * hd), 0.02), None, bits, group, dtype)
        o[prefix + ".self_attn.q_norm.weight"] = norm_w(hd)
        o[prefix + ".self_attn.k_norm.weight"] = norm_w(hd)
        _put_linear(o, prefix + ".mlp.gate_proj", r.normal((inter, hid), 0.02), None, bits, group, dtype)
        _put_linear(o, prefix + ".mlp.up_proj", r.normal((inter, hid), 0.02), None, bits, group, dtype)
        _put_linear(o, prefix + ".mlp.down_proj", r.normal((hid, inter), 0.02), None, bits, group, dtype)

    for i in range(t.num_hidden_layers):
        layer(f"talker.model.layers.{i}", H, t.num_attention_heads, t.num_key_value_heads, t.head_dim, t.intermediate_size)
    o["talker.model.norm.weight"] = norm_w(H)
    _put_linear(o, "talker.codec_head", r.normal((t.vocab_size, H), head_std), None, bits, group, dtype)

    for i in range(cp.num_code_groups - 1):
        o[f"talker.code_predictor.model.codec_embedding.{i}.weight"] = r.normal((cp.vocab_size, H), emb_std).to(dt)
    for i in range(cp.num_hidden_layers):
        layer(f"talker.code_predictor.model.layers.{i}", cp.hidden_size, cp.num_attention_heads,
              cp.num_key_value_heads, cp.head_dim, cp.intermediate_size)
    o["talker.code_predictor.model.norm.weight"] = norm_w(cp.hidden_size)
    for i in range(cp.num_code_groups - 1):
        _put_linear(o, f"talker.code_predictor.lm_head.{i}", r.normal((cp.vocab_size, cp.hidden_size), head_std),
                    None, bits, group, dtype)
    if cp.hidden_size != H:  # Qwen3CodePredictor.swift:171-175
        _put_linear(o, "talker.code_predictor.small_to_mtp_projection", r.normal((cp.hidden_size, H), 0.02 if base else 0.03),
                    r.normal((cp.hidden_size,), 0.02), bits, group, dtype)
    return o


def codec_tensors(c: CodecDims, seed: int, gain: float = 0.6, visible: bool = True) -> dict:
    """Seeded codec-decoder weights (fp32, PyTorch layouts as on disk).

    convs/linears N(0, gain^2 / fan_in); `visible=True` draws SnakeBeta alpha/beta, LayerScale and ConvNeXt
    gamma from O(1) ranges so every branch contributes (the reference's inits — 0, 0.01, 1e-6,
    SpeechTokenizer.swift:100-101, 220, 264 — would hide the transformer/ConvNeXt branches); `visible=False`
    uses the reference's init values.
    """
    r = _Rng(seed)
    o: dict = {}

    def conv(key, cout, cin_per_group, k, bias=True, g=gain):
        o[key + ".weight"] = r.normal((cout, cin_per_group, k), g / np.sqrt(cin_per_group * k))
        if bias:
            o[key + ".bias"] = r.normal((cout,), 0.02)

    def convT(key, cin, cout, k, stride):
        # each output sample sums k/stride taps of cin channels
        o[key + ".weight"] = r.normal((cin, cout, k), gain / np.sqrt(cin * max(1, k // stride)))
        o[key + ".bias"] = r.normal((cout,), 0.02)

    def lin(key, cout, cin, bias=True, g=gain):
        o[key + ".weight"] = r.normal((cout, cin), g / np.sqrt(cin))
        if bias:
            o[key + ".bias"] = r.normal((cout,), 0.02)

    def snake(key, ch):
        if visible:
            o[key + ".alpha"] = r.normal((ch,), 0.3)
            o[key + ".beta"] = r.normal((ch,), 0.3)
        else:
            o[key + ".alpha"] = torch.zeros(ch)
            o[key + ".beta"] = torch.zeros(ch)

    vq_dim = c.codebook_dim // 2
    n_rest = c.num_quantizers - c.num_semantic_quantizers
    for name, n in (("rvq_first", c.num_semantic_quantizers), ("rvq_rest", n_rest)):
        p = f"decoder.quantizer.{name}"
        for i in range(n):
            usage = r.uniform((c.codebook_size,), 0.5, 50.0)
            emb = r.normal((c.codebook_size, vq_dim), 1.0)
            dead = torch.arange(c.codebook_size) % 97 == 13  # exercise clip(cluster_usage, 1e-5) (AudioDecoder.swift:285-302)
            usage[dead] = 0.0
            esum = emb * usage[:, None]
            esum[dead] = r.normal((int(dead.sum()), vq_dim), 1e-6)
            o[f"{p}.vq.layers.{i}._codebook.embedding_sum"] = esum
            o[f"{p}.vq.layers.{i}._codebook.cluster_usage"] = usage
        o[f"{p}.input_proj.weight"] = r.normal((vq_dim, c.codebook_dim, 1), 1 / np.sqrt(c.codebook_dim))
        o[f"{p}.output_proj.weight"] = r.normal((c.codebook_dim, vq_dim, 1), 1 / np.sqrt(vq_dim * c.num_quantizers))
    conv("decoder.pre_conv.conv", c.latent_dim, c.codebook_dim, 3, g=1.0)
    pt = "decoder.pre_transformer"
    lin(pt + ".input_proj", c.hidden_size, c.latent_dim, g=1.0)
    lin(pt + ".output_proj", c.latent_dim, c.hidden_size, g=1.0)
    o[pt + ".norm.weight"] = r.uniform((c.hidden_size,), 0.8, 1.2)
    qd = c.num_attention_heads * c.head_dim
    kvd = c.num_key_value_heads * c.head_dim
    for i in range(c.num_hidden_layers):
        lp = f"{pt}.layers.{i}"
        lin(lp + ".self_attn.q_proj", qd, c.hidden_size, bias=c.attention_bias, g=1.0)
        lin(lp + ".self_attn.k_proj", kvd, c.hidden_size, bias=c.attention_bias, g=1.0)
        lin(lp + ".self_attn.v_proj", kvd, c.hidden_size, bias=c.attention_bias, g=1.0)
        lin(lp + ".self_attn.o_proj", c.hidden_size, qd, bias=c.attention_bias)
        lin(lp + ".mlp.gate_proj", c.intermediate_size, c.hidden_size, bias=False, g=1.0)
        lin(lp + ".mlp.up_proj", c.intermediate_size, c.hidden_size, bias=False, g=1.0)
        lin(lp + ".mlp.down_proj", c.hidden_size, c.intermediate_size, bias=False)
        o[lp + ".input_layernorm.weight"] = r.uniform((c.hidden_size,), 0.8, 1.2)
        o[lp + ".post_attention_layernorm.weight"] = r.uniform((c.hidden_size,), 0.8, 1.2)
        ls = r.uniform((c.hidden_size,), 0.3, 0.7) if visible else torch.full((c.hidden_size,), c.layer_scale_initial_scale)
        o[lp + ".self_attn_layer_scale.scale"] = ls
        ls2 = r.uniform((c.hidden_size,), 0.3, 0.7) if visible else torch.full((c.hidden_size,), c.layer_scale_initial_scale)
        o[lp + ".mlp_layer_scale.scale"] = ls2
    for i, f in enumerate(c.upsampling_ratios):
        convT(f"decoder.upsample.{i}.0.conv", c.latent_dim, c.latent_dim, f, f)
        conv(f"decoder.upsample.{i}.1.dwconv.conv", c.latent_dim, 1, 7, g=1.0)
        o[f"decoder.upsample.{i}.1.norm.weight"] = r.uniform((c.latent_dim,), 0.8, 1.2)
        o[f"decoder.upsample.{i}.1.norm.bias"] = r.normal((c.latent_dim,), 0.05)
        lin(f"decoder.upsample.{i}.1.pwconv1", 4 * c.latent_dim, c.latent_dim, g=1.0)
        lin(f"decoder.upsample.{i}.1.pwconv2", c.latent_dim, 4 * c.latent_dim)
        o[f"decoder.upsample.{i}.1.gamma"] = r.uniform((c.latent_dim,), 0.3, 0.7) if visible else torch.full((c.latent_dim,), 1e-6)
    conv("decoder.decoder.0.conv", c.decoder_dim, c.latent_dim, 7, g=1.0)
    for i, s in enumerate(c.upsample_rates):
        cin = c.decoder_dim // (2 ** i)
        cout = c.decoder_dim // (2 ** (i + 1))
        p = f"decoder.decoder.{i + 1}.block"
        snake(p + ".0", cin)
        convT(p + ".1.conv", cin, cout, 2 * s, s)
        for j in (2, 3, 4):
            snake(f"{p}.{j}.act1", cout)
            conv(f"{p}.{j}.conv1.conv", cout, cout, 7)
            snake(f"{p}.{j}.act2", cout)
            conv(f"{p}.{j}.conv2.conv", cout, cout, 1)
    n_out = len(c.upsample_rates) + 1
    cl = c.decoder_dim // (2 ** len(c.upsample_rates))
    snake(f"decoder.decoder.{n_out}", cl)
    conv(f"decoder.decoder.{n_out + 1}.conv", 1, cl, 7, g=1.0)
    return o


def encoder_preset(name: str):
    """ICL reference-audio encoder dimensions: 'full' = Qwen3TTSTokenizerEncoderConfig defaults (SpeechTokenizer.swift:9-40); 'tiny' = CPU-test size."""
    from .audio_encoder import EncoderDims

    if name == "full":
        return EncoderDims()
    if name == "tiny":
        return EncoderDims(codebook_dim=32, codebook_size=256, hidden_size=128, intermediate_size=256, num_filters=8, num_hidden_layers=2,
                           num_quantizers=32, num_attention_heads=2, num_key_value_heads=2, vector_quantization_hidden_dimension=32)
    raise ValueError(name)


def encoder_tensors(e, seed: int) -> dict:
    """Seeded ICL-encoder weights under the `encoder.*` keys `Qwen3TTSAudioEncoder.sanitizeEncoderWeights` reads
    (Vocoder/Qwen3TTSAudioEncoder.swift:589-648; module tree :120-460), PyTorch layouts (conv [out, in, k])."""
    r = _Rng(seed)
    o: dict = {}

    def conv(key, cout, cin, k, g=1.0):
        o[key + ".weight"] = r.normal((cout, cin, k), g / np.sqrt(cin * k))
        o[key + ".bias"] = r.normal((cout,), 0.02)

    li = 0
    conv(f"encoder.encoder.layers.{li}.conv", e.num_filters, e.audio_channels, e.kernel_size)
    li += 1
    cur = e.num_filters
    for i, ratio in enumerate(reversed(e.upsampling_ratios)):
        for _ in range(e.num_residual_layers):
            conv(f"encoder.encoder.layers.{li}.block.1.conv", cur // 2, cur, 3)
            conv(f"encoder.encoder.layers.{li}.block.3.conv", cur, cur // 2, 1, g=0.5)
            li += 1
        li += 1  # ELU
        conv(f"encoder.encoder.layers.{li}.conv", e.num_filters * 2 ** (i + 1), cur, 2 * ratio)
        cur = e.num_filters * 2 ** (i + 1)
        li += 1
    li += 1  # ELU
    conv(f"encoder.encoder.layers.{li}.conv", e.hidden_size, cur, e.last_kernel_size)
    H = e.hidden_size
    for n in range(e.num_hidden_layers):
        p = f"encoder.encoder_transformer.layers.{n}"
        for ln in ("input_layernorm", "post_attention_layernorm"):
            o[f"{p}.{ln}.weight"] = r.uniform((H,), 0.8, 1.2)
            o[f"{p}.{ln}.bias"] = r.normal((H,), 0.05)
        for nm, od in (("q_proj", e.num_attention_heads * e.head_dim), ("k_proj", e.num_key_value_heads * e.head_dim),
                       ("v_proj", e.num_key_value_heads * e.head_dim)):
            o[f"{p}.self_attn.{nm}.weight"] = r.normal((od, H), 1.0 / np.sqrt(H))
        o[f"{p}.self_attn.o_proj.weight"] = r.normal((H, e.num_attention_heads * e.head_dim), 0.6 / np.sqrt(e.num_attention_heads * e.head_dim))
        o[f"{p}.mlp.fc1.weight"] = r.normal((e.intermediate_size, H), 1.0 / np.sqrt(H))
        o[f"{p}.mlp.fc1.bias"] = r.normal((e.intermediate_size,), 0.02)
        o[f"{p}.mlp.fc2.weight"] = r.normal((H, e.intermediate_size), 0.6 / np.sqrt(e.intermediate_size))
        o[f"{p}.mlp.fc2.bias"] = r.normal((H,), 0.02)
        o[f"{p}.self_attn_layer_scale.scale"] = r.uniform((H,), 0.3, 0.7)
        o[f"{p}.mlp_layer_scale.scale"] = r.uniform((H,), 0.3, 0.7)
    conv("encoder.downsample.conv.conv", H, H, 2 * e.compress)
    D = e.vector_quantization_hidden_dimension
    for name, n in (("semantic", e.num_semantic_quantizers), ("acoustic", e.num_quantizers - e.num_semantic_quantizers)):
        p = f"encoder.quantizer.{name}_residual_vector_quantizer"
        o[p + ".input_proj.weight"] = r.normal((D, H, 1), 1.0 / np.sqrt(H))
        o[p + ".output_proj.weight"] = r.normal((H, D, 1), 1.0 / np.sqrt(D))
        for i in range(n):
            usage = r.uniform((e.codebook_size,), 0.5, 50.0)
            emb = r.normal((e.codebook_size, D), 1.0 / (1.0 + 0.35 * i))  # residual codebooks shrink like a trained RVQ's
            dead = torch.arange(e.codebook_size) % 89 == 7               # clip(cluster_usage, 1e-5) path
            usage[dead] = 0.0
            esum = emb * usage[:, None]
            esum[dead] = r.normal((int(dead.sum()), D), 1e-6)
            o[f"{p}.layers.{i}._codebook.embedding_sum"] = esum
            o[f"{p}.layers.{i}._codebook.cluster_usage"] = usage
    return o


def talker_config_json(t: TalkerDims, bits: int, group: int = 64) -> dict:
    tc = {k: v for k, v in asdict(t).items()
          if k not in ("code_predictor", "mrope_section", "tts_model_type", "tts_bos_token_id", "tts_eos_token_id", "tts_pad_token_id")}
    tc["spk_id"] = dict(SPK_ID)
    tc["code_predictor_config"] = asdict(t.code_predictor)
    if t.mrope_section is not None:
        tc["rope_scaling"] = {"mrope_section": t.mrope_section, "interleaved": True}
    cfg = {"talker_config": tc, "tts_bos_token_id": t.tts_bos_token_id, "tts_eos_token_id": t.tts_eos_token_id,
           "tts_pad_token_id": t.tts_pad_token_id}
    if t.tts_model_type:
        cfg["tts_model_type"] = t.tts_model_type
    if bits:
        cfg["quantization"] = {"group_size": group, "bits": bits}
    return cfg


def codec_config_json(c: CodecDims) -> dict:
    return {"decoder_config": asdict(c), "input_sample_rate": 24000, "output_sample_rate": 24000,
            "decode_upsample_rate": int(np.prod(c.upsample_rates + c.upsampling_ratios)),
            "encoder_valid_num_quantizers": 16}


def write_checkpoint(path: str, name: str = "tiny", bits: int = 8, dtype: str = "bf16", seed: int = 0,
                     codec_seed: int | None = None, visible: bool = True, with_codec: bool = True,
                     normalize_codec: bool = True, init: str = "stress", quant_key: str = "quantization", encoder: str | None = None,
                     speaker_encoder: str | None = None) -> str:
    """Write a complete synthetic model directory; returns `path`.  Idempotent via a stamp file.

    init: see `talker_tensors`.  quant_key = "quantization_config" writes the packed leaves WITHOUT a top-level
    `quantization` block, i.e. the checkpoint form `Qwen3Talker.load` dequantises offline to fp16 (Qwen3Talker.swift:139-175)."""
    t, c = preset(name)
    stamp = {"name": name, "bits": bits, "dtype": dtype, "seed": seed, "codec_seed": codec_seed, "visible": visible,
             "with_codec": with_codec, "normalize_codec": normalize_codec, "init": init, "quant_key": quant_key, "encoder": encoder, "speaker_encoder": speaker_encoder, "v": 7}
    stamp_path = os.path.join(path, "synthetic_stamp.json")
    if os.path.exists(stamp_path):
        try:
            if json.load(open(stamp_path)) == stamp:
                return path
        except Exception:
            pass
    os.makedirs(os.path.join(path, "speech_tokenizer"), exist_ok=True)
    cfg = talker_config_json(t, bits)
    if bits and quant_key != "quantization":
        cfg[quant_key] = cfg.pop("quantization")
    json.dump(cfg, open(os.path.join(path, "config.json"), "w"), indent=1)
    tt = talker_tensors(t, bits, dtype, seed, init=init)
    if speaker_encoder:  # ECAPA-TDNN weights live in the talker file under `speaker_encoder.*` (SpeakerEncoder.swift:550-555)
        tt.update(speaker_encoder_tensors(speaker_encoder, seed + 3000))
    save_file(tt, os.path.join(path, "model.safetensors"))
    if with_codec:
        cj = codec_config_json(c)
        ct = codec_tensors(c, seed + 1000 if codec_seed is None else codec_seed, visible=visible)
        if normalize_codec:
            _normalize_codec_output(c, ct)
        if encoder:  # ICL reference-audio encoder weights live in the same file under `encoder.*` (Qwen3TTSAudioEncoder.swift:587-603)
            ed = encoder_preset(encoder)
            cj["encoder_config"] = asdict(ed)
            et = encoder_tensors(ed, seed + 2000)
            _normalize_encoder_latent(ed, et)
            ct.update(et)
        json.dump(cj, open(os.path.join(path, "speech_tokenizer", "config.json"), "w"), indent=1)
        save_file({k: v.contiguous() for k, v in ct.items()}, os.path.join(path, "speech_tokenizer", "model.safetensors"))
    json.dump(stamp, open(stamp_path, "w"))
    return path


def speaker_encoder_tensors(preset_name: str, seed: int) -> dict:
    """Seeded ECAPA-TDNN weights under the `speaker_encoder.*` keys `SpeakerEncoder.load` reads (SpeakerEncoder.swift:550-603), PyTorch conv
    layout [out, in, k], fp32.  "full" = the reference's only configuration (:399-418); "tiny" = the same graph at small widths."""
    ch, se, att, enc, mel = {"full": (512, 128, 128, 1024, 128), "tiny": (64, 16, 16, 256, 128)}[preset_name]
    mfa_ch = 3 * ch
    r = _Rng(seed)
    o: dict = {}

    def conv(key, cout, cin, k, g=1.0):
        o[f"speaker_encoder.{key}.weight"] = r.normal((cout, cin, k), g / np.sqrt(cin * k))
        o[f"speaker_encoder.{key}.bias"] = r.normal((cout,), 0.05)

    kernels = (5, 3, 3, 3, 1)
    conv("blocks.0.conv", ch, mel, kernels[0], g=0.3)  # log-mel inputs are O(5): keep the first activations O(1)
    for i in (1, 2, 3):
        p = f"blocks.{i}"
        conv(p + ".tdnn1.conv", ch, ch, 1, g=1.4)
        for j in range(7):
            conv(f"{p}.res2net_block.blocks.{j}.conv", ch // 8, ch // 8, kernels[i], g=1.4)
        conv(p + ".tdnn2.conv", ch, ch, 1, g=1.4)
        conv(p + ".se_block.conv1", se, ch, 1)
        conv(p + ".se_block.conv2", ch, se, 1)
    conv("mfa.conv", mfa_ch, mfa_ch, kernels[4], g=1.4)
    conv("asp.tdnn.conv", att, 3 * mfa_ch, 1)
    conv("asp.conv", mfa_ch, att, 1, g=2.0)
    conv("fc", enc, 2 * mfa_ch, 1)
    return o


def _normalize_encoder_latent(e, et: dict, seconds: float = 1.0):
    """Rescale the quantizers' input projections so the projected latent has unit RMS on noise input: with the codebooks drawn at
    O(1) scale the nearest-neighbour search then spreads over the codebook instead of collapsing onto its smallest entries."""
    import tempfile

    from .audio_encoder import AudioEncoderOracle

    with tempfile.TemporaryDirectory() as td:
        json.dump({"encoder_config": asdict(e)}, open(os.path.join(td, "config.json"), "w"))
        save_file({k: v.contiguous() for k, v in et.items()}, os.path.join(td, "model.safetensors"))
        orc = AudioEncoderOracle(td)
    g = torch.Generator().manual_seed(4321)
    audio = (torch.randn(int(24000 * seconds), generator=g) * 0.1).numpy()
    rec: dict = {}
    orc.encode(audio, rec)
    lat = torch.from_numpy(rec["latent"])[0]  # [T, hidden]
    for name in ("semantic", "acoustic"):
        k = f"encoder.quantizer.{name}_residual_vector_quantizer.input_proj.weight"
        proj = lat @ et[k][:, :, 0].T
        et[k] = et[k] / max(float(proj.pow(2).mean().sqrt()), 1e-12)


def _normalize_codec_output(c: CodecDims, ct: dict, target_rms: float = 0.2, frames: int = 6):
    """Rescale the final conv so PCM is not saturated by `clip(-1, 1)` (SpeechTokenizer.swift:951) —
    otherwise SNR on clipped output would say nothing."""
    from .codec import CodecDecoder

    dec = CodecDecoder(c, ct)
    g = torch.Generator().manual_seed(1234)
    codes = torch.randint(0, c.codebook_size, (1, c.num_quantizers, frames), generator=g, dtype=torch.int32)
    wav = dec.decode(codes, clip=False)
    rms = float(wav.pow(2).mean().sqrt())
    n_out = len(c.upsample_rates) + 2
    k = f"decoder.decoder.{n_out}.conv"
    s = target_rms / max(rms, 1e-12)
    ct[k + ".weight"] = ct[k + ".weight"] * s
    ct[k + ".bias"] = ct[k + ".bias"] * s
