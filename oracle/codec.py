"""Speech-tokenizer DECODER restated on torch-CPU fp32 (oracle; test infrastructure only; parity unpinned).

Follows `Vocoder/SpeechTokenizer.swift:92-988` (graph) and `Vocoder/AudioDecoder.swift:196-305` (checkpoint
sanitising), reading the on-disk PyTorch layouts directly:

    conv weight            disk [C_out, C_in/g, K]  (reference permutes to MLX [C_out, K, C_in/g], :276-277)
    transposed-conv weight disk [C_in, C_out, K]    (reference permutes to MLX [C_out, K, C_in],   :271-275)

Tensors stay NCL here so `torch.nn.functional.conv1d / conv_transpose1d` apply unchanged.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def snake_beta(x, alpha, beta):
    """`SnakeBeta` / `DecoderOutputSnake`: x + sin^2(x e^alpha) / (e^beta + 1e-9)  (SpeechTokenizer.swift:105-109, 816-820)."""
    a = torch.exp(alpha)[None, :, None]
    b = torch.exp(beta)[None, :, None]
    return x + (1.0 / (b + 1e-9)) * torch.sin(x * a).pow(2)


def causal_conv1d(x, w, b, dilation=1, groups=1):
    """`CausalConv1d` with stride 1: left-pad (K-1)*d zeros, no right pad (SpeechTokenizer.swift:114-170;
    `getExtraPadding` is 0 for stride 1, :154-158).  x [B,C,T], w [C_out, C_in/g, K]."""
    k = w.shape[-1]
    pad = (k - 1) * dilation
    return F.conv1d(F.pad(x, (pad, 0)), w, b, dilation=dilation, groups=groups)


def causal_conv_transpose1d(x, w, b, stride):
    """`CausalTransposeConv1d` / `DecoderBlockUpsample`: conv_transpose, then trim right K - stride
    (SpeechTokenizer.swift:174-204, 720-751).  x [B,C_in,T], w [C_in, C_out, K] -> [B, C_out, T*stride]."""
    k = w.shape[-1]
    y = F.conv_transpose1d(x, w, b, stride=stride)
    trim = k - stride
    if trim > 0 and y.shape[-1] - trim > 0:
        y = y[..., : y.shape[-1] - trim]
    return y


def rms_norm(x, w, eps):
    """`DecoderRMSNorm` (SpeechTokenizer.swift:240-256)."""
    v = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(v + eps))


def rotate_half(x):
    h = x.shape[-1] // 2
    return torch.cat([-x[..., h:], x[..., :h]], dim=-1)


class CodecDecoder:
    """`Qwen3TTSSpeechTokenizerDecoder` (SpeechTokenizer.swift:844-988)."""

    def __init__(self, cfg, tensors: dict):
        self.c = cfg
        self.w = {k: v.to(torch.float32) for k, v in tensors.items()}
        self.total_upsample = math.prod(list(cfg.upsample_rates) + list(cfg.upsampling_ratios))
        # codebook = embedding_sum / clip(cluster_usage, 1e-5)[:, None]   (AudioDecoder.swift:285-302)
        self.codebooks = []
        n_sem = cfg.num_semantic_quantizers
        for name, n in (("rvq_first", n_sem), ("rvq_rest", cfg.num_quantizers - n_sem)):
            for i in range(n):
                p = f"decoder.quantizer.{name}.vq.layers.{i}._codebook"
                usage = self.w[p + ".cluster_usage"].clamp(min=1e-5)
                self.codebooks.append(self.w[p + ".embedding_sum"] / usage[:, None])

    # --- a17: SplitResidualVectorQuantizer.decode (SpeechTokenizer.swift:684-691, 629-639, 566-582)
    def rvq_embed(self, codes):
        """codes [B,Q,T] -> (first [B,T,D], rest [B,T,D]) fp32 gather-sums in codebook order (bit-exact contract)."""
        n_sem = self.c.num_semantic_quantizers
        B, Q, T = codes.shape
        D = self.codebooks[0].shape[1]
        first = torch.zeros(B, T, D)
        for q in range(n_sem):
            first = first + self.codebooks[q][codes[:, q].long()]
        rest = torch.zeros(B, T, D)
        for q in range(n_sem, Q):
            rest = rest + self.codebooks[q][codes[:, q].long()]
        return first, rest

    def quantizer_decode(self, codes):
        first, rest = self.rvq_embed(codes)
        w1 = self.w["decoder.quantizer.rvq_first.output_proj.weight"][:, :, 0]
        w2 = self.w["decoder.quantizer.rvq_rest.output_proj.weight"][:, :, 0]
        out = first @ w1.T
        if codes.shape[1] > self.c.num_semantic_quantizers:
            out = out + rest @ w2.T
        return out.transpose(1, 2)  # [B, codebook_dim, T]

    # --- a18: DecoderTransformer (SpeechTokenizer.swift:439-488)
    def pre_transformer(self, x):
        c, w = self.c, self.w
        p = "decoder.pre_transformer"
        B, T, _ = x.shape
        x = x @ w[p + ".input_proj.weight"].T + w[p + ".input_proj.bias"]
        hd = c.head_dim
        inv = 1.0 / torch.pow(torch.tensor(c.rope_theta, dtype=torch.float32),
                              torch.arange(0, hd, 2, dtype=torch.float32) / hd)
        fr = torch.arange(T, dtype=torch.float32)[:, None] * inv[None, :]
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos()[None, None], emb.sin()[None, None]
        mask = None
        if T > 1:  # full causal mask; `sliding_window` is never applied (:474-477)
            mask = torch.triu(torch.full((T, T), -1e9), diagonal=1)
        nh, nkv = c.num_attention_heads, c.num_key_value_heads
        for i in range(c.num_hidden_layers):
            lp = f"{p}.layers.{i}"
            h = rms_norm(x, w[lp + ".input_layernorm.weight"], c.rms_norm_eps)

            def proj(name, n):
                y = h @ w[f"{lp}.self_attn.{name}.weight"].T
                bk = f"{lp}.self_attn.{name}.bias"
                if bk in w:
                    y = y + w[bk]
                return y.view(B, T, n, hd).transpose(1, 2)

            q, k, v = proj("q_proj", nh), proj("k_proj", nkv), proj("v_proj", nkv)
            q = q * cos + rotate_half(q) * sin
            k = k * cos + rotate_half(k) * sin
            if nkv != nh:
                k = k.repeat_interleave(nh // nkv, dim=1)
                v = v.repeat_interleave(nh // nkv, dim=1)
            s = (q @ k.transpose(-1, -2)) * (hd ** -0.5)
            if mask is not None:
                s = s + mask
            a = torch.softmax(s, -1) @ v
            a = a.transpose(1, 2).reshape(B, T, nh * hd) @ w[lp + ".self_attn.o_proj.weight"].T
            if lp + ".self_attn.o_proj.bias" in w:
                a = a + w[lp + ".self_attn.o_proj.bias"]
            x = x + w[lp + ".self_attn_layer_scale.scale"] * a
            h = rms_norm(x, w[lp + ".post_attention_layernorm.weight"], c.rms_norm_eps)
            m = (F.silu(h @ w[lp + ".mlp.gate_proj.weight"].T) * (h @ w[lp + ".mlp.up_proj.weight"].T)) @ w[lp + ".mlp.down_proj.weight"].T
            x = x + w[lp + ".mlp_layer_scale.scale"] * m
        x = rms_norm(x, w[p + ".norm.weight"], c.rms_norm_eps)
        return x @ w[p + ".output_proj.weight"].T + w[p + ".output_proj.bias"]

    # --- a21: ConvNeXtBlock (SpeechTokenizer.swift:208-236)
    def convnext(self, x, p):
        w = self.w
        C = x.shape[1]
        h = causal_conv1d(x, w[p + ".dwconv.conv.weight"], w[p + ".dwconv.conv.bias"], groups=C)
        h = h.transpose(1, 2)
        h = F.layer_norm(h, (C,), w[p + ".norm.weight"], w[p + ".norm.bias"], eps=1e-6)
        h = h @ w[p + ".pwconv1.weight"].T + w[p + ".pwconv1.bias"]
        h = F.gelu(h)  # exact erf form
        h = h @ w[p + ".pwconv2.weight"].T + w[p + ".pwconv2.bias"]
        h = w[p + ".gamma"] * h
        return x + h.transpose(1, 2)

    # --- a23: decodeImpl (SpeechTokenizer.swift:917-952)
    def decode(self, codes, clip=True, taps: dict | None = None):
        """codes int [B, Q, T] -> wav [B, 1, T*1920]."""
        c, w = self.c, self.w
        if codes.shape[1] != c.num_quantizers:
            return torch.zeros(codes.shape[0], 1, 0)
        h = self.quantizer_decode(codes)
        if taps is not None:
            taps["rvq"] = h
        h = causal_conv1d(h, w["decoder.pre_conv.conv.weight"], w["decoder.pre_conv.conv.bias"])
        if taps is not None:
            taps["pre_conv"] = h
        h = self.pre_transformer(h.transpose(1, 2)).transpose(1, 2)
        if taps is not None:
            taps["pre_transformer"] = h
        for i, f in enumerate(c.upsampling_ratios):
            h = causal_conv_transpose1d(h, w[f"decoder.upsample.{i}.0.conv.weight"], w[f"decoder.upsample.{i}.0.conv.bias"], f)
            h = self.convnext(h, f"decoder.upsample.{i}.1")
            if taps is not None:
                taps[f"upsample{i}"] = h
        h = causal_conv1d(h, w["decoder.decoder.0.conv.weight"], w["decoder.decoder.0.conv.bias"])
        if taps is not None:
            taps["init_conv"] = h
        for i, s in enumerate(c.upsample_rates):
            p = f"decoder.decoder.{i + 1}.block"
            h = snake_beta(h, w[p + ".0.alpha"], w[p + ".0.beta"])
            h = causal_conv_transpose1d(h, w[p + ".1.conv.weight"], w[p + ".1.conv.bias"], s)
            for j, d in ((2, 1), (3, 3), (4, 9)):
                r = h
                u = snake_beta(h, w[f"{p}.{j}.act1.alpha"], w[f"{p}.{j}.act1.beta"])
                u = causal_conv1d(u, w[f"{p}.{j}.conv1.conv.weight"], w[f"{p}.{j}.conv1.conv.bias"], dilation=d)
                u = snake_beta(u, w[f"{p}.{j}.act2.alpha"], w[f"{p}.{j}.act2.beta"])
                u = causal_conv1d(u, w[f"{p}.{j}.conv2.conv.weight"], w[f"{p}.{j}.conv2.conv.bias"])
                h = u + r
            if taps is not None:
                taps[f"block{i}"] = h
        n = len(c.upsample_rates) + 1
        h = snake_beta(h, w[f"decoder.decoder.{n}.alpha"], w[f"decoder.decoder.{n}.beta"])
        h = causal_conv1d(h, w[f"decoder.decoder.{n + 1}.conv.weight"], w[f"decoder.decoder.{n + 1}.conv.bias"])
        return h.clamp(-1.0, 1.0) if clip else h

    # --- a23: chunkedDecode (SpeechTokenizer.swift:954-987)
    def chunked_decode(self, codes, chunk_size=100, left_context=10):
        """Left-pads with code id 0 (quirk 8), right-pads to a multiple of `chunk_size`, stacks chunks on the
        batch axis (chunk-major), decodes, drops the context samples and re-interleaves."""
        B, Q, T = codes.shape
        n_chunks = (T + chunk_size - 1) // chunk_size
        right = n_chunks * chunk_size - T
        padded = F.pad(codes, (left_context, right))
        chunks = [padded[:, :, i * chunk_size: i * chunk_size + chunk_size + left_context] for i in range(n_chunks)]
        out = self.decode(torch.cat(chunks, 0))
        valid = out[:, :, left_context * self.total_upsample:]
        target = T * self.total_upsample
        if B == 1:
            return valid.reshape(1, 1, -1)[:, :, :target]
        v = valid.reshape(n_chunks, B, 1, valid.shape[2]).permute(1, 2, 0, 3)
        return v.reshape(B, 1, -1)[:, :, :target]


def load_codec(model_dir: str) -> CodecDecoder:
    """Read `<model_dir>/speech_tokenizer/{config.json, model.safetensors}` (Qwen3TTSPipeline.swift:191-208)."""
    import json
    import os

    from safetensors.torch import load_file

    from .checkpoint import CodecDims

    d = os.path.join(model_dir, "speech_tokenizer")
    cj = json.load(open(os.path.join(d, "config.json"))).get("decoder_config", {})
    cfg = CodecDims()
    for k, v in cj.items():
        if hasattr(cfg, k):
            setattr(cfg, k, v)
    t = load_file(os.path.join(d, "model.safetensors"))
    t = {(k[len("audio_decoder."):] if k.startswith("audio_decoder.") else k): v for k, v in t.items()}
    return CodecDecoder(cfg, t)
