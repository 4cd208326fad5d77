"""ICL reference-audio encoder: CPU restatement of `Qwen3TTSAudioEncoder` (oracle; test infrastructure only; parity unpinned).

Follows `Vocoder/Qwen3TTSAudioEncoder.swift` line by line (torch CPU fp32):
  * `MimiConv1d` :24-85        causal conv: all padding left ((k-1)*d + 1 - stride), extra right padding so the last frame is whole
  * `MimiResnetBlock` :89-116  x + conv1(ELU(conv3(ELU(x))))
  * `MimiSEANetEncoder` :120-190   1->64 k7; for ratio in reversed(upsampling_ratios): resnet, ELU, conv k=2r stride r (channels x2); ELU, conv -> hidden k3
  * `EncoderAttention` :194-245, `EncoderMLP` :249-263, `EncoderTransformerLayer` :267-305, `EncoderTransformer` :309-335
        pre-LayerNorm, bidirectional MHA (no mask), RoPE theta 1e4, exact-erf GELU MLP with biases, LayerScale on both branches
  * `EncoderDownsample` :339-358   conv k = 2*compress, stride compress
  * `EncoderResidualVectorQuantizer` :380-420 / `EuclideanCodebook.encode` (`SpeechTokenizer.swift:511-519`)
        1x1 input projection (no bias), residual nearest-neighbour search: argmin(|x|^2 - 2 x.e + |e|^2), first index on ties
  * `EncoderSplitResidualVectorQuantizer` :424-460, `Qwen3TTSAudioEncoder.callAsFunction` :530-572 (keep the first 16 quantizers)
  * `sanitizeEncoderWeights` :589-648: `encoder.*` keys only, codebook = embedding_sum / clip(cluster_usage, 1e-5)

Checkpoint keys (speech_tokenizer/model.safetensors, PyTorch layouts): see `encoder_tensors` in oracle/checkpoint.py.
"""
from __future__ import annotations

import json
import math
import os
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F
from safetensors.torch import load_file


@dataclass
class EncoderDims:  # Qwen3TTSTokenizerEncoderConfig (SpeechTokenizer.swift:9-40)
    audio_channels: int = 1
    codebook_dim: int = 256
    codebook_size: int = 2048
    compress: int = 2
    dilation_growth_rate: int = 2
    hidden_size: int = 512
    intermediate_size: int = 2048
    kernel_size: int = 7
    last_kernel_size: int = 3
    num_filters: int = 64
    num_hidden_layers: int = 8
    num_residual_layers: int = 1
    num_quantizers: int = 32
    num_semantic_quantizers: int = 1
    residual_kernel_size: int = 3
    upsampling_ratios: list = field(default_factory=lambda: [8, 6, 5, 4])
    head_dim: int = 64
    num_attention_heads: int = 8
    num_key_value_heads: int = 8
    norm_eps: float = 1e-5
    rope_theta: float = 10000.0
    max_position_embeddings: int = 8000
    layer_scale_initial_scale: float = 0.01
    vector_quantization_hidden_dimension: int = 256


def mimi_extra_padding(length: int, kernel: int, stride: int, dilation: int) -> tuple[int, int]:
    """(paddingLeft, extraPadding) of `MimiConv1d` (:42-62), with its Float arithmetic."""
    eff = (kernel - 1) * dilation + 1
    pad_left = eff - stride
    n_frames = np.float32(length - eff + pad_left) / np.float32(stride) + np.float32(1)
    ideal = (int(math.ceil(float(n_frames))) - 1) * stride + (eff - pad_left)
    return pad_left, max(0, ideal - length)


def mimi_conv1d(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor | None, stride: int = 1, dilation: int = 1) -> torch.Tensor:
    """x [B, C, T], w [out, in, k] (PyTorch layout) -> [B, out, T']."""
    pl, extra = mimi_extra_padding(x.shape[2], w.shape[2], stride, dilation)
    return F.conv1d(F.pad(x, (pl, extra)), w, b, stride=stride, dilation=dilation)


def elu(x: torch.Tensor) -> torch.Tensor:
    """`ELUActivation` (:8-20): max(x, 0) + min(alpha * (exp(x) - 1), 0), alpha = 1."""
    return torch.clamp(x, min=0) + torch.clamp(torch.exp(x) - 1, max=0)


class AudioEncoderOracle:
    def __init__(self, tokenizer_dir: str):
        cfg_path = None
        for name in ("config.json", "configuration.json", "speech_tokenizer_config.json"):
            p = os.path.join(tokenizer_dir, name)
            if os.path.exists(p):
                cfg_path = p
                break
        raw = json.load(open(cfg_path)) if cfg_path else {}
        self.cfg = EncoderDims()
        for k, v in (raw.get("encoder_config") or {}).items():
            if hasattr(self.cfg, k):
                setattr(self.cfg, k, v)
        self.valid_q = int(raw.get("encoder_valid_num_quantizers") or 16)
        allw = load_file(os.path.join(tokenizer_dir, "model.safetensors"))
        self.w = {k[len("encoder."):]: v.to(torch.float32) for k, v in allw.items() if k.startswith("encoder.")}  # :600-603
        c = self.cfg
        # codebooks: embedding_sum / clip(cluster_usage, 1e-5) (:627-645)
        self.codebooks = {}
        for name, n in (("semantic", c.num_semantic_quantizers), ("acoustic", c.num_quantizers - c.num_semantic_quantizers)):
            books = []
            for i in range(n):
                p = f"quantizer.{name}_residual_vector_quantizer.layers.{i}._codebook"
                usage = torch.clamp(self.w[p + ".cluster_usage"], min=1e-5)
                books.append(self.w[p + ".embedding_sum"] / usage[:, None])
            self.codebooks[name] = books
        self.inv_freq = 1.0 / torch.pow(torch.tensor(c.rope_theta, dtype=torch.float32),
                                        torch.arange(0, c.head_dim, 2, dtype=torch.float32) / torch.tensor(float(c.head_dim)))

    # ---- SEANet CNN (:120-190)
    def _conv(self, key, x, stride=1, dilation=1):
        return mimi_conv1d(x, self.w[key + ".weight"], self.w.get(key + ".bias"), stride, dilation)

    def seanet(self, x: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        li = 0
        h = self._conv(f"encoder.layers.{li}.conv", x)
        li += 1
        for ratio in reversed(c.upsampling_ratios):
            for j in range(c.num_residual_layers):
                d = c.dilation_growth_rate ** j
                r = elu(h)
                r = self._conv(f"encoder.layers.{li}.block.1.conv", r, dilation=d)
                r = elu(r)
                r = self._conv(f"encoder.layers.{li}.block.3.conv", r)
                h = h + r
                li += 1
            h = elu(h)
            li += 1  # the ELU occupies a slot of `layers`
            h = self._conv(f"encoder.layers.{li}.conv", h, stride=ratio)
            li += 1
        h = elu(h)
        li += 1
        return self._conv(f"encoder.layers.{li}.conv", h)

    # ---- transformer (:194-335)
    def transformer(self, x: torch.Tensor) -> torch.Tensor:
        c = self.cfg
        B, T, H = x.shape
        pos = torch.arange(T, dtype=torch.float32)
        fr = pos[:, None] * self.inv_freq[None, :]
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos()[None, None], emb.sin()[None, None]

        def rot(t):
            h = t.shape[-1] // 2
            return torch.cat([-t[..., h:], t[..., :h]], -1)

        h = x
        for n in range(c.num_hidden_layers):
            p = f"encoder_transformer.layers.{n}"
            r = h
            y = F.layer_norm(h, (H,), self.w[p + ".input_layernorm.weight"], self.w[p + ".input_layernorm.bias"], c.norm_eps)
            q = (y @ self.w[p + ".self_attn.q_proj.weight"].T).view(B, T, c.num_attention_heads, c.head_dim).transpose(1, 2)
            k = (y @ self.w[p + ".self_attn.k_proj.weight"].T).view(B, T, c.num_key_value_heads, c.head_dim).transpose(1, 2)
            v = (y @ self.w[p + ".self_attn.v_proj.weight"].T).view(B, T, c.num_key_value_heads, c.head_dim).transpose(1, 2)
            q = q * cos + rot(q) * sin
            k = k * cos + rot(k) * sin
            if c.num_key_value_heads != c.num_attention_heads:
                g = c.num_attention_heads // c.num_key_value_heads
                k, v = k.repeat_interleave(g, 1), v.repeat_interleave(g, 1)
            att = torch.softmax((q @ k.transpose(-1, -2)) * (c.head_dim ** -0.5), -1) @ v  # bidirectional: no mask (:331)
            att = att.transpose(1, 2).reshape(B, T, -1) @ self.w[p + ".self_attn.o_proj.weight"].T
            h = r + self.w[p + ".self_attn_layer_scale.scale"] * att
            r = h
            y = F.layer_norm(h, (H,), self.w[p + ".post_attention_layernorm.weight"], self.w[p + ".post_attention_layernorm.bias"], c.norm_eps)
            y = F.gelu(y @ self.w[p + ".mlp.fc1.weight"].T + self.w[p + ".mlp.fc1.bias"])  # exact erf GELU
            y = y @ self.w[p + ".mlp.fc2.weight"].T + self.w[p + ".mlp.fc2.bias"]
            h = r + self.w[p + ".mlp_layer_scale.scale"] * y
        return h

    # ---- quantizer (:380-460)
    def _rvq(self, name: str, x_btc: torch.Tensor, n_layers: int, record=None):
        p = f"quantizer.{name}_residual_vector_quantizer"
        proj = x_btc @ self.w[p + ".input_proj.weight"][:, :, 0].T  # Conv1d k=1, no bias
        res = proj
        codes = []
        for i in range(n_layers):
            e = self.codebooks[name][i]
            x_sq = (res * res).sum(-1, keepdim=True)
            e_sq = (e * e).sum(-1)
            dist = x_sq - 2 * (res @ e.T) + e_sq
            idx = torch.argmin(dist, dim=-1)  # first minimal index
            if record is not None:
                top2 = torch.topk(dist, 2, dim=-1, largest=False).values
                record.append((top2[..., 1] - top2[..., 0]).numpy())
            codes.append(idx.to(torch.int32))
            res = res - e[idx]
        return codes

    def encode(self, audio, record: dict | None = None) -> np.ndarray:
        """audio float32 [L] or [B, L] -> codes int32 [B, valid_q, T] (`Qwen3TTSAudioEncoder.callAsFunction` :530-572)."""
        c = self.cfg
        x = torch.as_tensor(np.asarray(audio, dtype=np.float32))
        if x.ndim == 1:
            x = x[None]
        h = self.seanet(x[:, None, :])                      # [B, hidden, L/960]
        h = self.transformer(h.transpose(1, 2))             # [B, T, hidden]
        h = self._conv("downsample.conv.conv", h.transpose(1, 2), stride=c.compress)  # [B, hidden, T/2]
        if record is not None:
            record["latent"] = h.transpose(1, 2).numpy().copy()
        hb = h.transpose(1, 2)
        n_sem = c.num_semantic_quantizers
        n_ac = min(c.num_quantizers - n_sem, max(0, self.valid_q - n_sem))  # later acoustic layers never reach the output
        margins = [] if record is not None else None
        codes = self._rvq("semantic", hb, n_sem, margins) + self._rvq("acoustic", hb, n_ac, margins)
        out = torch.stack(codes, 1)[:, : self.valid_q]      # [B, Q, T]
        if record is not None:
            record["margins"] = np.stack(margins, 1)[:, : self.valid_q]
        return out.numpy()
