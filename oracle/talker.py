"""Talker + code predictor + sampler + AR loop restated on torch-CPU fp32 (oracle; test infrastructure only;
parity unpinned — see `oracle/__init__.py`).

Follows `Model/Qwen3Talker.swift` (forward :73-101, sampler :274-322, generateCodes :327-577, generateStream
:633-885), `Model/Qwen3Layers.swift` (RMSNorm :8-26, RoPE :30-101, attention :128-219, MLP :223-238, layer
:242-262, text projection :266-280, KV trim :105-124) and `Model/Qwen3CodePredictor.swift` (:8-216).

Activations are fp32 throughout (MLX type promotion makes the residual stream fp32 from
layer 0 onward; the stated logit tolerance is against this fp32-activation oracle).  Weights are the exact values
on disk: packed leaves are dequantised with `mlx_quant.dequantize` (fp32), float leaves are widened.

Host-side string work (chat template, BPE) is OUT of this path: requests carry token ids, as the C-ABI does.
"""
from __future__ import annotations

import json
import os
from dataclasses import dataclass, field

import numpy as np
import torch
import torch.nn.functional as F
from safetensors.torch import load_file

from . import mlx_quant
from .checkpoint import CodePredictorDims, TalkerDims

MAX_KV_WINDOW = 192  # Qwen3Layers.swift:108
MIN_TEXT_TOKENS = 9  # Qwen3Talker.swift:348
CODEBOOK_SIZE = 2048  # Qwen3Talker.swift:23, 573


@dataclass
class Request:
    """One utterance, in ids (mirrors `q3tts_request` in include/qwen3tts_b200.h)."""
    text_ids: list  # ids of "<|im_start|>assistant\n{text}<|im_end|>\n<|im_start|>assistant\n" (Qwen3Talker.swift:344)
    speaker_id: int = -1  # codec_embedding row (config.spk_id[name]); -1 = none (:370-373)
    speaker_embedding: np.ndarray | None = None  # raw [H] vector (:374-376)
    instruct_ids: list | None = None  # ids of "<|im_start|>user\n{instruct}<|im_end|>\n" (:389-394, 408-413)
    ref_text_ids: list | None = None  # ICL transcript ids (:396-398)
    ref_codes: np.ndarray | None = None  # ICL codes [16][T_ref]; only row 0 is used (:402-403)
    temperature: float = 0.9
    top_k: int = 0
    top_p: float = 1.0  # extension (the reference has no top-p); 1.0 = off
    repetition_penalty: float = 1.05
    max_tokens: int = 1200
    seed: int = 0
    stream_variant: bool = False  # True = generateStream (no repetition penalty on code-predictor groups, :821)


def load_config(model_dir: str) -> tuple[TalkerDims, dict]:
    """`Qwen3TTSConfig.init(from:)` (Qwen3Config.swift:208-253)."""
    raw = json.load(open(os.path.join(model_dir, "config.json")))
    src = raw.get("talker_config", raw)
    t = TalkerDims()
    for k in ("hidden_size", "num_hidden_layers", "vocab_size", "text_vocab_size", "num_attention_heads",
              "intermediate_size", "rms_norm_eps", "max_position_embeddings", "rope_theta"):
        setattr(t, k, src[k])
    for k in ("text_hidden_size", "num_key_value_heads", "head_dim", "codec_bos_id", "codec_eos_token_id", "codec_pad_id",
              "codec_nothink_id", "codec_think_bos_id", "codec_think_eos_id"):
        if k in src:
            setattr(t, k, src[k])
    for k in ("tts_bos_token_id", "tts_eos_token_id", "tts_pad_token_id"):
        if k in raw:
            setattr(t, k, raw[k])
    cp = CodePredictorDims()
    for k, v in src.get("code_predictor_config", {}).items():
        if hasattr(cp, k):
            setattr(cp, k, v)
    t.code_predictor = cp
    rs = src.get("rope_scaling")
    t.mrope_section = rs.get("mrope_section") if rs else None
    t.tts_model_type = raw.get("tts_model_type")
    extra = {"spk_id": src.get("spk_id", {}), "quantization": raw.get("quantization"),
             "quantization_config": raw.get("quantization_config")}
    return t, extra


def counter_uniform(seed: int, counter: int, n: int) -> np.ndarray:
    """Counter-based uniforms in (0,1): splitmix64(seed, counter, index) -> 23-bit mantissa.  OUR documented
    stream (MLX's threefry stream is not reproducible outside MLX); the CUDA sampler
    implements the same integer hash."""
    M = np.uint64(0xFFFFFFFFFFFFFFFF)
    with np.errstate(over="ignore"):
        x = (np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15) + np.uint64(counter) * np.uint64(0xD1B54A32D192ED03)
             + np.arange(n, dtype=np.uint64) * np.uint64(0x8CB92BA72F3D8DD7) + np.uint64(0x2545F4914F6CDD1D)) & M
        x ^= x >> np.uint64(30)
        x = (x * np.uint64(0xBF58476D1CE4E5B9)) & M
        x ^= x >> np.uint64(27)
        x = (x * np.uint64(0x94D049BB133111EB)) & M
        x ^= x >> np.uint64(31)
    return ((x >> np.uint64(41)).astype(np.float32) + np.float32(0.5)) * np.float32(1.0 / 8388608.0)


class TalkerOracle:
    def __init__(self, model_dir: str, runtime_quantization: bool = False):
        """runtime_quantization: `Qwen3TTSPipelineConfiguration.applyRuntimeQuantization` (Qwen3TTSPipeline.swift:25, 184, 961-980) -- a
        checkpoint without a `quantization` block has every Linear / Embedding weight MLX-quantised at load (group 64; 6 bits for embeddings,
        q/k/v projections and the heads, 4 bits for the rest); the oracle multiplies with the dequantised values."""
        self.cfg, extra = load_config(model_dir)
        self.spk_id = extra["spk_id"]
        # usePreQuantized = config.quantization != nil (Qwen3Talker.swift:139); otherwise packed leaves are dequantised offline
        # to fp16 with quantization_config's (group_size ?? 64, bits ?? 8) (:141-164)
        self.pre_quantized = extra["quantization"] is not None
        q = extra["quantization"] if self.pre_quantized else (extra["quantization_config"] or {})
        self.bits = int(q.get("bits") or (0 if self.pre_quantized else 8))
        self.group = int(q.get("group_size", 64))
        deq_dtype = "f32" if self.pre_quantized else "f16"
        raw = load_file(os.path.join(model_dir, "model.safetensors"))
        self.w: dict[str, torch.Tensor] = {}
        # key remap of Qwen3Talker.load (:117-137)
        tmp = {}
        for k, v in raw.items():
            if k.startswith("audio_decoder."):
                continue
            nk = k[len("talker."):] if k.startswith("talker.") else k
            if nk.startswith("code_predictor.model."):
                nk = "code_predictor." + nk[len("code_predictor.model."):]
            if nk.startswith("model."):
                nk = nk[len("model."):]
            tmp[nk] = v
        for k, v in tmp.items():
            if k.endswith(".scales") or k.endswith(".biases"):
                continue
            if k.endswith(".weight") and v.dtype in (torch.uint32, torch.int32) and k[:-7] + ".scales" in tmp:
                base = k[:-7]
                packed = v.view(torch.int32).numpy().view(np.uint32)
                self.w[k] = torch.from_numpy(mlx_quant.dequantize(packed, tmp[base + ".scales"], tmp[base + ".biases"],
                                                                  self.group, self.bits, deq_dtype))
            else:
                self.w[k] = v.to(torch.float32)
        if runtime_quantization and not self.pre_quantized:
            names = {torch.bfloat16: "bf16", torch.float16: "f16", torch.float32: "f32"}
            for k in list(self.w):
                w = self.w[k]
                if not k.endswith(".weight") or w.ndim != 2 or w.shape[1] % 64 != 0:
                    continue
                raw = tmp[k]
                sdt = "f16" if raw.dtype in (torch.uint32, torch.int32) else names[raw.dtype]  # offline-dequantised leaves are fp16
                bits = mlx_quant.runtime_bits(k)
                out_dt = sdt if "embed" in k.lower() else "f32"  # QuantizedEmbedding hands out rows in the storage dtype
                self.w[k] = torch.from_numpy(mlx_quant.fake_quantize(w.numpy(), bits, sdt, out_dt))
        c = self.cfg
        self.inv_freq = self._inv_freq(c.rope_theta, c.head_dim)
        self.cp_inv_freq = self._inv_freq(c.code_predictor.rope_theta, c.code_predictor.head_dim)

    @staticmethod
    def _inv_freq(base, dim):
        # (0..<dim/2).map { 1.0 / pow(base, Float($0 * 2) / Float(dim)) }  — fp32 (Qwen3Layers.swift:45)
        e = torch.arange(0, dim, 2, dtype=torch.float32) / torch.tensor(float(dim), dtype=torch.float32)
        return 1.0 / torch.pow(torch.tensor(base, dtype=torch.float32), e)

    # ------------------------------------------------------------------ building blocks
    def linear(self, name, x):
        y = x @ self.w[name + ".weight"].T
        if name + ".bias" in self.w:
            y = y + self.w[name + ".bias"]
        return y

    @staticmethod
    def rms_norm(x, w, eps):
        return x * torch.rsqrt(x.pow(2).mean(-1, keepdim=True) + eps) * w

    def text_project(self, ids):
        """`encodeText` = text_projection(text_embedding(ids)) (Qwen3Talker.swift:103-106; Qwen3Layers.swift:276-279)."""
        e = self.w["text_embedding.weight"][torch.as_tensor(ids, dtype=torch.long)]
        return self.linear("text_projection.linear_fc2", F.silu(self.linear("text_projection.linear_fc1", e)))

    def _attention(self, prefix, x, cache, positions, inv_freq, nh, nkv, hd, eps):
        """`Qwen3Attention` / `CodePredictorAttention` (Qwen3Layers.swift:167-218; Qwen3CodePredictor.swift:67-113).
        With the three identical position streams the reference feeds, interleaved MRoPE == plain RoPE (:75-92)."""
        L = x.shape[0]
        q = self.linear(prefix + ".q_proj", x).view(L, nh, hd)
        k = self.linear(prefix + ".k_proj", x).view(L, nkv, hd)
        v = self.linear(prefix + ".v_proj", x).view(L, nkv, hd)
        q = self.rms_norm(q, self.w[prefix + ".q_norm.weight"], eps).transpose(0, 1)
        k = self.rms_norm(k, self.w[prefix + ".k_norm.weight"], eps).transpose(0, 1)
        v = v.transpose(0, 1)
        fr = positions.to(torch.float32)[:, None] * inv_freq[None, :]
        emb = torch.cat([fr, fr], -1)
        cos, sin = emb.cos()[None], emb.sin()[None]

        def rot(t):
            h = t.shape[-1] // 2
            return torch.cat([-t[..., h:], t[..., :h]], -1)

        q = q * cos + rot(q) * sin
        k = k * cos + rot(k) * sin
        if cache is not None:
            k = torch.cat([cache[0], k], 1)
            v = torch.cat([cache[1], v], 1)
        new_cache = (k, v)
        g = nh // nkv
        kk = k.repeat_interleave(g, 0) if g > 1 else k
        vv = v.repeat_interleave(g, 0) if g > 1 else v
        s = (q @ kk.transpose(-1, -2)) * (1.0 / np.sqrt(np.float32(hd)))
        if L > 1:  # additive causal mask, only when L > 1; with a cache the reference passes none for L == 1
            S = kk.shape[1]
            m = torch.triu(torch.full((L, L), -1e9), diagonal=1)
            if S > L:  # (never happens in the reference: L > 1 only without a cache, except CP pass 0)
                m = torch.cat([torch.zeros(L, S - L), m], 1)
            s = s + m
        o = torch.softmax(s, -1) @ vv
        o = o.transpose(0, 1).reshape(L, nh * hd)
        return self.linear(prefix + ".o_proj", o), new_cache

    def _layer(self, prefix, x, cache, positions, inv_freq, nh, nkv, hd, eps):
        r, c = self._attention(prefix + ".self_attn", self.rms_norm(x, self.w[prefix + ".input_layernorm.weight"], eps),
                               cache, positions, inv_freq, nh, nkv, hd, eps)
        h = x + r
        n = self.rms_norm(h, self.w[prefix + ".post_attention_layernorm.weight"], eps)
        m = self.linear(prefix + ".mlp.down_proj", F.silu(self.linear(prefix + ".mlp.gate_proj", n)) * self.linear(prefix + ".mlp.up_proj", n))
        return h + m, c

    def forward(self, x, cache, offset):
        """`Qwen3Talker.callAsFunction` (Qwen3Talker.swift:73-101). x [L,H]; cache list of (k,v) or None."""
        c = self.cfg
        L = x.shape[0]
        pos = torch.arange(offset, offset + L)
        new = []
        h = x
        for i in range(c.num_hidden_layers):
            h, kv = self._layer(f"layers.{i}", h, None if cache is None else cache[i], pos, self.inv_freq,
                                c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.rms_norm_eps)
            new.append(kv)
        return self.rms_norm(h, self.w["norm.weight"], c.rms_norm_eps), new

    def cp_forward(self, x, cache, step):
        """`Qwen3CodePredictor.callAsFunction` (Qwen3CodePredictor.swift:180-215)."""
        cp = self.cfg.code_predictor
        if "code_predictor.small_to_mtp_projection.weight" in self.w:
            x = self.linear("code_predictor.small_to_mtp_projection", x)
        L = x.shape[0]
        offset = 0 if cache is None else cache[0][0].shape[1]
        pos = torch.arange(offset, offset + L)
        new = []
        for i in range(cp.num_hidden_layers):
            x, kv = self._layer(f"code_predictor.layers.{i}", x, None if cache is None else cache[i], pos, self.cp_inv_freq,
                                cp.num_attention_heads, cp.num_key_value_heads, cp.head_dim, cp.rms_norm_eps)
            new.append(kv)
        x = self.rms_norm(x, self.w["code_predictor.norm.weight"], cp.rms_norm_eps)
        return self.linear(f"code_predictor.lm_head.{step}", x), new

    # ------------------------------------------------------------------ a14 sampler
    def sample(self, logits: np.ndarray, req: Request, token_set, counter: int):
        """`sampleToken` (Qwen3Talker.swift:274-322) on the last-position logits [V] (fp32 numpy).
        Returns (id, processed_logits_before_sampling)."""
        lg = np.array(logits, dtype=np.float32, copy=True)
        V = lg.shape[0]
        if token_set and req.repetition_penalty != 1.0:
            pen = np.ones(V, dtype=np.float32)
            for t in token_set:
                if 0 <= t < V:
                    pen[t] = np.float32(req.repetition_penalty)
            lg = lg / pen  # division regardless of sign (:288-299)
        if not req.temperature > 0:
            return int(np.argmax(lg)), lg  # greedy returns BEFORE the valid-token mask (:301-305)
        lg = lg / np.float32(req.temperature)
        if 0 < req.top_k < V:
            thr = np.sort(lg)[V - req.top_k]
            lg = np.where(lg < thr, -np.inf, lg).astype(np.float32)
        if V == self.cfg.vocab_size:  # valid-token mask only for the codec-head vocabulary (:316-319, 19-33)
            idx = np.arange(V)
            valid = (idx < CODEBOOK_SIZE) | (idx == 2148) | (idx == 2150)
            lg = np.where(valid, lg, -np.inf).astype(np.float32)
        if req.top_p < 1.0:  # extension: keep i iff mass of strictly-more-probable tokens < top_p
            m = lg.max()
            p = np.exp((lg - m).astype(np.float64))
            p = p / p.sum()
            order = np.argsort(-p, kind="stable")
            ps = p[order]
            # mass of strictly greater probabilities (ties share the same prefix mass)
            cum = np.concatenate([[0.0], np.cumsum(ps)[:-1]])
            first_of_tie = np.concatenate([[True], ps[1:] != ps[:-1]])
            greater = np.maximum.accumulate(np.where(first_of_tie, cum, 0.0))
            keep = np.zeros(V, dtype=bool)
            keep[order] = greater < req.top_p
            lg = np.where(keep, lg, -np.inf).astype(np.float32)
        u = counter_uniform(req.seed, counter, V)
        g = -np.log(-np.log(u))  # Gumbel-max == categorical
        z = np.where(np.isfinite(lg), lg + g.astype(np.float32), -np.inf)
        return int(np.argmax(z)), lg

    # ------------------------------------------------------------------ prompt (Qwen3Talker.swift:344-433)
    def build_prompt(self, req: Request):
        c, w = self.cfg, self.w
        ids = list(req.text_ids)
        tts = self.text_project([c.tts_bos_token_id, c.tts_eos_token_id, c.tts_pad_token_id])
        tts_bos, tts_eos, tts_pad = tts[0:1], tts[1:2], tts[2:3]
        ce = w["codec_embedding.weight"]
        codec = ce[[c.codec_nothink_id, c.codec_think_bos_id, c.codec_think_eos_id]]
        suffix = ce[[c.codec_pad_id, c.codec_bos_id]]
        if req.speaker_id >= 0:
            codec = torch.cat([codec, ce[[req.speaker_id]], suffix], 0)
        elif req.speaker_embedding is not None:
            codec = torch.cat([codec, torch.as_tensor(req.speaker_embedding, dtype=torch.float32).reshape(1, -1), suffix], 0)
        else:
            codec = torch.cat([codec, suffix], 0)
        role = self.text_project(ids[0:3])
        n = codec.shape[0]
        combined = torch.cat([tts_pad.repeat(n - 2, 1), tts_bos], 0) + codec[: n - 1]
        instruct = None
        use_icl = req.ref_codes is not None and req.ref_text_ids is not None and len(req.ref_text_ids) > 0
        if req.instruct_ids is not None and len(req.instruct_ids) > 0:
            instruct = self.text_project(req.instruct_ids)
        elif use_icl:
            instruct = self.text_project(req.ref_text_ids)
            rc = np.asarray(req.ref_codes)
            if rc.size > 0 and rc.shape[1] > 0:
                instruct = torch.cat([instruct, ce[torch.as_tensor(rc[0], dtype=torch.long)]], 0)
        parts = ([instruct] if instruct is not None else []) + [role, combined]
        first_text = self.text_project(ids[3:4]) + codec[n - 1:]
        embeds = torch.cat(parts + [first_text], 0)
        trailing_len = len(ids) - 4 - 5
        if trailing_len > 0:
            trailing = torch.cat([self.text_project(ids[4: len(ids) - 5]), tts_eos], 0)
        else:
            trailing = tts_eos
        return embeds, trailing, tts_pad

    # ------------------------------------------------------------------ a2 AR loop
    def generate_codes(self, req: Request, forced=None, record: dict | None = None, filter_invalid=True):
        """`generateCodes` (:327-577) or, with `req.stream_variant`, the loop of `generateStream` (:633-885).

        forced: optional int array [F,16]; when given, the id fed back at every sample point is forced[f][g]
                (teacher forcing) and generation runs exactly F frames (EOS/pad stop rules are skipped).
        record: optional dict; receives 'code0_logits' [F,V0] (raw codec_head logits used at each frame),
                'cp_logits' [F,15,Vc], 'margins' [F,16] (top-1 minus top-2 of the processed logits) and
                'raw_frames' (before the code0 < 2048 filter).
        Returns list of frames (each 16 ints), filtered to code0 in [0, 2048) (:571-576) unless told otherwise.
        """
        c = self.cfg
        cp = c.code_predictor
        G = cp.num_code_groups
        if len(req.text_ids) < MIN_TEXT_TOKENS:
            return []
        embeds, trailing, tts_pad = self.build_prompt(req)
        h, cache = self.forward(embeds, None, 0)
        pos = embeds.shape[0]
        logits = self.linear("codec_head", h)  # all positions (quirk 6); sampler takes the last
        frames = []
        set0: set = set()
        sets = [set() for _ in range(G - 1)]
        eos, pad = c.codec_eos_token_id, c.codec_pad_id
        trailing_idx, consecutive_pad, counter = 0, 0, 0
        total_text = trailing.shape[0]
        ce = self.w["codec_embedding.weight"]
        rec0, rec_cp, margins = [], [], []
        n_steps = req.max_tokens if forced is None else len(forced)
        for step in range(n_steps):
            last = logits[-1].numpy().copy()
            if record is not None:
                rec0.append(last.copy())
            if trailing_idx < total_text:  # EOS/pad suppression while text remains (:470-475)
                last[eos] = -np.inf
                last[pad] = -np.inf
            code0, proc = self.sample(last, req, set0 if set0 else None, counter)
            counter += 1
            mrow = [_margin(proc)]
            if forced is not None:
                code0 = int(forced[step][0])
            else:
                if code0 == eos:
                    break
                if code0 == pad:
                    consecutive_pad += 1
                    if consecutive_pad > 6:
                        break
                else:
                    consecutive_pad = 0
            codes = [code0]
            code_hidden = h[-1:]
            cp_cache = None
            cp_rows = []
            for gi in range(G - 1):
                if gi == 0:
                    inp = torch.cat([code_hidden, ce[[code0]]], 0)
                else:
                    inp = self.w[f"code_predictor.codec_embedding.{gi - 1}.weight"][[codes[gi]]]
                lg, cp_cache = self.cp_forward(inp, cp_cache, gi)
                row = lg[-1].numpy()
                if record is not None:
                    cp_rows.append(row.copy())
                ts = None if req.stream_variant else (sets[gi] if sets[gi] else None)  # quirk 3
                tok, proc = self.sample(row, req, ts, counter)
                counter += 1
                mrow.append(_margin(proc))
                if forced is not None:
                    tok = int(forced[step][gi + 1])
                codes.append(tok)
                sets[gi].add(tok)
            frames.append(codes)
            set0.add(code0)
            if record is not None:
                rec_cp.append(np.stack(cp_rows))
                margins.append(mrow)
            if trailing_idx < total_text:
                text_embed = trailing[trailing_idx: trailing_idx + 1]
                trailing_idx += 1
            else:
                text_embed = tts_pad
            s = ce[[code0]]
            for gi in range(G - 1):
                s = s + self.w[f"code_predictor.codec_embedding.{gi}.weight"][[codes[gi + 1]]]
            x = text_embed + s
            h, cache = self.forward(x, cache, pos)
            logits = self.linear("codec_head", h)
            pos += 1
            if (step + 1) % 15 == 0 and cache[0][0].shape[1] > MAX_KV_WINDOW:  # trimKVCache (:556-558; Qwen3Layers.swift:111-124)
                cache = [(k[:, -MAX_KV_WINDOW:], v[:, -MAX_KV_WINDOW:]) for k, v in cache]
        if record is not None:
            record["code0_logits"] = np.stack(rec0) if rec0 else np.zeros((0, c.vocab_size), np.float32)
            record["cp_logits"] = np.stack(rec_cp) if rec_cp else np.zeros((0, G - 1, cp.vocab_size), np.float32)
            record["margins"] = np.asarray(margins, dtype=np.float32)
            record["raw_frames"] = [list(f) for f in frames]
        if filter_invalid:
            frames = [f for f in frames if 0 <= f[0] < CODEBOOK_SIZE]
        return frames


def _margin(lg: np.ndarray) -> float:
    f = lg[np.isfinite(lg)]
    if f.size < 2:
        return float("inf")
    top2 = np.partition(f, -2)[-2:]
    return float(top2[1] - top2[0])
