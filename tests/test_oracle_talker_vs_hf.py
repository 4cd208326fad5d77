"""Third-party anchor for the oracle's decoder stacks (parity is unpinned by the reference, so every independent pin counts):
HuggingFace `transformers` ships `Qwen3Model` -- the architecture `Qwen3Talker` / `Qwen3CodePredictor` restate in Swift (pre-norm blocks, per-head
RMSNorm on q and k, rotate-half RoPE, GQA, additive causal mask, SwiGLU; Model/Qwen3Layers.swift:128-262, Model/Qwen3CodePredictor.swift:8-216).
With the SAME weights, the oracle's `forward` / code-predictor layers must reproduce its hidden states, for a prefill and for cached decode steps."""
import numpy as np
import pytest
import torch

from conftest import ckpt
from oracle import talker as ot

transformers = pytest.importorskip("transformers")


def _hf_stack(orc, prefix, layers, hidden, inter, nh, nkv, hd, eps, theta, final_norm):
    from transformers import Qwen3Config, Qwen3Model

    cfg = Qwen3Config(vocab_size=32, hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers, num_attention_heads=nh,
                      num_key_value_heads=nkv, head_dim=hd, rms_norm_eps=eps, rope_theta=theta, attention_bias=False,
                      max_position_embeddings=4096, use_sliding_window=False, attn_implementation="eager")
    m = Qwen3Model(cfg).to(torch.float32).eval()
    sd = m.state_dict()
    with torch.no_grad():
        for k in sd:
            if k == "embed_tokens.weight":
                continue
            src = final_norm if k == "norm.weight" else prefix + k
            assert src in orc.w, f"no oracle tensor for {k}"
            sd[k].copy_(orc.w[src])
    return m


@pytest.mark.parametrize("bits,dtype", [(0, "bf16"), (8, "bf16"), (4, "bf16")])
def test_talker_stack_equals_hf_qwen3(bits, dtype):
    """Dense and dequantised (MLX 4 / 8-bit) weights: same hidden states as transformers' Qwen3Model, prefill and three cached decode steps."""
    d = ckpt("tiny", bits, dtype)
    orc = ot.TalkerOracle(d)
    c = orc.cfg
    m = _hf_stack(orc, "", c.num_hidden_layers, c.hidden_size, c.intermediate_size, c.num_attention_heads, c.num_key_value_heads, c.head_dim,
                  c.rms_norm_eps, c.rope_theta, "norm.weight")
    g = torch.Generator().manual_seed(bits)
    x = torch.randn(11, c.hidden_size, generator=g)
    with torch.no_grad():
        out = m(inputs_embeds=x[None], use_cache=True)
        got, cache = orc.forward(x, None, 0)
        scale = out.last_hidden_state.abs().max().item()
        assert (got - out.last_hidden_state[0]).abs().max().item() <= 1e-5 * scale
        past, off = out.past_key_values, 11
        for step in range(3):
            x1 = torch.randn(1, c.hidden_size, generator=g)
            o1 = m(inputs_embeds=x1[None], past_key_values=past, use_cache=True)
            g1, cache = orc.forward(x1, cache, off)
            assert (g1 - o1.last_hidden_state[0]).abs().max().item() <= 1e-5 * scale, f"decode step {step}"
            past, off = o1.past_key_values, off + 1


def test_code_predictor_stack_equals_hf_qwen3():
    """Code predictor: pass 0 (two positions, no cache) and passes 1..3 (one position, cache of g + 1 keys); the oracle applies norm +
    lm_head[g], HF stops at the final norm -- compare through the same head."""
    d = ckpt("tiny", 0, "bf16")
    orc = ot.TalkerOracle(d)
    cp = orc.cfg.code_predictor
    m = _hf_stack(orc, "code_predictor.", cp.num_hidden_layers, cp.hidden_size, cp.intermediate_size, cp.num_attention_heads,
                  cp.num_key_value_heads, cp.head_dim, cp.rms_norm_eps, cp.rope_theta, "code_predictor.norm.weight")
    has_proj = "code_predictor.small_to_mtp_projection.weight" in orc.w
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, orc.cfg.hidden_size, generator=g)
    with torch.no_grad():
        xin = orc.linear("code_predictor.small_to_mtp_projection", x) if has_proj else x
        out = m(inputs_embeds=xin[None], use_cache=True)
        logits, cache = orc.cp_forward(x, None, 0)
        want = orc.linear("code_predictor.lm_head.0", out.last_hidden_state[0])
        scale = want.abs().max().item()
        assert (logits - want).abs().max().item() <= 1e-5 * scale
        past = out.past_key_values
        for step in (1, 2, 3):
            x1 = torch.randn(1, orc.cfg.hidden_size, generator=g)
            x1in = orc.linear("code_predictor.small_to_mtp_projection", x1) if has_proj else x1
            o1 = m(inputs_embeds=x1in[None], past_key_values=past, use_cache=True)
            l1, cache = orc.cp_forward(x1, cache, step)
            w1 = orc.linear(f"code_predictor.lm_head.{step}", o1.last_hidden_state[0])
            assert (l1 - w1).abs().max().item() <= 1e-5 * scale, f"pass {step}"
            past = o1.past_key_values


def test_code_predictor_loop_equals_hf_qwen3_omni_code_predictor():
    """The per-frame loop of the code predictor (Qwen3Talker.swift:501-523; Qwen3CodePredictor.swift:158-212): pass 0 = [talker hidden,
    codec_embedding(code0)] -> lm_head[0]; pass g >= 1 = code_predictor.codec_embedding[g - 1](code_g) -> lm_head[g], KV cache growing by one.
    HuggingFace's Qwen3-Omni talker code predictor implements the same multi-token-prediction loop with the SAME checkpoint key names
    (`model.codec_embedding.{i}`, `lm_head.{i}`, `generation_steps`): drive both with the same greedy tokens, compare every pass's logits."""
    hf = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    from transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe import Qwen3OmniMoeTalkerCodePredictorConfig

    d = ckpt("tiny-cp", 0, "bf16")
    orc = ot.TalkerOracle(d)
    cp = orc.cfg.code_predictor
    assert "code_predictor.small_to_mtp_projection.weight" not in orc.w and cp.hidden_size == orc.cfg.hidden_size
    cfg = Qwen3OmniMoeTalkerCodePredictorConfig(vocab_size=cp.vocab_size, hidden_size=cp.hidden_size, intermediate_size=cp.intermediate_size,
                                                num_hidden_layers=cp.num_hidden_layers, num_attention_heads=cp.num_attention_heads,
                                                num_key_value_heads=cp.num_key_value_heads, head_dim=cp.head_dim, rms_norm_eps=cp.rms_norm_eps,
                                                rope_theta=cp.rope_theta, num_code_groups=cp.num_code_groups, attn_implementation="eager")
    m = hf.Qwen3OmniMoeTalkerCodePredictorModelForConditionalGeneration(cfg).to(torch.float32).eval()
    sd = m.state_dict()
    with torch.no_grad():
        for k in sd:
            src = "code_predictor." + (k[len("model."):] if k.startswith("model.") else k)
            assert src in orc.w, f"no oracle tensor for {k}"
            sd[k].copy_(orc.w[src])
    g = torch.Generator().manual_seed(5)
    hidden = torch.randn(1, orc.cfg.hidden_size, generator=g)  # the talker's last hidden state
    code0 = 17
    ce = orc.w["codec_embedding.weight"]
    with torch.no_grad():
        inp = torch.cat([hidden, ce[[code0]]], 0)
        lg, cache = orc.cp_forward(inp, None, 0)
        out = m(inputs_embeds=inp[None], use_cache=True)
        scale = float(out.logits.abs().max())
        assert (lg[-1] - out.logits[0, -1]).abs().max().item() <= 1e-5 * scale
        tok = int(torch.argmax(lg[-1]))
        past, steps = out.past_key_values, out.generation_steps
        for gi in range(1, cp.num_code_groups - 1):
            assert steps == gi
            mine_in = orc.w[f"code_predictor.codec_embedding.{gi - 1}.weight"][[tok]]
            lg, cache = orc.cp_forward(mine_in, cache, gi)
            out = m(input_ids=torch.tensor([[tok]]), past_key_values=past, use_cache=True, generation_steps=steps)
            assert (lg[-1] - out.logits[0, -1]).abs().max().item() <= 1e-5 * scale, f"pass {gi}"
            tok = int(torch.argmax(lg[-1]))
            past, steps = out.past_key_values, out.generation_steps


def test_interleaved_mrope_with_identical_streams_is_plain_rope():
    """Quirk a7: the reference routes positions through "interleaved MRoPE" (Qwen3Layers.swift:60-92, a port of upstream's
    `apply_interleaved_mrope`) whenever `rope_scaling.mrope_section` is present, but always feeds three IDENTICAL position streams (:77-79) --
    which makes it plain 1-D RoPE.  HF's interleaved-MRoPE rotary embedding (Qwen3-Omni talker) on identical streams against the oracle's
    plain cos / sin tables."""
    hf = pytest.importorskip("transformers.models.qwen3_omni_moe.modeling_qwen3_omni_moe")
    import transformers.models.qwen3_omni_moe.configuration_qwen3_omni_moe as hc

    cfg = hc.Qwen3OmniMoeTalkerTextConfig(hidden_size=256, num_attention_heads=2, num_key_value_heads=1, head_dim=128,
                                          rope_parameters={"rope_type": "default", "rope_theta": 1000000.0, "mrope_section": [24, 20, 20]})
    rot = hf.Qwen3OmniMoeTalkerRotaryEmbedding(cfg)
    pos = torch.arange(3, 230)[None]
    cos, sin = rot(torch.zeros(1, pos.shape[1], 256), pos)  # a 2-D position tensor is expanded to three identical streams
    inv = ot.TalkerOracle._inv_freq(1000000.0, 128)
    fr = pos[0].float()[:, None] * inv[None, :]
    emb = torch.cat([fr, fr], -1)
    assert torch.equal(cos[0], emb.cos()) and torch.equal(sin[0], emb.sin())
