"""CPU: talker / code-predictor / sampler restatement (oracle/talker.py) — self-consistency and the reference's quirks."""
import numpy as np
import pytest
import torch

from conftest import TEXT_IDS, ckpt
from oracle import talker as otalker


@pytest.fixture(scope="module")
def orc():
    return otalker.TalkerOracle(ckpt("tiny", 8))


def test_config_roundtrip(orc):
    c = orc.cfg
    assert (c.hidden_size, c.num_hidden_layers, c.vocab_size, c.head_dim) == (256, 2, 3072, 128)
    assert (c.codec_pad_id, c.codec_bos_id, c.codec_eos_token_id) == (2148, 2149, 2150)  # ConfigTests.swift:6-18 pins these ids
    assert c.code_predictor.num_code_groups == 16 and c.code_predictor.vocab_size == 2048
    assert orc.spk_id["aiden"] == 2861 and orc.bits == 8 and orc.group == 64


def test_incremental_kv_equals_full_forward(orc):
    torch.manual_seed(0)
    x = torch.randn(7, orc.cfg.hidden_size) * 0.3
    full, _ = orc.forward(x, None, 0)
    h, cache = orc.forward(x[:4], None, 0)
    outs = [h]
    for i in range(4, 7):
        h, cache = orc.forward(x[i:i + 1], cache, i)
        outs.append(h)
    assert torch.allclose(torch.cat(outs), full, atol=2e-5)


def test_attention_matches_sdpa(orc):
    torch.manual_seed(1)
    c = orc.cfg
    x = torch.randn(5, c.hidden_size) * 0.3
    o, _ = orc._attention("layers.0.self_attn", x, None, torch.arange(5), orc.inv_freq, c.num_attention_heads, c.num_key_value_heads, c.head_dim, c.rms_norm_eps)
    # independent formulation: HF-style rotate-half RoPE + torch SDPA with GQA expansion
    q = orc.linear("layers.0.self_attn.q_proj", x).view(5, c.num_attention_heads, 128)
    k = orc.linear("layers.0.self_attn.k_proj", x).view(5, c.num_key_value_heads, 128)
    v = orc.linear("layers.0.self_attn.v_proj", x).view(5, c.num_key_value_heads, 128)
    q = orc.rms_norm(q, orc.w["layers.0.self_attn.q_norm.weight"], c.rms_norm_eps)
    k = orc.rms_norm(k, orc.w["layers.0.self_attn.k_norm.weight"], c.rms_norm_eps)
    ang = torch.arange(5, dtype=torch.float32)[:, None] * orc.inv_freq[None]
    cos, sin = torch.cat([ang, ang], -1).cos()[:, None], torch.cat([ang, ang], -1).sin()[:, None]
    rot = lambda t: torch.cat([-t[..., 64:], t[..., :64]], -1)
    q, k = q * cos + rot(q) * sin, k * cos + rot(k) * sin
    g = c.num_attention_heads // c.num_key_value_heads
    a = torch.nn.functional.scaled_dot_product_attention(q.transpose(0, 1)[None], k.transpose(0, 1).repeat_interleave(g, 0)[None],
                                                         v.transpose(0, 1).repeat_interleave(g, 0)[None], is_causal=True)[0]
    want = orc.linear("layers.0.self_attn.o_proj", a.transpose(0, 1).reshape(5, -1))
    assert torch.allclose(o, want, atol=2e-5)


def test_prompt_layout(orc):
    """Prefill = [instruct?] + role(3) + combined(n-1) + first-text(1); trailing = text[4:-5] + tts_eos (Qwen3Talker.swift:381-433)."""
    e, tr, pad = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861))
    assert e.shape[0] == 3 + 5 + 1 and tr.shape[0] == len(TEXT_IDS) - 9 + 1
    e2, tr2, _ = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS[:11]))  # "Hello world!": 11 ids, no speaker -> 3 + 4 + 1
    assert e2.shape[0] == 8 and tr2.shape[0] == 3
    e3, _, _ = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, instruct_ids=[1, 2, 3]))
    assert e3.shape[0] == 9 + 3 and torch.equal(e3[3:], e)
    rc = np.zeros((16, 5), np.int32)
    e4, _, _ = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, ref_text_ids=[4, 5], ref_codes=rc))
    assert e4.shape[0] == 9 + 2 + 5
    # instruct wins over ICL (:389-395)
    e5, _, _ = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, instruct_ids=[1, 2, 3], ref_text_ids=[4, 5], ref_codes=rc))
    assert e5.shape[0] == 12
    # the speaker row is codec_embedding[spk_id]; a raw embedding takes the same place
    emb = torch.randn(orc.cfg.hidden_size)
    e6, _, _ = orc.build_prompt(otalker.Request(text_ids=TEXT_IDS, speaker_embedding=emb.numpy()))
    assert e6.shape[0] == 9
    tts_pad = orc.text_project([orc.cfg.tts_pad_token_id])
    assert torch.allclose(e6[6], tts_pad[0] + emb, atol=1e-6)


def test_sampler_quirks(orc):
    V = orc.cfg.vocab_size
    lg = np.full(V, -5.0, np.float32)
    lg[10], lg[20] = -1.0, -1.02
    greedy = otalker.Request(text_ids=[], temperature=0.0)
    assert orc.sample(lg, greedy, None, 0)[0] == 10
    assert orc.sample(lg, greedy, {20}, 0)[0] == 20  # division regardless of sign makes a negative logit LARGER (quirk 3)
    lg[2500] = 3.0
    assert orc.sample(lg, greedy, None, 0)[0] == 2500  # greedy returns before the valid-id mask (quirk 4)
    hot = otalker.Request(text_ids=[], temperature=0.7, seed=5)
    ids = {orc.sample(lg, hot, None, c)[0] for c in range(200)}
    assert all(i < 2048 or i in (2148, 2150) for i in ids) and 2500 not in ids
    # ties resolve to the first index
    lg2 = np.zeros(2048, np.float32)
    assert orc.sample(lg2, greedy, None, 0)[0] == 0
    # top-k keeps exactly k (no ties here); counter changes the draw, same counter reproduces it
    lg3 = np.arange(2048, dtype=np.float32) / 100
    tk = otalker.Request(text_ids=[], temperature=1.0, top_k=5, seed=1)
    draws = {orc.sample(lg3, tk, None, c)[0] for c in range(300)}
    assert draws <= set(range(2043, 2048)) and len(draws) > 1
    assert orc.sample(lg3, tk, None, 7)[0] == orc.sample(lg3, tk, None, 7)[0]
    tp = otalker.Request(text_ids=[], temperature=1.0, top_p=0.4, seed=1)
    lg4 = np.log(np.array([0.5, 0.3, 0.15, 0.05] + [1e-9] * 2044, np.float32))
    assert {orc.sample(lg4, tp, None, c)[0] for c in range(100)} == {0}  # mass of strictly-more-probable ids < 0.4 only for id 0


def test_counter_uniform_range_and_determinism():
    u = otalker.counter_uniform(3, 9, 100000)
    assert u.min() > 0 and u.max() < 1 and abs(u.mean() - 0.5) < 0.01
    assert np.array_equal(u, otalker.counter_uniform(3, 9, 100000))
    assert not np.array_equal(u, otalker.counter_uniform(3, 10, 100000))


def test_generate_loop_rules(orc):
    req = otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=40)
    rec = {}
    frames = orc.generate_codes(req, record=rec)
    raw = rec["raw_frames"]
    assert len(raw) <= 40 and all(len(f) == 16 for f in raw)
    assert frames == [f for f in raw if 0 <= f[0] < 2048]  # final filter (:571-576)
    # EOS / pad are suppressed while trailing text remains: the first len(trailing) frames cannot be 2150 / 2148 (:470-475)
    n_trailing = len(TEXT_IDS) - 9 + 1
    assert all(f[0] not in (2148, 2150) for f in raw[:n_trailing])
    assert orc.generate_codes(otalker.Request(text_ids=TEXT_IDS[:8], temperature=0.0, max_tokens=5)) == []  # < 9 ids (:348-352)
    # stream variant: no repetition penalty on the code-predictor groups -> different ids sooner or later (quirk 3)
    s = orc.generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=40, stream_variant=True), filter_invalid=False)
    assert s[0] == raw[0] and s != raw


def test_teacher_forcing_and_window(orc):
    F = 210  # > 192 + 15: the sliding window trims at steps 15, 30, ...
    forced = np.random.default_rng(0).integers(0, 2048, (F, 16))
    rec = {}
    out = orc.generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=F), forced=forced, record=rec, filter_invalid=False)
    assert np.array_equal(np.asarray(out), forced) and rec["code0_logits"].shape == (F, 3072) and rec["cp_logits"].shape == (F, 15, 2048)
    assert np.isfinite(rec["code0_logits"]).all()
