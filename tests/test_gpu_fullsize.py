"""GPU parity at the BASELINE.json configurations, at full model size and at the stated bar (no relative tolerances here).

* talker: teacher-forced logits max-abs <= 1e-2 and greedy ids identical wherever the oracle's top-2 margin exceeds 2e-2, for the
  batched tensor-core decode step at **64 rows** (64 distinct utterances in one handle), on checkpoints drawn with BASELINE.md's
  init (every matrix and head N(0, 0.02^2), norms = 1; `oracle.checkpoint` init="baseline"):
    config 1  0.6B 8-bit g64        config 2  0.6B 4-bit g64        config 3  1.7B bf16 (assumed dims, 2048 -> 1024 projection)
* codec (config 4): PCM SNR >= 40 dB at the full decoder dimensions for T = 26 (stream window), 110 (chunkedDecode window) and
  750 (whole 60 s clip), default fp16 vocoder residual stream.
Model/Qwen3Talker.swift:464-562; Vocoder/SpeechTokenizer.swift:917-952.
"""
import numpy as np
import pytest
import torch

from conftest import ckpt

pytestmark = [pytest.mark.gpu, pytest.mark.slow]

LOGIT_TOL = 1e-2
MARGIN_TOL = 2e-2
SPEAKERS = [3066, 3065, 3010, 3061, 2861, 2873, 2864, 2875, 2878]

CONFIGS = {
    "cfg1_0.6b_8bit": ("0.6b", 8),
    "cfg2_0.6b_4bit": ("0.6b", 4),
    "cfg3_1.7b_bf16": ("1.7b", 0),
}


def _requests(q, n, frames, forced=None, **kw):
    rng = np.random.default_rng(11)
    reqs = []
    for i in range(n):
        ids = rng.integers(0, 150000, size=int(rng.integers(8, 41)) + 9).tolist()
        reqs.append(q.GenRequest(text_ids=ids, speaker_id=SPEAKERS[i % len(SPEAKERS)], temperature=0.0, max_tokens=frames, keep_invalid_frames=True,
                                 forced_codes=None if forced is None else forced[i], **kw))
    return reqs


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_batch64_tensor_core_step_strict_logits_and_greedy(cfg):
    import qwen3tts_b200 as q
    from oracle import talker as otalker

    name, bits = CONFIGS[cfg]
    d = ckpt(name, bits, init="baseline")
    orc = otalker.TalkerOracle(d)  # not from the session cache: a full-size fp32 copy is 4-10 GB of host memory, dropped with the test
    B, F = 64, 3
    forced = np.random.default_rng(4).integers(0, 2048, size=(B, F, 16)).astype(np.int32)
    eng = q.Engine(d, max_batch=B, max_frames=64, load_codec=False)
    try:
        assert eng.info.quant_bits == bits
        # ---- teacher-forced logits of three of the 64 rows (first, middle, last slot), all 64 utterances in flight
        worst0 = worstc = 0.0
        for j in (0, 31, 63):
            reqs = _requests(q, B, F, forced)
            reqs[j].want_logits = F
            outs, lg = eng.generate_codes_batch(reqs)
            t = eng.timing()
            assert t.persistent_launches == 0 and t.graph_replays >= F, "the 64-row step must run the batched tensor-core path"
            assert all(o.tolist() == forced[i].tolist() for i, o in enumerate(outs))
            rec = {}
            orc.generate_codes(otalker.Request(text_ids=reqs[j].text_ids, speaker_id=reqs[j].speaker_id, temperature=0.0, max_tokens=F), forced=forced[j],
                               record=rec, filter_invalid=False)
            e0 = float(np.abs(lg["code0_logits"] - rec["code0_logits"]).max())
            ec = float(np.abs(lg["cp_logits"] - rec["cp_logits"]).max())
            print(f"[{cfg}] row {j} of 64: teacher-forced max-abs logit error code0 {e0:.3e} (logit rms {rec['code0_logits'].std():.2f}), "
                  f"code predictor {ec:.3e} (rms {rec['cp_logits'].std():.2f})")
            worst0, worstc = max(worst0, e0), max(worstc, ec)
        assert worst0 <= LOGIT_TOL and worstc <= LOGIT_TOL, (worst0, worstc)
        # ---- greedy ids of the same batch, free running: identical wherever the oracle's margin allows
        G = 4
        outs = eng.generate_codes_batch(_requests(q, B, G))
        reqs = _requests(q, B, G)
        checked = flips = 0
        for i in (0, 9, 22, 37, 50, 63):
            rec = {}
            want = orc.generate_codes(otalker.Request(text_ids=reqs[i].text_ids, speaker_id=reqs[i].speaker_id, temperature=0.0, max_tokens=G), record=rec,
                                      filter_invalid=False)
            got = outs[i].tolist()
            for f in range(min(len(got), len(want))):
                if got[f] != want[f]:
                    g = next(k for k in range(16) if got[f][k] != want[f][k])
                    assert rec["margins"][f][g] < MARGIN_TOL, f"[{cfg}] utterance {i} frame {f} group {g}: ids diverge at margin {rec['margins'][f][g]:.4f}"
                    flips += 1
                    break
                checked += 1
        print(f"[{cfg}] greedy: {checked} frames identical to the oracle, {flips} utterances left it at a near-tie (margin < {MARGIN_TOL})")
    finally:
        eng.close()


def test_stress_init_logits_scale_with_logit_rms():
    """The same 64-row step on the unit tests' *stress* init (heads N(0, 0.25^2): logit rms ~8, 13x the BASELINE init's): the
    fp16-operand error scales with the logits, so this case is held to a RELATIVE bar (5e-3 x rms); the absolute 1e-2 bar is the
    test above."""
    import qwen3tts_b200 as q
    from oracle import talker as otalker

    d = ckpt("0.6b", 4)
    B, F = 64, 3
    forced = np.random.default_rng(4).integers(0, 2048, size=(B, F, 16)).astype(np.int32)
    eng = q.Engine(d, max_batch=B, max_frames=64, load_codec=False)
    try:
        reqs = _requests(q, B, F, forced)
        reqs[40].want_logits = F
        _, lg = eng.generate_codes_batch(reqs)
    finally:
        eng.close()
    rec = {}
    otalker.TalkerOracle(d).generate_codes(otalker.Request(text_ids=reqs[40].text_ids, speaker_id=reqs[40].speaker_id, temperature=0.0, max_tokens=F), forced=forced[40],
                              record=rec, filter_invalid=False)
    e0 = np.abs(lg["code0_logits"] - rec["code0_logits"]).max()
    ec = np.abs(lg["cp_logits"] - rec["cp_logits"]).max()
    rms0, rmsc = float(rec["code0_logits"].std()), float(rec["cp_logits"].std())
    print(f"[0.6b 4-bit stress init] row 40 of 64: code0 {e0:.3e} (rms {rms0:.2f}), code predictor {ec:.3e} (rms {rmsc:.2f})")
    assert e0 <= 5e-3 * rms0 and ec <= 5e-3 * rmsc


# ------------------------------------------------------------------------------------------------ codec, full dimensions
def _snr_db(got, ref):
    got, ref = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    return 10 * np.log10(np.sum(ref ** 2) / max(np.sum((got - ref) ** 2), 1e-300))


@pytest.mark.parametrize("T", [26, 110, 750])
def test_codec_full_dims_snr(T, oracles):
    """BASELINE config 4 decoder (SpeechTokenizer.swift:42-74 defaults) on the tcgen05 pipeline with the fp16 vocoder residual
    stream (the default): whole-sequence decode of T frames vs the fp32 CPU restatement."""
    import qwen3tts_b200 as q

    d = ckpt("codecfull", 8)
    codes = np.random.default_rng(3 + T).integers(0, 2048, size=(1, T, 16)).astype(np.int32)
    eng = q.Engine(d, load_talker=False, codec_max_frames=max(T, 256))
    try:
        got = eng.decode(codes)
    finally:
        eng.close()
    want = oracles(d, "codec").decode(torch.as_tensor(codes).transpose(1, 2).contiguous()).reshape(1, -1).numpy()
    s = _snr_db(got, want)
    print(f"[codec full dims] T={T}: SNR {s:.1f} dB, max|err| {np.abs(got - want).max():.2e}")
    assert got.shape == want.shape and s >= 40.0
