"""GPU parity: ICL reference-audio encoder (csrc/audio_encoder.cu) through the C ABI vs the CPU oracle (oracle/audio_encoder.py).

The output is a table of nearest-neighbour indices, so the bar is: quantiser input (latent) within 1e-3 of the oracle's, and code ids
IDENTICAL wherever the oracle's distance margin (second-best minus best) exceeds 1e-3; a residual chain that took a different
codeword at a near-tie is compared no further (its later layers quantise a different residual)."""
import numpy as np
import pytest

from conftest import ckpt

pytestmark = pytest.mark.gpu

MARGIN = 1e-3


def _check(eng, orc, audio):
    rec = {}
    want = orc.encode(audio, rec)[0]            # [Q, T]
    got, lat = eng.encode_reference_audio(audio, want_latent=True)
    assert got.shape == want.shape and got.dtype == np.int32
    err = np.abs(lat - rec["latent"][0]).max()
    scale = np.abs(rec["latent"][0]).max()
    assert err <= 1e-3 * max(1.0, scale), f"latent max-abs error {err:.3e} (scale {scale:.2f})"
    Q, T = want.shape
    n_sem = orc.cfg.num_semantic_quantizers
    same = flips = 0
    for t in range(T):
        for chain in (range(0, n_sem), range(n_sem, Q)):
            for q in chain:
                if got[q, t] != want[q, t]:
                    assert rec["margins"][0, q, t] < MARGIN, f"frame {t} quantiser {q}: ids differ at margin {rec['margins'][0, q, t]:.4f}"
                    flips += 1
                    break  # the rest of this chain quantises a different residual
                same += 1
    return same, flips, err


@pytest.mark.parametrize("L", [960, 1919, 1921, 24000, 24000 * 3 + 517])
def test_encoder_codes_tiny(L, engines):
    from oracle import audio_encoder as ae

    d = ckpt("tiny", 8, encoder="tiny")
    eng = engines(d, load_talker=False)
    assert eng.info.has_audio_encoder == 1
    orc = ae.AudioEncoderOracle(d + "/speech_tokenizer")
    audio = (np.random.default_rng(L).standard_normal(L) * 0.1).astype(np.float32)
    same, flips, err = _check(eng, orc, audio)
    print(f"[tiny encoder] L={L}: {same} code ids identical, {flips} near-tie chain flips, latent max-abs error {err:.2e}")


@pytest.mark.slow
def test_encoder_codes_full_dims(engines):
    """Qwen3TTSTokenizerEncoderConfig defaults (64 filters, ratios 8/6/5/4, 8 x 512 transformer, 32 x 2048 x 256 codebooks), 6 s of audio."""
    import qwen3tts_b200 as q
    from oracle import audio_encoder as ae

    d = ckpt("tiny", 8, encoder="full")
    eng = q.Engine(d, load_talker=False)
    try:
        orc = ae.AudioEncoderOracle(d + "/speech_tokenizer")
        t = np.arange(24000 * 6) / 24000.0
        audio = (0.2 * np.sin(2 * np.pi * 220 * t) * np.sin(2 * np.pi * 3 * t) + 0.05 * np.random.default_rng(1).standard_normal(t.size)).astype(np.float32)
        same, flips, err = _check(eng, orc, audio)
        print(f"[full encoder] 6 s: {same} code ids identical, {flips} near-tie chain flips, latent max-abs error {err:.2e}; encode {eng.timing().device_ms:.2f} ms on the device")
        assert same >= 0.9 * 16 * 75
    finally:
        eng.close()


def test_no_encoder_weights_means_no_icl(tiny8, engines):
    """Like the reference without `encoder.*` tensors: supportsICL is false and encodeReferenceAudio returns nil."""
    eng = engines(tiny8)
    assert eng.info.has_audio_encoder == 0
    assert eng.encode_reference_audio(np.zeros(24000, np.float32)) is None


def test_icl_round_trip_through_the_pipeline(engines):
    """encodeReferenceAudio -> referenceAudioCodes of generate_to_file (BASELINE config 5's data path): the codes feed
    codec_embedding(refCodes[0]) rows of the prompt (Model/Qwen3Talker.swift:395-404)."""
    import qwen3tts_b200 as q
    from oracle import talker as otalker

    d = ckpt("tiny", 8, encoder="tiny")
    p = q.Qwen3TTSPipeline(d, q.Qwen3TTSPipelineConfiguration(default_max_tokens=40))
    try:
        assert p.supports_icl
        audio = (np.random.default_rng(3).standard_normal(24000) * 0.1).astype(np.float32)
        codes = p.encode_reference_audio(audio)
        assert codes.shape == (16, 13) and codes.dtype == np.int32
        req = p._request("Hello there, this is a cloned voice.", reference_transcript="the reference text", reference_audio_codes=codes, temperature=0.0, max_tokens=12)
        got = p.engine.generate_codes(q.GenRequest(**{**req.__dict__, "keep_invalid_frames": True}))
        rec = {}
        want = otalker.TalkerOracle(d).generate_codes(otalker.Request(text_ids=req.text_ids, ref_text_ids=req.ref_text_ids, ref_codes=codes, temperature=0.0,
                                                                      max_tokens=12), record=rec, filter_invalid=False)
        n = min(len(got), len(want))
        for f in range(n):
            if got[f].tolist() != want[f]:
                g = next(k for k in range(16) if got[f][k] != want[f][k])
                assert rec["margins"][f][g] < 2e-2
                break
    finally:
        p.close()
