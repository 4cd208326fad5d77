"""GPU parity: ECAPA-TDNN speaker encoder (csrc/speaker_encoder.cu) through the C ABI vs the CPU oracle (oracle/speaker_encoder.py).

Floating point end to end (fp32 SIMT kernels): log-mel within 2e-3 absolute of the oracle's rfft-based front end (the log of a
clipped 1e-5 floor amplifies fp32 summation noise in silent bins, so the bar is stated on the log values), embedding within
1e-3 x max(1, |embedding|_inf)."""
import numpy as np
import pytest

from conftest import ckpt

pytestmark = pytest.mark.gpu


def _audio(L, seed):
    t = np.arange(L) / 24000.0
    rng = np.random.default_rng(seed)
    return (0.2 * np.sin(2 * np.pi * 180 * t) * (1 + 0.5 * np.sin(2 * np.pi * 2.5 * t)) + 0.05 * rng.standard_normal(L)).astype(np.float32)


def _check(eng, orc, audio):
    rec = {}
    want = orc.extract_embedding(audio, rec)
    got, mels = eng.extract_speaker_embedding(audio, want_mels=True)
    assert got.shape == want.shape and got.dtype == np.float32
    assert mels.shape == rec["mels"].shape
    mel_err = np.abs(mels - rec["mels"]).max()
    assert mel_err <= 2e-3, f"log-mel max-abs error {mel_err:.3e}"
    err = np.abs(got - want).max()
    scale = max(1.0, np.abs(want).max())
    assert err <= 1e-3 * scale, f"embedding max-abs error {err:.3e} (scale {scale:.2f})"
    return mel_err, err, scale


@pytest.mark.parametrize("L", [1024, 1279, 5000, 24000, 24000 * 3 + 517])
def test_speaker_embedding_tiny(L, engines):
    from oracle import speaker_encoder as se

    d = ckpt("tiny", 8, speaker_encoder="tiny")
    eng = engines(d, load_codec=False)
    assert eng.info.has_speaker_encoder == 1 and eng.info.speaker_embedding_dim == 256
    mel_err, err, scale = _check(eng, se.SpeakerEncoderOracle(d), _audio(L, L))
    print(f"[tiny speaker encoder] L={L}: log-mel max-abs error {mel_err:.2e}, embedding max-abs error {err:.2e} (scale {scale:.2f})")


@pytest.mark.slow
def test_speaker_embedding_full_dims(engines):
    """SpeakerEncoderConfig defaults (512-channel blocks, 1536-channel MFA, 1024-d embedding; SpeakerEncoder.swift:399-418), 5 s of audio."""
    import qwen3tts_b200 as q
    from oracle import speaker_encoder as se

    d = ckpt("tiny", 8, speaker_encoder="full")
    eng = q.Engine(d, load_codec=False)
    try:
        assert eng.info.speaker_embedding_dim == 1024
        mel_err, err, scale = _check(eng, se.SpeakerEncoderOracle(d), _audio(24000 * 5, 7))
        print(f"[full speaker encoder] 5 s: log-mel max-abs error {mel_err:.2e}, embedding max-abs error {err:.2e} (scale {scale:.2f}); "
              f"{eng.timing().device_ms:.2f} ms on the device")
    finally:
        eng.close()


def test_silence_hits_the_log_floor(engines):
    """log(clip(., 1e-5)) (:66-69): all-zero audio gives log(1e-5) in every bin on both sides."""
    d = ckpt("tiny", 8, speaker_encoder="tiny")
    eng = engines(d, load_codec=False)
    _, mels = eng.extract_speaker_embedding(np.zeros(4096, np.float32), want_mels=True)
    assert np.allclose(mels, np.log(np.float32(1e-5)), atol=1e-6)


def test_too_short_audio_is_an_argument_error(engines):
    import qwen3tts_b200 as q

    d = ckpt("tiny", 8, speaker_encoder="tiny")
    eng = engines(d, load_codec=False)
    with pytest.raises(q.Q3Error):
        eng.extract_speaker_embedding(np.zeros(1023, np.float32))
    # the handle stays usable
    assert eng.extract_speaker_embedding(_audio(2048, 0)).shape == (256,)


def test_no_speaker_encoder_weights_means_no_voice_cloning(tiny8, engines):
    """Like the reference without `speaker_encoder.*` tensors: supportsVoiceCloning is false and extractSpeakerEmbedding returns nil."""
    eng = engines(tiny8)
    assert eng.info.has_speaker_encoder == 0
    assert eng.extract_speaker_embedding(np.zeros(24000, np.float32)) is None


def test_voice_cloning_round_trip_through_the_pipeline():
    """extractSpeakerEmbedding -> speakerEmbedding of generate (Qwen3TTSPipeline.swift:279-306): the embedding becomes the speaker row of
    the prompt (Model/Qwen3Talker.swift:374-376); codes equal the oracle's fed the ORACLE's embedding wherever its margin allows."""
    import qwen3tts_b200 as q
    from oracle import speaker_encoder as se, talker as otalker

    d = ckpt("tiny", 8, speaker_encoder="tiny")
    p = q.Qwen3TTSPipeline(d, q.Qwen3TTSPipelineConfiguration(default_max_tokens=40))
    try:
        assert p.supports_voice_cloning
        audio = _audio(24000, 11)
        emb = p.extract_speaker_embedding(audio)
        assert emb.shape == (256,)
        req = p._request("Hello there, this is a cloned voice.", speaker_embedding=emb, temperature=0.0, max_tokens=12)
        got = p.engine.generate_codes(q.GenRequest(**{**req.__dict__, "keep_invalid_frames": True}))
        rec = {}
        want_emb = se.SpeakerEncoderOracle(d).extract_embedding(audio)
        want = otalker.TalkerOracle(d).generate_codes(otalker.Request(text_ids=req.text_ids, speaker_embedding=want_emb, temperature=0.0, max_tokens=12),
                                                      record=rec, filter_invalid=False)
        assert len(got) > 0
        n = min(len(got), len(want))
        for f in range(n):
            if got[f].tolist() != want[f]:
                g = next(k for k in range(16) if got[f][k] != want[f][k])
                assert rec["margins"][f][g] < 2e-2
                break
    finally:
        p.close()
