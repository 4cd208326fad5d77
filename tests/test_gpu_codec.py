"""GPU parity: speech-tokenizer decoder through the C ABI vs the CPU oracle.

Contracts (BASELINE.json north_star): RVQ code->embedding lookup bit-exact; PCM SNR >= 40 dB (measured far above on
the fp32 path; the bar is written here)."""
import numpy as np
import pytest
import torch

from conftest import TEXT_IDS, ckpt

pytestmark = pytest.mark.gpu

SNR_DB = 40.0


def snr_db(got, ref):
    got, ref = np.asarray(got, np.float64).ravel(), np.asarray(ref, np.float64).ravel()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    err = np.sum((got - ref) ** 2)
    return 10 * np.log10(np.sum(ref ** 2) / max(err, 1e-300))


def rand_codes(B, T, seed):
    return np.random.default_rng(seed).integers(0, 2048, size=(B, T, 16)).astype(np.int32)


def oracle_decode(codec, codes):  # codes [B,T,16] -> [B, T*1920]
    t = torch.as_tensor(codes).transpose(1, 2).contiguous()
    return codec.decode(t).reshape(codes.shape[0], -1).numpy()


def test_rvq_embed_bit_exact(tiny8, engines, oracles):
    codes = rand_codes(3, 11, 0)
    codes[0, 0, :] = 13  # a dead codebook entry (cluster_usage clipped to 1e-5)
    first, rest = engines(tiny8).rvq_embed(codes)
    of, orr = oracles(tiny8, "codec").rvq_embed(torch.as_tensor(codes).transpose(1, 2))
    assert np.array_equal(first.view(np.uint32), of.numpy().view(np.uint32))
    assert np.array_equal(rest.view(np.uint32), orr.numpy().view(np.uint32))


@pytest.mark.parametrize("B,T", [(1, 1), (1, 2), (1, 7), (2, 18), (1, 26), (3, 33), (1, 110)])
def test_decode_matches_oracle(B, T, tiny8, engines, oracles):
    codes = rand_codes(B, T, B * 100 + T)
    got = engines(tiny8).decode(codes)
    want = oracle_decode(oracles(tiny8, "codec"), codes)
    s = snr_db(got, want)
    print(f"decode B={B} T={T}: SNR {s:.1f} dB, max|err| {np.abs(got - want).max():.2e}")
    assert s >= SNR_DB
    assert np.all(np.abs(got) <= 1.0)


def test_decode_reference_init_weights(engines, oracles):
    """alpha = beta = 0, LayerScale 0.01, gamma 1e-6 — the reference's own init values (SpeechTokenizer.swift:100-101, 220, 264)."""
    d = ckpt("tiny", 8, visible=False)
    codes = rand_codes(2, 9, 5)
    s = snr_db(engines(d).decode(codes), oracle_decode(oracles(d, "codec"), codes))
    assert s >= SNR_DB


@pytest.mark.parametrize("B,T,chunk,left", [(1, 25, 10, 3), (2, 25, 10, 3), (3, 40, 10, 2), (1, 7, 100, 10), (2, 30, 10, 0)])
def test_chunked_decode(B, T, chunk, left, tiny8, engines, oracles):
    codes = rand_codes(B, T, T + chunk)
    got = engines(tiny8).decode_chunked(codes, chunk, left)
    want = oracles(tiny8, "codec").chunked_decode(torch.as_tensor(codes).transpose(1, 2).contiguous(), chunk, left).reshape(B, -1).numpy()
    assert got.shape == want.shape
    assert snr_db(got, want) >= SNR_DB


@pytest.mark.parametrize("mode,chunk", [("whole", 0), ("file", 16), ("batchapi", 24), ("stream", 18)])
def test_generate_pcm_modes(mode, chunk, tiny8, engines, oracles):
    import qwen3tts_b200 as q
    from oracle import pipeline as opipe, talker as otalker

    eng = engines(tiny8)
    steps = 60
    frames = oracles(tiny8).generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=steps))
    codec = oracles(tiny8, "codec")
    if mode == "whole":
        want = opipe.decode_whole(codec, frames)
    else:
        want = np.concatenate([w for w, _ in opipe.decode_windowed(codec, frames, chunk, 8)])
    m = {"whole": q.DECODE_WHOLE, "file": q.DECODE_FILE, "batchapi": q.DECODE_BATCHAPI, "stream": q.DECODE_STREAM}[mode]
    pcm, n = eng.generate_pcm(q.GenRequest(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=steps), m)
    if n != len(frames):
        pytest.skip("greedy ids diverged at a near-tie; covered by test_generate_codes_greedy")
    assert pcm.shape == want.shape
    assert snr_db(pcm, want) >= SNR_DB


def test_stream_audio_chunks(tiny8, engines, oracles):
    """_generateStreamImpl: windows 18 then 8+18, flush with isFinal, trailing empty isFinal chunk (quirk 9)."""
    import qwen3tts_b200 as q
    from oracle import pipeline as opipe, talker as otalker

    steps = 50
    raw = oracles(tiny8).generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=steps, stream_variant=True),
                                        filter_invalid=False)
    want = opipe.stream_chunks(oracles(tiny8, "codec"), [raw[i:i + 12] for i in range(0, len(raw), 12)])
    st = engines(tiny8).stream(q.GenRequest(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=steps), 12)
    got = []
    while True:
        s, rng, fin, done = st.next_audio()
        got.append((s, rng, fin))
        if done:
            break
    st.close()
    assert [(g[1], g[2], g[0].size) for g in got] == [(w["token_range"], w["is_final"], w["samples"].size) for w in want]
    for g, w in zip(got, want):
        if w["samples"].size:
            assert snr_db(g[0], w["samples"]) >= SNR_DB
    assert got[-1][0].size == 0 and got[-1][2]


def test_stream_code_chunks(tiny8, engines, oracles):
    import qwen3tts_b200 as q
    from oracle import talker as otalker

    raw = oracles(tiny8).generate_codes(otalker.Request(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=29, stream_variant=True),
                                        filter_invalid=False)
    st = engines(tiny8).stream(q.GenRequest(text_ids=TEXT_IDS, speaker_id=2861, temperature=0.0, max_tokens=29), 12)
    chunks = []
    while True:
        c, done = st.next_codes()
        if len(c):
            chunks.append(c.tolist())
        if done:
            break
    st.close()
    assert [len(c) for c in chunks] == [12, 12, 5]  # chunkSize groups + final partial (Qwen3Talker.swift:831-835, 871-873)
    assert [f for c in chunks for f in c] == raw


def test_pipeline_api(tiny8, tmp_path):
    """The reference-shaped surface: generate / generate_stream / generate_to_file / errors."""
    import qwen3tts_b200 as q

    p = q.Qwen3TTSPipeline(tiny8, q.Qwen3TTSPipelineConfiguration(default_max_tokens=40))
    assert p.available_speakers == sorted(["serena", "vivian", "uncle_fu", "ryan", "aiden", "ono_anna", "sohee", "eric", "dylan"])
    assert q.Qwen3TTSPipeline.sample_rate == 24000 and p.model_type is None
    a = p.generate("Hello world, this is a test.", speaker="Aiden", temperature=0.0, max_tokens=20)
    assert a.dtype == np.float32 and a.size % 1920 == 0 and a.size > 0 and np.all(np.abs(a) <= 1.0)
    chunks = list(p.generate_stream("Hello world, this is a test.", speaker="aiden", temperature=0.0, max_tokens=30))
    assert chunks[-1].is_final and chunks[-1].samples.size == 0
    n = p.generate_to_file("Hello there. " * 30, tmp_path / "o.wav", speaker="aiden", temperature=0.0)
    data = (tmp_path / "o.wav").read_bytes()
    assert data[:4] == b"RIFF" and len(data) == 44 + 2 * n
    assert p.generate("Hi", speaker="aiden").size == 0 or True  # short prompts still carry the 9-token template
    with pytest.raises(q.FileNotFound):
        q.Qwen3TTSPipeline(str(tmp_path / "missing"))
    p.clear_cache()
    p.close()


# ------------------------------------------------------------------------------------------------ tcgen05 pipeline
@pytest.mark.parametrize("preset,cases", [("tcsmall", [(1, 1), (1, 7), (2, 18), (1, 26), (3, 33), (1, 140)]), ("codecfull", [(1, 3), (2, 5)])])
def test_decode_tensor_core_path(preset, cases, engines, oracles):
    """Same parity bar on checkpoints whose codec runs the tcgen05/TMEM pipeline (fp16 operands, fp32 accumulate/residuals)."""
    d = ckpt(preset, 8)
    eng, codec = engines(d), oracles(d, "codec")
    for B, T in cases:
        codes = rand_codes(B, T, B * 31 + T)
        got = eng.decode(codes)
        want = oracle_decode(codec, codes)
        s = snr_db(got, want)
        print(f"[{preset}] decode B={B} T={T}: SNR {s:.1f} dB, max|err| {np.abs(got - want).max():.2e}")
        assert s >= SNR_DB


def test_tensor_core_vs_simt_pipeline(engines, monkeypatch):
    """A/B: the two codec pipelines of the engine agree with each other far inside the parity bar."""
    import qwen3tts_b200 as q

    d = ckpt("tcsmall", 8)
    codes = rand_codes(2, 21, 9)
    a = engines(d).decode(codes)
    monkeypatch.setenv("Q3TTS_CODEC_SIMT", "1")
    e2 = q.Engine(d, max_frames=64, load_talker=False)
    b = e2.decode(codes)
    e2.close()
    assert snr_db(a, b) >= 45.0


@pytest.mark.parametrize("preset,B,T", [("tcsmall", 1, 140), ("tcsmall", 3, 70), ("codecfull", 2, 16)])
def test_fused_residual_unit_vs_two_kernel_form(preset, B, T, oracles, monkeypatch):
    """DecoderResidualUnit as one persistent kernel (csrc/codec_unit.cu: conv7 -> SnakeBeta -> 1x1 with the intermediate in tensor
    memory) against the two-launch form of the same unit: same parity bar against the oracle, and the two agree far inside it."""
    import qwen3tts_b200 as q

    d = ckpt(preset, 8)
    codes = rand_codes(B, T, 77 + T)
    want = oracle_decode(oracles(d, "codec"), codes)
    out = {}
    for unit in ("1", "0"):
        monkeypatch.setenv("Q3TTS_CODEC_UNIT", unit)
        e = q.Engine(d, max_frames=256, load_talker=False)
        out[unit] = e.decode(codes)
        e.close()
    s1, s0, sab = snr_db(out["1"], want), snr_db(out["0"], want), snr_db(out["1"], out["0"])
    print(f"[{preset} B={B} T={T}] fused unit {s1:.1f} dB, two-kernel {s0:.1f} dB vs oracle; fused vs two-kernel {sab:.1f} dB")
    assert s1 >= SNR_DB and s0 >= SNR_DB and sab >= 45.0


def test_chunked_decode_tensor_core(engines, oracles):
    d = ckpt("tcsmall", 8)
    codes = rand_codes(2, 45, 3)
    got = engines(d).decode_chunked(codes, 20, 4)
    want = oracles(d, "codec").chunked_decode(torch.as_tensor(codes).transpose(1, 2).contiguous(), 20, 4).reshape(2, -1).numpy()
    assert snr_db(got, want) >= SNR_DB


# ------------------------------------------------------------------------------------------------ long-text drivers (§8 f3)
LONG_TEXT = ("The quick brown fox jumps over the lazy dog near the quiet river bank. " * 9 + "Then it rests. ") * 2


def _oracle_file_pcm(p, oracles, d, text, steps, mode_chunk):
    """What generateToFile / generateBatch decode per text chunk: greedy codes of the oracle, windows of `mode_chunk` + 8."""
    from oracle import pipeline as opipe, talker as otalker
    import qwen3tts_b200 as q

    outs, margins = [], []
    for tc in q.TextChunker.chunk(text, q.TextChunker.default_max_words):
        r = p._request(tc, speaker="aiden", temperature=0.0, max_tokens=steps)
        rec = {}
        raw = oracles(d).generate_codes(otalker.Request(text_ids=r.text_ids, speaker_id=r.speaker_id, temperature=0.0, max_tokens=steps), record=rec,
                                        filter_invalid=False)
        frames = opipe.valid_frames(raw)
        pcm = np.concatenate([w for w, _ in opipe.decode_windowed(oracles(d, "codec"), frames, mode_chunk, 8)]) if frames else np.zeros(0, np.float32)
        outs.append((raw, pcm))
        margins.append(rec["margins"])
    return outs, margins


def test_generate_to_file_bytes_vs_oracle(tiny8, oracles, tmp_path):
    """generateToFile (Qwen3TTSPipeline.swift:644-757): TextChunker -> per-chunk generateCodes (maxTokens 600; shortened here through
    the configuration) -> 16 + 8 windows -> StreamingWAVWriter int16 PCM.  Header and sample count exact, int16 payload >= 40 dB
    against the oracle's; a 4-slot handle (chunks generated side by side) writes the SAME BYTES as a 1-slot handle of the same
    numeric path would -- compared here against its own sequential run."""
    import struct

    import qwen3tts_b200 as q

    steps = 40
    p = q.Qwen3TTSPipeline(tiny8, q.Qwen3TTSPipelineConfiguration(max_batch=4))
    try:
        chunks = q.TextChunker.chunk(LONG_TEXT, q.TextChunker.default_max_words)
        assert len(chunks) >= 3
        # the driver hard-codes maxTokens 600 (:690); keep the test short by capping frames through the request factory
        orig = p._request
        p._request = lambda *a, **kw: orig(*a, **{**kw, "max_tokens": steps, "temperature": 0.0})
        n = p.generate_to_file(LONG_TEXT, tmp_path / "par.wav", speaker="aiden", temperature=0.0)
        par = (tmp_path / "par.wav").read_bytes()
        # the same handle, one chunk at a time
        seq_pcm = []
        for tc in chunks:
            pcm, fr = p.engine.generate_pcm(p._request(tc, speaker="aiden"), q.DECODE_FILE)
            seq_pcm.append(pcm)
        w = q.StreamingWAVWriter(tmp_path / "seq.wav")
        for pcm in seq_pcm:
            if pcm.size:
                w.write(pcm)
        assert w.finalize() == n
        assert par == (tmp_path / "seq.wav").read_bytes(), "chunk-parallel generation changed the file"
        # header (AudioSampleWriter.swift:44-106): 44 bytes, PCM16 mono 24 kHz
        assert par[:4] == b"RIFF" and par[8:16] == b"WAVEfmt " and struct.unpack("<IHHIIHH", par[16:36]) == (16, 1, 1, 24000, 48000, 2, 16)
        assert struct.unpack("<I", par[40:44])[0] == 2 * n and len(par) == 44 + 2 * n
        want, margins = _oracle_file_pcm(p, oracles, tiny8, LONG_TEXT, steps, 16)
        got16 = np.frombuffer(par[44:], dtype="<i2").astype(np.float64)
        off = 0
        compared = 0
        for (raw, pcm), tc in zip(want, chunks):
            mine = p.engine.generate_codes(q.GenRequest(**{**p._request(tc, speaker="aiden").__dict__, "keep_invalid_frames": True})).tolist()
            k = pcm.size
            if mine == raw:  # greedy ids identical: the PCM of this chunk is comparable
                ref16 = np.trunc(np.clip(pcm, -1, 1) * 32767.0)  # Int16(clamped * 32767) truncation (AudioSampleWriter.swift:93-96)
                seg = got16[off: off + k]
                err = np.sum((seg - ref16) ** 2)
                assert 10 * np.log10(np.sum(ref16 ** 2) / max(err, 1e-9)) >= 40.0
                compared += 1
            off += len(opipe_valid(mine)) * 1920
        assert off == n and compared >= 1
    finally:
        p.close()


def opipe_valid(frames):
    from oracle import pipeline as opipe

    return opipe.valid_frames(frames)


def test_generate_batch_crossfade_vs_oracle(tiny8, oracles):
    """generateBatch (Qwen3TTSPipeline.swift:774-898): windows of 24 + 8 per chunk, 480-sample linear crossfade between chunks."""
    import qwen3tts_b200 as q
    from oracle import pipeline as opipe

    steps = 30
    p = q.Qwen3TTSPipeline(tiny8, q.Qwen3TTSPipelineConfiguration())
    try:
        orig = p._request
        p._request = lambda *a, **kw: orig(*a, **{**kw, "max_tokens": steps, "temperature": 0.0})
        got = p.generate_batch(LONG_TEXT, speaker="aiden", temperature=0.0)
        want, _ = _oracle_file_pcm(p, oracles, tiny8, LONG_TEXT, steps, 24)
        chunks = q.TextChunker.chunk(LONG_TEXT, q.TextChunker.default_max_words)
        same = all(p.engine.generate_codes(q.GenRequest(**{**p._request(tc, speaker="aiden").__dict__, "keep_invalid_frames": True})).tolist() == raw
                   for tc, (raw, _) in zip(chunks, want))
        ref = opipe.crossfade_concat([pcm for _, pcm in want], p.config.crossfade_samples)
        if not same:
            pytest.skip("greedy ids diverged at a near-tie in one chunk; covered by test_generate_codes_greedy")
        assert got.shape == ref.shape and snr_db(got, ref) >= SNR_DB
    finally:
        p.close()


def test_generate_batch_chunk_parallel_equals_one_chunk_at_a_time(tiny8):
    """generateBatch's chunks are independent generations (Qwen3TTSPipeline.swift:813-864): a handle with several slots runs them through one
    batched call; the samples must be those of the same handle run chunk by chunk."""
    import qwen3tts_b200 as q
    from oracle import pipeline as opipe

    steps = 24
    p = q.Qwen3TTSPipeline(tiny8, q.Qwen3TTSPipelineConfiguration(max_batch=4))
    try:
        orig = p._request
        p._request = lambda *a, **kw: orig(*a, **{**kw, "max_tokens": steps, "temperature": 0.0})
        got = p.generate_batch(LONG_TEXT, speaker="aiden", temperature=0.0)
        chunks = q.TextChunker.chunk(LONG_TEXT, q.TextChunker.default_max_words)
        singles = [p.engine.generate_pcm(p._request(tc, speaker="aiden"), q.DECODE_BATCHAPI)[0] for tc in chunks]
        ref = opipe.crossfade_concat([s for s in singles if s.size], p.config.crossfade_samples)
        assert got.shape == ref.shape and np.array_equal(got, ref)
    finally:
        p.close()
