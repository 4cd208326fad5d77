"""CPU self-checks of the ECAPA-TDNN speaker-encoder restatement (oracle/speaker_encoder.py) against independent formulations:
reflect padding vs torch's, the STFT front end vs torch.stft, the Slaney filterbank's published properties, attentive statistics
pooling vs explicit loops, the Res2Net chunk recurrence, and the checkpoint keys `SpeakerEncoder.load` reads."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import ckpt
from oracle import speaker_encoder as se


@pytest.mark.parametrize("n,pad", [(10, 3), (600, 512), (5, 4), (7, 0), (1025, 512)])
def test_reflect_indices_equal_torch_reflect_pad(n, pad):
    """reflectPadSignal / reflectPad1d (SpeakerEncoder.swift:148-167, 213-232) are torch's `reflect` mode for pad < n."""
    x = torch.arange(n, dtype=torch.float32)
    want = F.pad(x[None, None], (pad, pad), mode="reflect")[0, 0] if pad else x
    assert torch.equal(x[torch.from_numpy(se.reflect_indices(n, pad))], want)


def test_mel_front_end_matches_torch_stft():
    """speakerEncoderSTFT (:169-209) = centred, reflect-padded STFT with a symmetric (periodic=False) Hann window."""
    a = (np.random.default_rng(0).standard_normal(24000 + 77) * 0.1).astype(np.float32)
    got = se.mel_spectrogram(a)
    assert got.shape == (a.size // 256 + 1, 128) and got.dtype == np.float32
    spec = torch.stft(torch.from_numpy(a), 1024, 256, 1024, torch.hann_window(1024, periodic=False), center=True, pad_mode="reflect", return_complex=True)
    want = torch.log(torch.clamp(spec.abs().T @ torch.from_numpy(se.mel_filterbank()), min=1e-5)).numpy()
    assert np.abs(got - want).max() < 2e-4


def test_slaney_filterbank_properties():
    """createMelFilterbankImpl (:75-146): triangles on the Slaney scale (linear below 1 kHz, log above), area-normalised."""
    fb = se.mel_filterbank()
    assert fb.shape == (513, 128) and fb.dtype == np.float32 and (fb >= 0).all()
    peaks = fb.argmax(0)
    assert (np.diff(peaks) >= 0).all() and peaks[0] >= 1 and peaks[-1] <= 512
    freqs = np.arange(513) * 12000.0 / 512
    centre = freqs[peaks]
    low = centre < 900
    assert np.allclose(np.diff(centre[low]), np.diff(centre[low])[0], atol=24.0)   # linear spacing (one bin = 23.4 Hz)
    hi = centre > 1500
    ratio = centre[hi][1:] / centre[hi][:-1]
    assert ratio.std() < 0.01                                                      # geometric spacing
    # Slaney normalisation: every triangle integrates to ~1 over frequency
    area = fb.sum(0) * (12000.0 / 512)
    assert np.allclose(area[8:], 1.0, atol=0.12)
    if True:  # librosa's published corner case: hz_to_mel(1000) = 15, the filterbank spans 0 .. fmax
        assert fb[0].sum() == 0.0 and fb[-1].sum() == 0.0


def test_forward_against_explicit_loops():
    d = ckpt("tiny", 8, speaker_encoder="tiny")
    orc = se.SpeakerEncoderOracle(d)
    assert orc.present
    a = (np.random.default_rng(1).standard_normal(24000) * 0.1).astype(np.float32)
    rec = {}
    emb = orc.extract_embedding(a, rec)
    assert emb.shape == (256,) and np.isfinite(emb).all() and np.abs(emb).max() > 1e-3
    # attentive statistics pooling (:366-396) by explicit loops over channels, in float64
    h = torch.from_numpy(rec["mfa"]).double().T  # [C, T]
    C, T = h.shape
    w = {k: v.double() for k, v in orc.w.items()}
    mean, std = h.mean(1), torch.sqrt(((h - h.mean(1, keepdim=True)) ** 2).mean(1) + 1e-12)
    att_in = torch.cat([h, mean[:, None].expand(C, T), std[:, None].expand(C, T)], 0)
    a1 = torch.tanh(torch.relu(w["asp.tdnn.conv.weight"][:, :, 0] @ att_in + w["asp.tdnn.conv.bias"][:, None]))
    a2 = w["asp.conv.weight"][:, :, 0] @ a1 + w["asp.conv.bias"][:, None]
    pooled = torch.zeros(2 * C, dtype=torch.float64)
    for c in range(C):
        p = torch.softmax(a2[c], 0)
        m = (p * h[c]).sum()
        pooled[c] = m
        pooled[C + c] = torch.sqrt(torch.clamp((p * (h[c] - m) ** 2).sum(), min=1e-12))
    want = (w["fc.weight"][:, :, 0] @ pooled + w["fc.bias"]).numpy()
    assert np.abs(want - emb).max() < 1e-4


def test_res2net_recurrence_and_tdnn_padding():
    """TimeDelayNetBlock keeps the length for odd kernels at any dilation; chunk 0 of Res2Net passes through untouched."""
    d = ckpt("tiny", 8, speaker_encoder="tiny")
    orc = se.SpeakerEncoderOracle(d)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(64, 37, generator=g)
    for i, (k, dil) in enumerate(zip(se.KERNELS[1:4], se.DILATIONS[1:4]), start=1):
        y = orc._se_res2net(f"blocks.{i}", x, k, dil)
        assert y.shape == x.shape
    c = torch.randn(8, 37, generator=g)
    y = orc._tdnn("blocks.1.res2net_block.blocks.0", c, 3, 2)
    pad = F.pad(c[None], (2, 2), mode="reflect")
    want = torch.relu(F.conv1d(pad, orc.w["blocks.1.res2net_block.blocks.0.conv.weight"], orc.w["blocks.1.res2net_block.blocks.0.conv.bias"], dilation=2))[0]
    assert torch.allclose(y, want, atol=1e-6)


def test_checkpoint_keys_are_the_ones_the_reference_loader_reads():
    """SpeakerEncoder.load (:550-603) strips `speaker_encoder.` and applies `blocks.N.…`, `mfa.conv`, `asp.tdnn.conv`, `asp.conv`, `fc`."""
    from oracle import checkpoint

    t = checkpoint.speaker_encoder_tensors("full", 0)
    keys = {k[len("speaker_encoder."):] for k in t}
    need = {"blocks.0.conv.weight", "blocks.1.tdnn1.conv.weight", "blocks.3.res2net_block.blocks.6.conv.bias", "blocks.2.se_block.conv2.weight", "mfa.conv.weight",
            "asp.tdnn.conv.weight", "asp.conv.bias", "fc.weight", "fc.bias"}
    assert need <= keys
    assert tuple(t["speaker_encoder.blocks.0.conv.weight"].shape) == (512, 128, 5)
    assert tuple(t["speaker_encoder.mfa.conv.weight"].shape) == (1536, 1536, 1)
    assert tuple(t["speaker_encoder.asp.tdnn.conv.weight"].shape) == (128, 4608, 1)
    assert tuple(t["speaker_encoder.fc.weight"].shape) == (1024, 3072, 1)


def test_filterbank_against_an_independent_float64_slaney_construction():
    """The published Slaney construction (librosa.filters.mel, htk=False, norm='slaney'), written independently in float64 with array
    operations, against the oracle's scalar Float restatement of createMelFilterbankImpl (:75-146)."""
    sr, n_fft, n_mels, fmin, fmax = 24000, 1024, 128, 0.0, 12000.0

    def hz_to_mel(f):
        f = np.asarray(f, dtype=np.float64)
        f_sp = 200.0 / 3
        mels = f / f_sp
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return np.where(f >= min_log_hz, min_log_mel + np.log(np.maximum(f, 1e-12) / min_log_hz) / logstep, mels)

    def mel_to_hz(m):
        m = np.asarray(m, dtype=np.float64)
        f_sp = 200.0 / 3
        min_log_hz, logstep = 1000.0, np.log(6.4) / 27.0
        min_log_mel = min_log_hz / f_sp
        return np.where(m >= min_log_mel, min_log_hz * np.exp(logstep * (m - min_log_mel)), f_sp * m)

    fft_freqs = np.linspace(0, sr / 2, n_fft // 2 + 1)
    mel_f = mel_to_hz(np.linspace(hz_to_mel(fmin), hz_to_mel(fmax), n_mels + 2))
    fdiff = np.diff(mel_f)
    ramps = mel_f[:, None] - fft_freqs[None, :]
    lower = -ramps[:-2] / fdiff[:-1, None]
    upper = ramps[2:] / fdiff[1:, None]
    weights = np.maximum(0, np.minimum(lower, upper)) * (2.0 / (mel_f[2:] - mel_f[:-2]))[:, None]  # [n_mels, n_freqs]
    got = se.mel_filterbank()
    assert got.shape == weights.T.shape
    assert np.abs(got - weights.T).max() < 2e-6 * max(1.0, np.abs(weights).max())


def test_forward_against_torch_nn_modules_with_native_reflect_padding():
    """The whole ECAPA-TDNN graph rebuilt from torch.nn layers (Conv1d with padding_mode='reflect' -- torch's own reflect padding and
    dilation handling, nn.functional statistics) against the oracle's index-based restatement, float64."""
    import torch.nn as nn

    d = ckpt("tiny", 8, speaker_encoder="tiny")
    orc = se.SpeakerEncoderOracle(d)
    w = {k: v.double() for k, v in orc.w.items()}

    def tdnn(prefix, k, dil):
        wt = w[prefix + ".conv.weight"]
        c = nn.Conv1d(wt.shape[1], wt.shape[0], k, dilation=dil, padding=(k - 1) * dil // 2, padding_mode="reflect" if k > 1 else "zeros").double()
        with torch.no_grad():
            c.weight.copy_(wt); c.bias.copy_(w[prefix + ".conv.bias"])
        return nn.Sequential(c, nn.ReLU())

    def conv1(prefix):
        wt = w[prefix + ".weight"]
        c = nn.Conv1d(wt.shape[1], wt.shape[0], 1).double()
        with torch.no_grad():
            c.weight.copy_(wt); c.bias.copy_(w[prefix + ".bias"])
        return c

    a = (np.random.default_rng(2).standard_normal(9000) * 0.1).astype(np.float32)
    mels = se.mel_spectrogram(a)
    x = torch.from_numpy(mels).double().T[None]  # [1, mel, T]
    with torch.no_grad():
        h = tdnn("blocks.0", se.KERNELS[0], se.DILATIONS[0])(x)
        outs = []
        for i in (1, 2, 3):
            p = f"blocks.{i}"
            r = h
            y = tdnn(p + ".tdnn1", 1, 1)(h)
            chunks = torch.chunk(y, se.SCALE, dim=1)
            parts = [chunks[0]]
            for j in range(1, se.SCALE):
                inp = chunks[j] if j == 1 else chunks[j] + parts[-1]
                parts.append(tdnn(f"{p}.res2net_block.blocks.{j - 1}", se.KERNELS[i], se.DILATIONS[i])(inp))
            y = tdnn(p + ".tdnn2", 1, 1)(torch.cat(parts, 1))
            s = torch.sigmoid(conv1(p + ".se_block.conv2")(torch.relu(conv1(p + ".se_block.conv1")(y.mean(2, keepdim=True)))))
            h = y * s + r
            outs.append(h)
        m = tdnn("mfa", se.KERNELS[4], se.DILATIONS[4])(torch.cat(outs, 1))
        T = m.shape[2]
        mean = m.mean(2, keepdim=True)
        std = torch.sqrt(m.var(2, keepdim=True, unbiased=False) + se.EPS)
        att = torch.tanh(tdnn("asp.tdnn", 1, 1)(torch.cat([m, mean.expand(-1, -1, T), std.expand(-1, -1, T)], 1)))
        att = torch.softmax(conv1("asp.conv")(att), dim=2)
        wm = (att * m).sum(2, keepdim=True)
        ws = torch.sqrt(torch.clamp((att * (m - wm) ** 2).sum(2, keepdim=True), min=se.EPS))
        want = conv1("fc")(torch.cat([wm, ws], 1))[0, :, 0].numpy()
    got = orc.forward(mels)
    assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())


def test_forward_equals_hf_ecapa_time_delay_net():
    """Third-party anchor: HuggingFace transformers ships the same ECAPA-TDNN (Qwen2.5-Omni's speaker encoder, `ECAPA_TimeDelayNet` with
    TimeDelayNetBlock / Res2NetBlock / SqueezeExcitationRes2NetBlock / AttentiveStatisticsPooling -- the very class names the reference
    restates in SpeakerEncoder/SpeakerEncoder.swift:234-524).  Same `speaker_encoder.*` weights in, same embedding out."""
    pytest.importorskip("transformers")
    hq = pytest.importorskip("transformers.models.qwen2_5_omni.modeling_qwen2_5_omni")
    from transformers.models.qwen2_5_omni.configuration_qwen2_5_omni import Qwen2_5OmniDiTConfig

    d = ckpt("tiny", 8, speaker_encoder="tiny")
    orc = se.SpeakerEncoderOracle(d)
    ch, mfa = orc.w["blocks.0.conv.weight"].shape[0], orc.w["mfa.conv.weight"].shape[0]
    cfg = Qwen2_5OmniDiTConfig(mel_dim=128, enc_dim=orc.w["fc.weight"].shape[0], enc_channels=[ch, ch, ch, ch, mfa], enc_kernel_sizes=list(se.KERNELS),
                               enc_dilations=list(se.DILATIONS), enc_res2net_scale=se.SCALE, enc_se_channels=orc.w["blocks.1.se_block.conv1.weight"].shape[0],
                               enc_attention_channels=orc.w["asp.tdnn.conv.weight"].shape[0])
    m = hq.ECAPA_TimeDelayNet(cfg).to(torch.float32).eval()
    sd = m.state_dict()
    with torch.no_grad():
        for k in sd:
            assert k in orc.w, f"no oracle tensor for {k}"
            sd[k].copy_(orc.w[k])
    for L in (1024, 5000, 24000):
        a = (np.random.default_rng(L).standard_normal(L) * 0.1).astype(np.float32)
        mels = se.mel_spectrogram(a)
        with torch.no_grad():
            want = m(torch.from_numpy(mels)[None])[0].numpy()
        got = orc.forward(mels)
        assert np.abs(got - want).max() <= 1e-5 * max(1.0, np.abs(want).max()), L
