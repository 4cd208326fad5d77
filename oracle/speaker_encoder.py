"""ECAPA-TDNN speaker encoder: CPU restatement of `SpeakerEncoder` (oracle; test infrastructure only; parity unpinned).

Follows `SpeakerEncoder/SpeakerEncoder.swift` (torch CPU fp32):
  * `melSpectrogram` :37-73         reflect pad n_fft/2, frames of 1024 @ hop 256, SYMMETRIC Hann (i / (N-1)), |rfft|, mel filterbank
                                     [513, 128] (Slaney scale, Slaney norm: `createMelFilterbankImpl` :75-146), log(clip(., 1e-5))
  * `TimeDelayNetBlock` :234-257    reflect pad (k-1)*d/2 both sides, Conv1d, ReLU
  * `Res2NetBlock` :260-302         8 channel chunks: y0 = x0, y1 = tdnn0(x1), y_i = tdnn_{i-1}(x_i + y_{i-1})
  * `SqueezeExcitationBlock` :304-322   x * sigmoid(conv2(relu(conv1(mean_t x))))
  * `SqueezeExcitationRes2NetBlock` :324-353   tdnn1 -> res2net -> tdnn2 -> se, + residual
  * `AttentiveStatisticsPooling` :355-397      attention over time from [x, mean, std]; weighted mean and std
  * `SpeakerEncoder.callAsFunction` :496-524, `extractEmbedding` :526-542, `load` :550-603 (keys `speaker_encoder.*`, conv [out, in, k])

The reference always builds the default configuration (`SpeakerEncoder()`, Qwen3TTSPipeline.swift:159); here every width is read from the
tensor shapes, kernel sizes / dilations / Res2Net scale are the defaults, so that small seeded presets exercise the same graph.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn.functional as F
from safetensors.torch import load_file

N_FFT, HOP, WIN, N_MELS, SR, FMIN, FMAX = 1024, 256, 1024, 128, 24000, 0.0, 12000.0
KERNELS = (5, 3, 3, 3, 1)
DILATIONS = (1, 2, 3, 4, 1)
SCALE = 8
EPS = 1e-12


def mel_filterbank(sample_rate=SR, n_fft=N_FFT, n_mels=N_MELS, fmin=FMIN, fmax=FMAX) -> np.ndarray:
    """[n_fft/2 + 1, n_mels] float32, the Float arithmetic of `createMelFilterbankImpl` (:75-146)."""
    f32 = np.float32
    n_freqs = n_fft // 2 + 1
    f_sp = f32(200.0) / f32(3.0)
    min_log_hz = f32(1000.0)
    min_log_mel = min_log_hz / f_sp
    log_step = f32(math.log(6.4) / 27.0)

    def hz_to_mel(hz):
        hz = f32(hz)
        if hz >= min_log_hz:
            return f32(min_log_mel + f32(math.log(float(hz / min_log_hz))) / log_step)
        return f32(hz / f_sp)

    def mel_to_hz(mel):
        mel = f32(mel)
        if mel >= min_log_mel:
            return f32(min_log_hz * f32(math.exp(float(log_step * (mel - min_log_mel)))))
        return f32(f_sp * mel)

    all_freqs = [f32(i) * f32(sample_rate // 2) / f32(n_freqs - 1) for i in range(n_freqs)]
    m_min, m_max = hz_to_mel(fmin), hz_to_mel(fmax)
    m_pts = [f32(m_min + f32(i) * (m_max - m_min) / f32(n_mels + 1)) for i in range(n_mels + 2)]
    f_pts = [mel_to_hz(m) for m in m_pts]
    f_diff = [f32(f_pts[i + 1] - f_pts[i]) for i in range(len(f_pts) - 1)]
    fb = np.zeros((n_freqs, n_mels), dtype=np.float32)
    for k in range(n_freqs):
        for m in range(n_mels):
            down = f32(all_freqs[k] - f_pts[m]) / f_diff[m]
            up = f32(f_pts[m + 2] - all_freqs[k]) / f_diff[m + 1]
            fb[k, m] = max(f32(0.0), min(down, up))
    for m in range(n_mels):
        fb[:, m] *= f32(2.0) / f32(f_pts[m + 2] - f_pts[m])
    return fb


def reflect_indices(n: int, pad: int) -> np.ndarray:
    """`reflectPadSignal` / `reflectPad1d` (:148-167, :213-232): pad..1, 0..n-1, n-2..max(n-pad-1, 0)."""
    return np.array(list(range(pad, 0, -1)) + list(range(n)) + list(range(n - 2, max(n - pad - 1, 0) - 1, -1)), dtype=np.int64)


def hann_symmetric(n: int) -> np.ndarray:
    i = np.arange(n, dtype=np.float32)
    return (np.float32(0.5) * (np.float32(1.0) - np.cos(np.float32(2.0) * np.float32(np.pi) * i / np.float32(n - 1)))).astype(np.float32)


def mel_spectrogram(audio: np.ndarray) -> np.ndarray:
    """audio float32 [L] -> log-mel [frames, 128] (`melSpectrogram` :37-73, `speakerEncoderSTFT` :169-209)."""
    x = np.asarray(audio, dtype=np.float32)
    padded = x[reflect_indices(x.shape[0], N_FFT // 2)]
    frames = (padded.shape[0] - N_FFT) // HOP + 1
    idx = np.arange(frames)[:, None] * HOP + np.arange(N_FFT)[None, :]
    fr = torch.from_numpy(padded[idx] * hann_symmetric(WIN)[None, :])
    mag = torch.fft.rfft(fr, dim=1).abs().to(torch.float32)
    mel = mag @ torch.from_numpy(mel_filterbank())
    return torch.log(torch.clamp(mel, min=1e-5)).numpy()


class SpeakerEncoderOracle:
    def __init__(self, model_dir: str):
        raw = load_file(os.path.join(model_dir, "model.safetensors"))
        self.w = {k[len("speaker_encoder."):]: v.to(torch.float32) for k, v in raw.items() if k.startswith("speaker_encoder.")}  # :551-555
        self.present = len(self.w) > 0

    def _tdnn(self, prefix: str, x: torch.Tensor, k: int, d: int) -> torch.Tensor:
        """x [C, T] -> relu(conv(reflect_pad(x)))  (`TimeDelayNetBlock`)."""
        pad = (k - 1) * d // 2
        if pad > 0:
            x = x[:, torch.from_numpy(reflect_indices(x.shape[1], pad))]
        y = F.conv1d(x[None], self.w[prefix + ".conv.weight"], self.w[prefix + ".conv.bias"], dilation=d)[0]
        return torch.relu(y)

    def _se_res2net(self, prefix: str, x: torch.Tensor, k: int, d: int) -> torch.Tensor:
        h = self._tdnn(prefix + ".tdnn1", x, 1, 1)
        chunk = h.shape[0] // SCALE
        outs, part = [], None
        for i in range(SCALE):
            c = h[i * chunk:(i + 1) * chunk]
            if i == 0:
                part = c
            elif i == 1:
                part = self._tdnn(f"{prefix}.res2net_block.blocks.{i - 1}", c, k, d)
            else:
                part = self._tdnn(f"{prefix}.res2net_block.blocks.{i - 1}", c + part, k, d)
            outs.append(part)
        h = torch.cat(outs, 0)
        h = self._tdnn(prefix + ".tdnn2", h, 1, 1)
        m = h.mean(1, keepdim=True)  # [C, 1]
        se = torch.relu(F.conv1d(m[None], self.w[prefix + ".se_block.conv1.weight"], self.w[prefix + ".se_block.conv1.bias"])[0])
        se = torch.sigmoid(F.conv1d(se[None], self.w[prefix + ".se_block.conv2.weight"], self.w[prefix + ".se_block.conv2.bias"])[0])
        return h * se + x

    def forward(self, mels: np.ndarray, record: dict | None = None) -> np.ndarray:
        """log-mel [frames, 128] -> embedding [enc_dim]  (`callAsFunction` :496-524)."""
        h = torch.from_numpy(np.asarray(mels, dtype=np.float32)).T.contiguous()  # [mel, T]
        h = self._tdnn("blocks.0", h, KERNELS[0], DILATIONS[0])
        hs = []
        for i in (1, 2, 3):
            h = self._se_res2net(f"blocks.{i}", h, KERNELS[i], DILATIONS[i])
            hs.append(h)
        h = self._tdnn("mfa", torch.cat(hs, 0), KERNELS[4], DILATIONS[4])
        if record is not None:
            record["mfa"] = h.T.numpy().copy()
        # AttentiveStatisticsPooling (:355-397)
        mean = h.mean(1, keepdim=True)
        std = torch.sqrt(h.var(1, keepdim=True, unbiased=False) + EPS)
        T = h.shape[1]
        att = torch.cat([h, mean.expand(-1, T), std.expand(-1, T)], 0)
        att = torch.tanh(self._tdnn("asp.tdnn", att, 1, 1))
        att = F.conv1d(att[None], self.w["asp.conv.weight"], self.w["asp.conv.bias"])[0]
        att = torch.softmax(att, dim=1)
        wmean = (att * h).sum(1, keepdim=True)
        wstd = torch.sqrt(torch.clamp((att * (h - wmean) ** 2).sum(1, keepdim=True), min=EPS))
        pooled = torch.cat([wmean, wstd], 0)  # [2C, 1]
        out = F.conv1d(pooled[None], self.w["fc.weight"], self.w["fc.bias"])[0, :, 0]
        return out.numpy()

    def extract_embedding(self, audio: np.ndarray, record: dict | None = None) -> np.ndarray:
        """`extractEmbedding` (:526-542)."""
        mels = mel_spectrogram(audio)
        if record is not None:
            record["mels"] = mels
        return self.forward(mels, record)
