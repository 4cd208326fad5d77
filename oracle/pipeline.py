"""Driver-level restatement: windowed codec decoding as scheduled by `Qwen3TTSPipeline` (oracle; test
infrastructure only; parity unpinned).

Follows `Qwen3TTSPipeline.swift`: `_generateStreamImpl` :484-624 (windows 18 / 8+18, final empty chunk),
`generateToFile` :644-757 (windows 16 + 8 left context), `generateBatch` :774-898 (24 + 8, crossfade :869-888),
`generate` via `Qwen3Talker.generate` (`Model/Qwen3Talker.swift:580-630`, whole-sequence decode + scrub).
"""
from __future__ import annotations

import numpy as np
import torch

SAMPLES_PER_FRAME = 1920


def _decode_frames(codec, frames) -> np.ndarray:
    """`AudioDecoder.mlxDecode(codes: [1,F,16])` (AudioDecoder.swift:167-175): transpose to [1,16,F], decode."""
    codes = torch.as_tensor(np.asarray(frames, dtype=np.int32)).reshape(1, len(frames), -1).transpose(1, 2).contiguous()
    return codec.decode(codes).reshape(-1).numpy()


def clean_samples(x: np.ndarray) -> np.ndarray:
    """NaN/Inf -> 0, clamp to [-1, 1] (Qwen3TTSPipeline.swift:565-570)."""
    x = np.where(np.isfinite(x), x, 0.0).astype(np.float32)
    return np.clip(x, -1.0, 1.0)


def valid_frames(frames):
    """Drop frames whose code0 is outside [0, 2048) (Qwen3TTSPipeline.swift:576-579)."""
    return [f for f in frames if 0 <= f[0] < 2048]


def decode_whole(codec, frames) -> np.ndarray:
    """`Qwen3Talker.generate` tail: one whole-sequence decode (Qwen3Talker.swift:606-629)."""
    if not frames:
        return np.zeros(0, np.float32)
    x = _decode_frames(codec, frames)
    if not np.all(np.isfinite(x)):
        x = clean_samples(x)
    return x


def decode_windowed(codec, frames, chunk: int, left: int):
    """Windowed decode used by stream (18, 8), file (16, 8) and batch (24, 8) modes.

    Returns list of (samples, (start, end)) per window.  Left context = the last `left` frames of the PREVIOUS
    window's new frames in stream mode (`leftContext = codes.suffix(8)`, :561) and `codes[max(0,end-8)..<end]`
    in file/batch mode (:734, :859) — identical whenever chunk >= left."""
    out = []
    ctx: list = []
    pos = 0
    while pos < len(frames):
        end = min(pos + chunk, len(frames))
        new = frames[pos:end]
        x = _decode_frames(codec, ctx + new)
        drop = len(ctx) * SAMPLES_PER_FRAME
        if drop > 0 and x.shape[0] > drop:
            x = x[drop:]
        out.append((clean_samples(x), (pos, end)))
        ctx = frames[max(0, end - left):end]
        pos = end
    return out


def stream_chunks(codec, code_chunks, decode_chunk=18, left=8):
    """`_generateStreamImpl` consumer (:572-607): code_chunks = list of frame groups as yielded by the talker
    stream (chunkSize frames each).  Returns list of dicts {samples, token_range, is_final} INCLUDING the
    trailing empty final chunk (:607)."""
    buf: list = []
    ctx: list = []
    first = True
    total = 0
    out = []

    def decode_batch(codes):
        nonlocal ctx, first
        inp = codes if first else ctx + codes
        first = False
        x = _decode_frames(codec, inp)
        drop = len(ctx) * SAMPLES_PER_FRAME
        if drop > 0 and x.shape[0] > drop:
            x = x[drop:]
        ctx = codes[-left:]
        return x

    for chunk in code_chunks:
        v = valid_frames(chunk)
        if not v:
            continue
        buf.extend(v)
        while len(buf) >= decode_chunk:
            batch, buf = buf[:decode_chunk], buf[decode_chunk:]
            s = decode_batch(batch)
            total += len(batch)
            if s.size:
                out.append({"samples": clean_samples(s), "token_range": (total - len(batch), total), "is_final": False})
    if buf:
        s = decode_batch(buf)
        total += len(buf)
        if s.size:
            out.append({"samples": clean_samples(s), "token_range": (total - len(buf), total), "is_final": True})
    out.append({"samples": np.zeros(0, np.float32), "token_range": (total, total), "is_final": True})
    return out


def crossfade_concat(chunks, crossfade=480):
    """`generateBatch` crossfade between text chunks (:869-888)."""
    all_s: list = []
    tail = np.zeros(0, np.float32)
    for i, ch in enumerate(chunks):
        ch = np.asarray(ch, dtype=np.float32)
        if ch.size == 0:
            continue
        last = i == len(chunks) - 1
        if tail.size and crossfade > 0:
            n = min(crossfade, tail.size, ch.size)
            k = np.arange(n, dtype=np.float32)
            fo = (np.float32(n) - k) / np.float32(n)
            fi = k / np.float32(n)
            all_s.append(tail[:n] * fo + ch[:n] * fi)
            ch = ch[n:]
        if last:
            all_s.append(ch)
        elif ch.size > crossfade:
            all_s.append(ch[: ch.size - crossfade])
            tail = ch[ch.size - crossfade:]
        else:
            tail = ch
    return np.concatenate(all_s) if all_s else np.zeros(0, np.float32)
