"""Committed golden vectors (tests/golden/tiny_golden.npz, made by tests/golden/make_golden.py from the oracle).
CPU: the oracle still reproduces them.  GPU: the CUDA path hits the same vectors through the C ABI."""
import os

import numpy as np
import pytest
import torch

from conftest import ckpt

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tiny_golden.npz"))


@pytest.mark.parametrize("bits", [4, 8])
def test_oracle_dequant_golden(bits):
    from oracle import mlx_quant

    p, s, b = mlx_quant.quantize(G["q_w"], 64, bits, "bf16")
    assert np.array_equal(p, G[f"q{bits}_packed"]) and np.array_equal(s, G[f"q{bits}_scales"]) and np.array_equal(b, G[f"q{bits}_biases"])
    assert np.array_equal(mlx_quant.dequantize(p, s, b, 64, bits, "f32"), G[f"q{bits}_deq_f32"])
    assert np.array_equal(mlx_quant.dequantize(p, s, b, 64, bits, "f16"), G[f"q{bits}_deq_f16"])


@pytest.mark.parametrize("name,bits", [("tiny8", 8), ("tiny4", 4)])
def test_oracle_talker_golden(name, bits):
    from oracle import talker as otalker

    orc = otalker.TalkerOracle(ckpt("tiny", bits))
    rec = {}
    orc.generate_codes(otalker.Request(text_ids=G["text_ids"].tolist(), speaker_id=2861, temperature=0.0, max_tokens=8), record=rec, filter_invalid=False)
    assert np.abs(rec["code0_logits"][:2] - G[f"{name}_code0_logits"]).max() < 2e-4
    assert np.abs(rec["cp_logits"][:1, :3] - G[f"{name}_cp_logits"]).max() < 2e-4
    if G[f"{name}_margins"].min() > 1e-3:
        assert np.array_equal(np.asarray(rec["raw_frames"]), G[f"{name}_frames"])


def test_oracle_codec_golden():
    from oracle import codec as ocodec

    cd = ocodec.load_codec(ckpt("tiny", 8))
    tc = torch.as_tensor(G["codec_codes"]).transpose(1, 2).contiguous()
    first, rest = cd.rvq_embed(tc)
    assert np.array_equal(first.numpy(), G["codec_first"]) and np.array_equal(rest.numpy(), G["codec_rest"])
    assert np.abs(cd.decode(tc).reshape(-1).numpy() - G["codec_pcm"]).max() < 1e-4


# ------------------------------------------------------------------------------------------------ GPU against the same vectors
@pytest.mark.gpu
@pytest.mark.parametrize("bits", [4, 8])
def test_gpu_dequant_golden(bits):
    import qwen3tts_b200 as q

    for odt in ("f32", "f16"):
        got = q.dequantize(G[f"q{bits}_packed"], G[f"q{bits}_scales"], G[f"q{bits}_biases"], 64, bits, "bf16", odt)
        assert np.array_equal(got.view(np.uint32), G[f"q{bits}_deq_{odt}"].view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("name,bits", [("tiny8", 8), ("tiny4", 4)])
def test_gpu_talker_golden(name, bits, engines):
    import qwen3tts_b200 as q

    eng = engines(ckpt("tiny", bits))
    forced = G[f"{name}_frames"]
    frames, lg = eng.generate_codes(q.GenRequest(text_ids=G["text_ids"].tolist(), speaker_id=2861, temperature=0.0, max_tokens=len(forced),
                                                 forced_codes=forced, keep_invalid_frames=True, want_logits=2))
    assert np.abs(lg["code0_logits"][:2] - G[f"{name}_code0_logits"]).max() <= 1e-2
    assert np.abs(lg["cp_logits"][:1, :3] - G[f"{name}_cp_logits"]).max() <= 1e-2
    if G[f"{name}_margins"].min() > 2e-2:
        free = eng.generate_codes(q.GenRequest(text_ids=G["text_ids"].tolist(), speaker_id=2861, temperature=0.0, max_tokens=len(forced), keep_invalid_frames=True))
        assert np.array_equal(free, forced)


@pytest.mark.gpu
def test_gpu_codec_golden(engines):
    eng = engines(ckpt("tiny", 8))
    first, rest = eng.rvq_embed(G["codec_codes"])
    assert np.array_equal(first[0].view(np.uint32), G["codec_first"][0].view(np.uint32))
    assert np.array_equal(rest[0].view(np.uint32), G["codec_rest"][0].view(np.uint32))
    pcm = eng.decode(G["codec_codes"])[0]
    err = np.sum((pcm.astype(np.float64) - G["codec_pcm"]) ** 2)
    assert 10 * np.log10(np.sum(G["codec_pcm"].astype(np.float64) ** 2) / max(err, 1e-300)) >= 40.0
