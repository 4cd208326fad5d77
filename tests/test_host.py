"""CPU: host-side logic and the C-ABI surface (no compute calls: there is no GPU here and the library has no CPU path)."""
import ctypes
import os
import re
import struct

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_header_symbol():
    import qwen3tts_b200 as q
    from qwen3tts_b200 import _abi

    L = q.lib()
    hdr = open(os.path.join(ROOT, "include", "qwen3tts_b200.h")).read()
    declared = set(re.findall(r"\b(q3tts_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"q3tts_handle", "q3tts_stream", "q3tts_status"}
    assert declared, "no prototypes found in the header"
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qwen3tts_b200.h but not exported"
    assert declared == set(_abi.SYMBOLS), (declared ^ set(_abi.SYMBOLS))
    assert L.q3tts_abi_version() == 1


def test_struct_sizes_match_c_defaults():
    """default_options / default_request write struct_size = sizeof(C struct): the ctypes mirror must agree."""
    import qwen3tts_b200 as q
    from qwen3tts_b200 import _abi

    o = _abi.Options()
    q.lib().q3tts_default_options(ctypes.byref(o))
    assert o.struct_size == ctypes.sizeof(_abi.Options)
    assert (o.max_batch, o.kv_capacity, o.max_frames, o.use_cuda_graph, o.load_codec) == (1, 512, 2400, 1, 1)
    r = _abi.Request()
    q.lib().q3tts_default_request(ctypes.byref(r))
    assert r.struct_size == ctypes.sizeof(_abi.Request)
    # sampler defaults of the reference (Qwen3Talker.swift:274-281, 335-336)
    assert r.speaker_id == -1 and abs(r.temperature - 0.9) < 1e-7 and r.top_k == 0 and abs(r.repetition_penalty - 1.05) < 1e-7 and r.max_tokens == 1200


def test_no_gpu_fails_loudly_not_silently():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is visible")
    import qwen3tts_b200 as q

    with pytest.raises(q.Q3Error) as e:
        q.Engine("/nonexistent")
    assert e.value.status == -9 and "no CPU path" in e.value.message  # Q3TTS_ERR_NO_DEVICE
    with pytest.raises(q.Q3Error):
        q.dequantize(np.zeros((1, 8), np.uint32), np.ones((1, 1), np.float32), np.zeros((1, 1), np.float32))


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "mlx-swift-qwen3-tts_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".h", ".swift")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


# ---- TextChunker: the assertions of Tests/Qwen3TTSTests/TextChunkerTests.swift ------------------------------------
def test_text_chunker_reference_assertions():
    from qwen3tts_b200 import TextChunker as TC

    assert TC.chunk("") == [] and TC.chunk("   \n  ") == []  # :6-14
    assert TC.chunk("Hello world.") == ["Hello world."]  # :16-21
    long = "This is the first sentence of the text and it keeps going for a while. " * 6
    chunks = TC.chunk(long)
    assert len(chunks) > 1 and all(len(c.split()) <= 35 for c in chunks)  # :50-60
    assert chunks[0].endswith(".")  # sentence boundary preferred (:23-30)
    commas = "word " * 20 + "more words here, " + "tail " * 30
    c2 = TC.chunk(commas)
    assert c2[0].endswith(",")  # comma boundary (:31-37)
    assert TC.estimate_tokens("one two three") == 50 and TC.estimate_tokens("w " * 100) == 500  # :39-48
    assert " ".join(TC.chunk(long)).split() == long.split()  # nothing lost


def test_streaming_wav_writer_format(tmp_path):
    from qwen3tts_b200 import Qwen3TTSPipeline, StreamingWAVWriter

    p = tmp_path / "a.wav"
    w = StreamingWAVWriter(str(p))
    w.write(np.array([0.0, 1.0, -1.0, 0.5, 2.0, -0.99999], np.float32))
    assert w.finalize() == 6
    data = p.read_bytes()
    assert len(data) == 44 + 12 and data[:4] == b"RIFF" and data[8:16] == b"WAVEfmt "
    riff, = struct.unpack("<I", data[4:8])
    fmt = struct.unpack("<IHHIIHH", data[16:36])
    assert riff == 36 + 12 and fmt == (16, 1, 1, 24000, 48000, 2, 16) and data[36:40] == b"data"
    pcm = np.frombuffer(data[44:], "<i2")
    # Int16(clamped * 32767) truncates toward zero (AudioSampleWriter.swift:73-76)
    assert pcm.tolist() == [0, 32767, -32767, 16383, 32767, -32766]
    back = Qwen3TTSPipeline.wav_to_float_samples(data)
    assert abs(back[3] - 16383 / 32767) < 1e-7


def test_synthetic_tokenizer_is_deterministic_and_in_vocab():
    from qwen3tts_b200 import SyntheticTokenizer

    t = SyntheticTokenizer(640)
    a = t.encode("<|im_start|>assistant\nHello world!<|im_end|>\n<|im_start|>assistant\n")
    assert a == t.encode("<|im_start|>assistant\nHello world!<|im_end|>\n<|im_start|>assistant\n")
    assert len(a) == 11 and all(0 <= i < 640 for i in a)  # 11 ids, like the reference's "Hello world!" prompt
    assert a[0] == a[8] and a[1] == a[9]


def test_checkpoint_writer_key_scheme(tmp_path):
    """The synthetic checkpoints use the reference's on-disk names."""
    from safetensors import safe_open

    from conftest import ckpt

    d = ckpt("tiny", 4)
    with safe_open(os.path.join(d, "model.safetensors"), "pt") as f:
        keys = set(f.keys())
        assert "talker.model.layers.0.self_attn.q_proj.weight" in keys and "talker.model.layers.0.self_attn.q_proj.scales" in keys
        assert "talker.code_predictor.model.codec_embedding.14.weight" in keys and "talker.code_predictor.lm_head.14.biases" in keys
        assert "talker.code_predictor.small_to_mtp_projection.bias" in keys and "talker.text_projection.linear_fc1.bias" in keys
        w = f.get_tensor("talker.model.layers.0.mlp.down_proj.weight")
        assert str(w.dtype) == "torch.uint32" and tuple(w.shape) == (256, 512 * 4 // 32)
        assert tuple(f.get_tensor("talker.model.layers.0.mlp.down_proj.scales").shape) == (256, 512 // 64)
    with safe_open(os.path.join(d, "speech_tokenizer", "model.safetensors"), "pt") as f:
        keys = set(f.keys())
        for k in ("decoder.quantizer.rvq_first.vq.layers.0._codebook.embedding_sum", "decoder.quantizer.rvq_rest.vq.layers.14._codebook.cluster_usage",
                  "decoder.pre_conv.conv.weight", "decoder.pre_transformer.layers.1.self_attn_layer_scale.scale", "decoder.upsample.1.1.pwconv2.bias",
                  "decoder.decoder.0.conv.weight", "decoder.decoder.4.block.4.conv2.conv.weight", "decoder.decoder.5.alpha", "decoder.decoder.6.conv.bias"):
            assert k in keys, k
        assert tuple(f.get_tensor("decoder.decoder.1.block.1.conv.weight").shape) == (192, 96, 16)  # transposed conv [C_in, C_out, 2s]
        assert tuple(f.get_tensor("decoder.upsample.0.1.dwconv.conv.weight").shape) == (128, 1, 7)


def test_safetensors_header_is_bounds_checked_without_a_gpu(tmp_path):
    """Untrusted checkpoint parsing (csrc/safetensors.h): header length, dimensions, offsets and byte counts are validated before
    anything is sized from them; q3tts_safetensors_check exposes the parser without a device."""
    import json
    import struct

    import qwen3tts_b200 as q
    from qwen3tts_b200 import _abi as A

    def write(name, hdr, data=b"\0" * 64, hlen=None):
        h = json.dumps(hdr).encode()
        p = tmp_path / name
        p.write_bytes(struct.pack("<Q", len(h) if hlen is None else hlen) + h + data)
        return str(p)

    good = {"a": {"dtype": "F32", "shape": [4, 4], "data_offsets": [0, 64]}, "__metadata__": {"format": "pt"}}
    assert q.safetensors_check(write("ok.safetensors", good)) == (1, 64)
    bad = {
        "huge_header_length": (good, dict(hlen=2 ** 63 + 5)),          # hlen + 8 would wrap around
        "negative_dims": ({"a": {"dtype": "F32", "shape": [-4, -4], "data_offsets": [0, 64]}}, {}),
        "bytes_vs_shape": ({"a": {"dtype": "F32", "shape": [4, 8], "data_offsets": [0, 64]}}, {}),
        "past_the_end": ({"a": {"dtype": "F32", "shape": [4, 4], "data_offsets": [32, 96]}}, {}),
        "reversed_offsets": ({"a": {"dtype": "F32", "shape": [4, 4], "data_offsets": [64, 0]}}, {}),
        "unknown_dtype": ({"a": {"dtype": "Q7", "shape": [64], "data_offsets": [0, 64]}}, {}),
    }
    for name, (hdr, kw) in bad.items():
        with pytest.raises(q.Q3Error) as e:
            q.safetensors_check(write(name + ".safetensors", hdr, **kw))
        assert e.value.status == A.ERR_BAD_WEIGHTS, name
    with pytest.raises(q.Q3Error) as e:
        q.safetensors_check(str(tmp_path / "missing.safetensors"))
    assert e.value.status == A.ERR_FILE_NOT_FOUND
    # many failing opens must not leak descriptors / mappings (the constructor cleans up on every failing path)
    import os
    before = len(os.listdir("/proc/self/fd"))
    for _ in range(200):
        with pytest.raises(q.Q3Error):
            q.safetensors_check(write("again.safetensors", bad["bytes_vs_shape"][0]))
    assert len(os.listdir("/proc/self/fd")) <= before + 2


def test_request_structs_own_their_buffers():
    """GenRequest.to_c() must not share keep-alive state between calls (the same request may appear several times in a batch)."""
    import qwen3tts_b200 as q

    r = q.GenRequest(text_ids=list(range(100, 120)), speaker_id=2861, forced_codes=np.arange(32, dtype=np.int32).reshape(2, 16))
    a, b = r.to_c(), r.to_c()
    assert a.text_ids[0] == 100 and b.text_ids[19] == 119 and a.n_text_ids == b.n_text_ids == 20
    assert [a.text_ids[i] for i in range(20)] == list(range(100, 120))  # still valid after the second to_c()
    assert a.forced_codes[31] == 31 and a.n_forced_frames == 2
    assert not hasattr(r, "_keep")


def test_baseline_init_checkpoint_statistics(tmp_path):
    """init="baseline" = BASELINE.md §3: every matrix, head and embedding N(0, 0.02^2), norm weights 1."""
    import torch
    from oracle import checkpoint

    t, _ = checkpoint.preset("tiny")
    w = checkpoint.talker_tensors(t, 0, "f32", 0, init="baseline")
    for k, v in w.items():
        if "norm" in k:
            assert torch.all(v == 1.0), k
        elif v.ndim == 2:
            assert abs(float(v.std()) - 0.02) < 0.004, (k, float(v.std()))
    s = checkpoint.talker_tensors(t, 0, "f32", 0)
    assert float(s["talker.codec_head.weight"].std()) > 0.2  # the stress init keeps its wide heads


def test_offline_dequant_checkpoint_form(tmp_path):
    """quant_key="quantization_config": packed leaves without `quantization` -> the oracle dequantises them to fp16 at load
    (Qwen3Talker.swift:139-175)."""
    import json
    from oracle import checkpoint, mlx_quant, talker as otalker
    from safetensors.numpy import load_file

    d = checkpoint.write_checkpoint(str(tmp_path / "qc"), "tiny", bits=4, quant_key="quantization_config", with_codec=False)
    cfg = json.load(open(d + "/config.json"))
    assert "quantization" not in cfg and cfg["quantization_config"] == {"group_size": 64, "bits": 4}
    o = otalker.TalkerOracle(d)
    assert not o.pre_quantized and o.bits == 4
    import torch
    from safetensors.torch import load_file as tload
    raw = tload(d + "/model.safetensors")
    k = "talker.model.layers.0.self_attn.q_proj"
    want = mlx_quant.dequantize(raw[k + ".weight"].view(torch.int32).numpy().view(np.uint32), raw[k + ".scales"], raw[k + ".biases"], 64, 4, "f16")
    assert np.array_equal(o.w["layers.0.self_attn.q_proj.weight"].numpy(), want)
    assert np.array_equal(want, want.astype(np.float16).astype(np.float32))  # fp16-representable
