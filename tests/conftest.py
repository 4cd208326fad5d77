import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

CKPT_ROOT = os.environ.get("Q3TTS_TEST_CKPT", "/tmp/q3tts_test_ckpt")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")
    config.addinivalue_line("markers", "slow: full-size model cases")


def ckpt(name="tiny", bits=8, dtype="bf16", seed=0, **kw):
    """Synthetic checkpoint directory (cached under /tmp by its stamp)."""
    from oracle import checkpoint

    tag = f"{name}_b{bits}_{dtype}_s{seed}" + "".join(f"_{k}{v}" for k, v in sorted(kw.items()))
    return checkpoint.write_checkpoint(os.path.join(CKPT_ROOT, tag), name, bits=bits, dtype=dtype, seed=seed, **kw)


TEXT_IDS = [11, 21, 22, 23, 24, 25, 26, 27, 28, 29, 30, 31, 32]  # 13 ids: role(3) + first + 4 trailing + 5 tail


@pytest.fixture(scope="session")
def tiny8():
    return ckpt("tiny", 8)


@pytest.fixture(scope="session")
def tiny4():
    return ckpt("tiny", 4)


@pytest.fixture(scope="session")
def tiny_bf16():
    return ckpt("tiny", 0, "bf16")


@pytest.fixture(scope="session")
def engines():
    """Cache of qwen3tts_b200.Engine per (dir, kwargs) for the GPU tests."""
    import qwen3tts_b200 as q

    cache = {}

    def get(d, **kw):
        key = (d, tuple(sorted(kw.items())))
        if key not in cache:
            kw.setdefault("max_frames", 256)
            cache[key] = q.Engine(d, **kw)
        return cache[key]

    yield get
    for e in cache.values():
        e.close()


@pytest.fixture(scope="session")
def oracles():
    from oracle import codec as ocodec, talker as otalker

    cache = {}

    def get(d, kind="talker"):
        key = (d, kind)
        if key not in cache:
            cache[key] = otalker.TalkerOracle(d) if kind == "talker" else ocodec.load_codec(d)
        return cache[key]

    return get
