"""CPU, world size 2, gloo: the request-parallel plumbing (sharding, max-over-ranks time, sum of work, length gather)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
    from qwen3tts_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11
    idx = parallel.shard_indices(n, rank, world)
    lengths = [1920 * (i + 1) for i in idx]  # pretend utterance i produced (i+1) frames
    secs = 1.0 + rank  # rank 1 is slower
    t, sums = parallel.aggregate(dist, "cpu", secs, sum(lengths), extra_sums=(len(idx),))
    full = parallel.gather_lengths(dist, "cpu", lengths, n, rank, world)
    # result gather: utterance i's "PCM" is a ramp of (i+1)*7 samples starting at 1000*i
    import numpy as np
    pcm = [np.arange((i + 1) * 7, dtype=np.float32) + 1000.0 * i for i in idx]
    got, glen, moved = parallel.gather_pcm(dist, "cpu", pcm, n, rank, world)
    ok = True
    if rank == 0:
        ok = len(got) == n and all(np.array_equal(got[i], np.arange((i + 1) * 7, dtype=np.float32) + 1000.0 * i) for i in range(n))
        ok = ok and moved == 4 * sum((i + 1) * 7 for i in parallel.shard_indices(n, 1, world))
    else:
        ok = got is None
    ok = ok and glen == [(i + 1) * 7 for i in range(n)]
    dist.barrier()
    q.put((rank, idx, t, sums, full, ok))
    dist.destroy_process_group()


def test_request_parallel_plumbing_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, i0, t0, s0, f0, ok0), (r1, i1, t1, s1, f1, ok1) = res
    assert ok0 and ok1  # rank 0 holds every utterance's PCM in request order
    assert sorted(i0 + i1) == list(range(11)) and not set(i0) & set(i1)  # a partition of the requests
    assert t0 == t1 == 2.0  # time = max over ranks
    assert s0 == s1 and s0[0] == 1920 * sum(range(1, 12)) and s0[1] == 11  # work = sum over ranks
    assert f0 == f1 == [1920 * (i + 1) for i in range(11)]


def test_single_process_is_identity():
    sys.path.insert(0, os.path.join(ROOT, "mlx-swift-qwen3-tts_b200"))
    from qwen3tts_b200 import parallel

    assert parallel.shard_indices(5, 0, 1) == [0, 1, 2, 3, 4]
    t, s = parallel.aggregate(None, "cpu", 3.5, 100, (7,))
    assert t == 3.5 and s == [100.0, 7.0]
    assert parallel.gather_lengths(None, "cpu", [1, 2, 3], 3, 0, 1) == [1, 2, 3]
    import numpy as np
    out, lens, moved = parallel.gather_pcm(None, "cpu", [np.ones(3, np.float32), np.zeros(5, np.float32)], 2, 0, 1)
    assert [o.size for o in out] == lens == [3, 5] and moved == 0
