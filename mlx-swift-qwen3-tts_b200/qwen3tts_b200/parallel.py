"""Request-parallel multi-GPU plumbing: one process per GPU, a full replica each, utterances sharded by
index, NO collective on the data path.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) carries only the barrier,
the max-over-ranks time, and -- after the timed region -- the gather of the results (per-utterance sample counts + PCM) to rank 0
(`gather_pcm`, BASELINE.json north_star: "NCCL only to gather results, none on the hot path")."""
from __future__ import annotations


def shard_indices(n_items: int, rank: int, world: int) -> list:
    """Utterance i -> rank i mod world (round-robin keeps length-sorted request lists balanced)."""
    return list(range(rank, n_items, world))


def aggregate(dist, device, seconds: float, samples: int, extra_sums=()):
    """(max over ranks of `seconds`, sum over ranks of `samples` and of each entry of `extra_sums`)."""
    import torch

    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    s = torch.tensor([float(samples)] + [float(x) for x in extra_sums], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
    return float(t[0]), [float(x) for x in s.tolist()]


def gather_lengths(dist, device, local_lengths: list, n_items: int, rank: int, world: int) -> list:
    """Sample count of every utterance of the job, in request order, on every rank (the only 'result gather' the path has:
    PCM stays on the rank that produced it unless the caller asks for it)."""
    import torch

    full = torch.zeros(n_items, dtype=torch.int64, device=device)
    idx = shard_indices(n_items, rank, world)
    if idx:
        full[torch.tensor(idx, device=device)] = torch.tensor(local_lengths, dtype=torch.int64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(full, op=dist.ReduceOp.SUM)
    return [int(x) for x in full.tolist()]


def gather_pcm(dist, device, local_pcm: list, n_items: int, rank: int, world: int, dst: int = 0):
    """Result gather over NCCL / gloo: every utterance's PCM, in request order, on rank `dst` (None elsewhere).

    local_pcm[j] is the float32 PCM of utterance shard_indices(n_items, rank, world)[j].  Two collectives: an all-reduce of the
    length table (so every rank can size its part), then one variable-size gather (`dist.gather` of per-rank flat buffers padded
    to the largest shard).  Returns (list of numpy arrays on dst | None, lengths, bytes moved into dst)."""
    import numpy as np
    import torch

    idx = shard_indices(n_items, rank, world)
    assert len(idx) == len(local_pcm)
    lengths = gather_lengths(dist, device, [int(p.size) for p in local_pcm], n_items, rank, world)
    multi = dist is not None and dist.is_initialized() and dist.get_world_size() > 1
    if not multi:
        return [np.asarray(p, dtype=np.float32) for p in local_pcm], lengths, 0
    per_rank = [sum(lengths[i] for i in shard_indices(n_items, r, world)) for r in range(world)]
    cap = max(max(per_rank), 1)
    flat = torch.zeros(cap, dtype=torch.float32, device=device)
    if local_pcm:
        cat = np.concatenate([np.asarray(p, dtype=np.float32).ravel() for p in local_pcm]) if per_rank[rank] else np.zeros(0, np.float32)
        flat[: cat.size] = torch.from_numpy(cat).to(device)
    bufs = [torch.empty(cap, dtype=torch.float32, device=device) for _ in range(world)] if rank == dst else None
    dist.gather(flat, bufs, dst=dst)
    if rank != dst:
        return None, lengths, 0
    out = [None] * n_items
    for r in range(world):
        host = bufs[r][: per_rank[r]].cpu().numpy()
        off = 0
        for i in shard_indices(n_items, r, world):
            out[i] = host[off: off + lengths[i]]
            off += lengths[i]
    return out, lengths, 4 * sum(per_rank[r] for r in range(world) if r != dst)
