"""Request-parallel multi-GPU plumbing: one process per GPU, a full replica each, utterances sharded by
index, NO collective on the data path.  `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) carries only the barrier,
the max-over-ranks time and the gather of per-utterance sample counts."""
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
