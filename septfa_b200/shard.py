"""Data-parallel sharding of independent mixtures / streams across the GPUs of one box.

Utterances never interact (GroupNorm and the TF pooling are per utterance, model/model.py:123-124,
189,195; the online PIT is per stream), so the batch is split into contiguous shards, weights are
replicated, and there is NO collective on the data path. ``torch.distributed`` is used only to
aggregate timings / counts (a few bytes) - SURVEY.md section 8(e).
"""
from __future__ import annotations


def shard_range(n_items: int, rank: int, world_size: int):
    """Contiguous [begin, end) of ``n_items`` for ``rank``; sizes differ by at most one."""
    base, rem = divmod(n_items, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def gather_max_time(seconds: float) -> float:
    """MAX over ranks of a per-rank elapsed time (all ranks get the result)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(seconds)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([seconds], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_counts(n: int) -> int:
    """SUM over ranks of a per-rank item count."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return int(n)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())
