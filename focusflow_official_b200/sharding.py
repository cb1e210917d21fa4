"""Pair sharding for multi-GPU runs.

The correlation path never mixes batch entries (`corr.py:55-59` is a batched matmul, lookups
index `[B*N]`), so work shards by image pair with NO data-path collective: one process per GPU,
weights replicated, each rank takes a contiguous slice of the pairs.  The only collectives are
the timing reductions of the benchmark (max over ranks) -- and, in the training configuration,
stock DDP gradient all-reduce, which is outside the correlation path.
"""
from __future__ import annotations

from typing import Tuple


def shard_range(total_pairs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of `total_pairs` for `rank`; sizes differ by at most one.

    The reference's training loader uses `BATCH_SIZE // world_size` per rank
    (`core/datasets.py:306`); for divisible sizes this gives the same split.
    """
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    if total_pairs < 0:
        raise ValueError("total_pairs must be >= 0")
    base, extra = divmod(total_pairs, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def job_throughput(pairs_this_rank: int, seconds_this_rank: float, world: int, device=None) -> float:
    """Whole-job pairs/s = (sum of pairs over ranks) / (max time over ranks)."""
    if world == 1:
        return pairs_this_rank / seconds_this_rank
    import torch
    import torch.distributed as dist

    t = torch.tensor([seconds_this_rank], dtype=torch.float64, device=device)
    n = torch.tensor([float(pairs_this_rank)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(n, op=dist.ReduceOp.SUM)
    return float(n.item() / t.item())
