"""Replica data parallelism over independent utterances: one process per GPU, every rank holds a
full model, utterance i goes to rank i mod N, results return through host memory.  There is NO data-path
collective (nothing shards below the request level); torch.distributed is used only to gather the finished
host-side results and to agree on timing (max over ranks)."""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence, TypeVar

T = TypeVar("T")
R = TypeVar("R")


def shard_indices(n_items: int, world: int, rank: int) -> List[int]:
    """Round-robin: utterance i -> rank i mod world (keeps long and short texts mixed on every rank)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    return list(range(rank, n_items, world))


def run_sharded(items: Sequence[T], fn: Callable[[T], R], dist=None) -> Optional[List[R]]:
    """Runs fn on this rank's shard; rank 0 returns the results in the ORIGINAL order, other ranks return None."""
    world = dist.get_world_size() if dist is not None and dist.is_initialized() else 1
    rank = dist.get_rank() if world > 1 else 0
    mine = shard_indices(len(items), world, rank)
    local = [(i, fn(items[i])) for i in mine]
    if world == 1:
        return [r for _, r in local]
    gathered = [None] * world if rank == 0 else None
    dist.gather_object(local, gathered, dst=0)
    if rank != 0:
        return None
    out: List[Optional[R]] = [None] * len(items)
    for part in gathered:
        for i, r in part:
            out[i] = r
    return out


def max_over_ranks(seconds: float, dist=None, device=None) -> float:
    """Aggregate throughput is reported against the slowest rank."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return seconds
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
