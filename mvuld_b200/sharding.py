"""Batch sharding of MVulD inference over the GPUs of one node (one process per GPU).

Functions (image + token row + CPG) are independent in eval mode, so the path shards with no data-path collective.
The reference shards with ``DistributedSampler`` (/root/reference/mvuld/data/bigvul_dataset.py:170-175), which pads
the last shard by repeating samples; here every function is scored exactly once: contiguous shards, the remainder
spread over the first ranks.  ``shard_by_cost`` balances the graph branch by node count instead of function count
(graph sizes are log-normal; Swin / UniXcoder cost is constant per function).  The only collective is the optional
``gather_rows`` of the ``[n_local, 2]`` logits for metric computation (the reference all-reduces per-batch scalars,
main_bigvul.py:427-428).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """[lo, hi) of rank's contiguous shard; sizes differ by at most one and cover [0, n_total) exactly once."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_cost(costs: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous shards [lo, hi) per rank whose summed cost (e.g. CPG node counts) is as even as a prefix split allows."""
    n = len(costs)
    total = float(sum(costs))
    out, lo, acc = [], 0, 0.0
    for r in range(world):
        target = total * (r + 1) / world
        hi = lo
        while hi < n and (acc + costs[hi] <= target or hi == lo) and n - (hi + 1) >= world - r - 1:
            acc += costs[hi]
            hi += 1
        if r == world - 1:
            hi = n
        out.append((lo, hi))
        lo = hi
    return out


def batches(lo: int, hi: int, batch: int):
    """Step boundaries [a, b) of one rank's shard."""
    for a in range(lo, hi, batch):
        yield a, min(a + batch, hi)


def gather_rows(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of per-rank row blocks of unequal length (contiguous shards, rank order) -> [n_total, ...]."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert local.shape[0] == n_total
        return local
    world = dist.get_world_size(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    longest = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((longest,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)
    return torch.cat([p[: hi - lo] for p, (lo, hi) in zip(parts, sizes)], dim=0)
