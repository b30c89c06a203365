"""CPU: the N>1 host logic -- shard ranges and the world-size-2 gather over gloo."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from mvuld_b200 import sharding, synth


@pytest.mark.parametrize("n,world", [(25816, 8), (25816, 1), (7, 8), (0, 4), (64, 3)])
def test_shard_range_covers_every_function_exactly_once(n, world):
    spans = [sharding.shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1 and sum(sizes) == n
    with pytest.raises(ValueError):
        sharding.shard_range(n, world, world)


def test_shard_by_cost_balances_node_counts():
    g = torch.Generator().manual_seed(3)
    costs = synth._num_nodes(4096, g)
    spans = sharding.shard_by_cost(costs, 8)
    assert spans[0][0] == 0 and spans[-1][1] == len(costs)
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    loads = [sum(costs[lo:hi]) for lo, hi in spans]
    assert max(loads) / (sum(loads) / 8) < 1.02          # within 2 % of perfect balance
    # degenerate: fewer items than ranks still gives every rank a (possibly empty) contiguous span
    spans = sharding.shard_by_cost([5, 1], 4)
    assert spans[-1][1] == 2 and all(a[1] == b[0] for a, b in zip(spans, spans[1:]))


def test_batches():
    assert list(sharding.batches(3, 10, 4)) == [(3, 7), (7, 10)]
    assert list(sharding.batches(5, 5, 4)) == []


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = sharding.shard_range(n_total, rank, world)
        # stand-in for the per-function logits: a deterministic function of the GLOBAL function index
        idx = torch.arange(lo, hi, dtype=torch.float32)
        local = torch.stack([idx * 2.0, -idx], dim=1)
        full = sharding.gather_rows(local, n_total)
        torch.save(full, os.path.join(out_dir, f"r{rank}.pt"))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_gather_rows_world2_gloo(tmp_path):
    n_total, world = 13, 2                       # odd: the shards are unequal
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    idx = torch.arange(n_total, dtype=torch.float32)
    want = torch.stack([idx * 2.0, -idx], dim=1)
    for r in range(world):
        assert torch.equal(torch.load(tmp_path / f"r{r}.pt"), want)
