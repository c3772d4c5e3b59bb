"""The N > 1 path on CPU: two processes over gloo. Tiles are independent streams, so the multi-GPU logic is sharding
plus the max-over-ranks reduction of the timings; there is no collective on the data path to test. Each rank takes its
contiguous tile range, generates its own tiles (the generator must give the same tiles whatever the split), encodes
them with the oracle (standing in for the device here: no GPU in this test), and the gathered sizes must equal the
single-process result."""
import os
import socket

import numpy as np
import pytest

import qb3_b200 as q
from helpers import oracle, synth_tiles

NT, W, H, B = 10, 32, 24, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch
    import torch.distributed as dist
    from bench import device_synth_tiles
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    try:
        lo, hi = q.shard_range(NT, rank, world)
        mine = device_synth_tiles(hi - lo, W, H, B, 0, torch.device("cpu"), t0=lo).numpy().reshape(hi - lo, H, W, B)
        assert np.array_equal(mine, synth_tiles(hi - lo, W, H, B, np.uint8, t0=lo))
        sizes = torch.zeros(NT, dtype=torch.int64)
        for t in range(lo, hi):
            sizes[t] = len(oracle().encode(mine[t - lo]))
        dist.all_reduce(sizes)                                  # every tile is written by exactly one rank
        ranges = [None] * world
        dist.all_gather_object(ranges, (lo, hi))
        times = torch.tensor([1.0 + rank, 5.0 - rank], dtype=torch.float64)
        dist.all_reduce(times, op=dist.ReduceOp.MAX)            # bench.py: step time = max over ranks
        if rank == 0:
            out.put((sizes.tolist(), ranges, times.tolist()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_without_overlap():
    for n in (0, 1, 7, 4096):
        for world in (1, 2, 3, 8):
            r = [q.shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[k][1] == r[k + 1][0] for k in range(world - 1))
            assert max(b - a for a, b in r) - min(b - a for a, b in r) <= 1
    with pytest.raises(ValueError):
        q.shard_range(4, 2, 2)


@pytest.mark.timeout(300)
def test_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    sizes, ranges, times = out.get(timeout=240)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert ranges == [(0, 5), (5, 10)]
    assert times == [2.0, 5.0]
    tiles = synth_tiles(NT, W, H, B, np.uint8)
    assert sizes == [len(oracle().encode(tiles[t])) for t in range(NT)]
