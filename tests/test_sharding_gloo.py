"""CPU, world_size 2 over gloo: the N > 1 host logic (stream sharding, validation gather,
max-over-ranks timing) without GPUs."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from distantspeech_b200.sharding import shard_bounds, gather_validation_streams, max_over_ranks


def test_shard_bounds_partition():
    for n in (1, 7, 8, 1024, 8192, 8193):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(n, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(8, 2, 2)


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_bounds(10, rank, world)
        # each rank "processes" its own streams: y[s] = s (stand-in for the per-stream pipeline)
        local = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 4)
        gathered = gather_validation_streams(local[0], dst=0)
        tmax = max_over_ranks(10.0 + rank)
        if rank == 0:
            q.put(([g.tolist() for g in gathered], tmax))
        else:
            assert gathered is None
            q.put((None, tmax))
    finally:
        dist.destroy_process_group()


def test_gather_and_timing_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tmaxes = sorted(r[1] for r in res)
    assert tmaxes == [11.0, 11.0]
    gathered = [r[0] for r in res if r[0] is not None][0]
    assert gathered == [[0.0] * 4, [5.0] * 4]          # first stream of rank 0 (stream 0) and rank 1 (stream 5)
