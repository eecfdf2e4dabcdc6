"""CPU: the N > 1 host logic (query sharding, max-over-ranks, route gather) with gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import multi_gpu


def test_shard_range_partitions_exactly():
    for q in (0, 1, 7, 4096, 4097):
        for world in (1, 2, 3, 8):
            cover = []
            for r in range(world):
                lo, hi = multi_gpu.shard_range(q, r, world)
                cover += list(range(lo, hi))
                assert 0 <= hi - lo <= q // world + 1
            assert cover == list(range(q))
    with pytest.raises(ValueError):
        multi_gpu.shard_range(4, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    starts, goals = np.arange(11), np.arange(11) + 100
    s, g = multi_gpu.shard_queries(starts, goals, rank, world)
    # a fake planner: the "route" of a query is [goal, start]
    routes = [[int(b), int(a)] for a, b in zip(s, g)]
    times = multi_gpu.max_over_ranks([1.0 + rank, 5.0 - rank])
    allr = multi_gpu.gather_routes(routes)
    dist.barrier()
    if rank == 0:
        out.put((times, allr))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    times, allr = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert times == [2.0, 5.0]
    assert allr == [[100 + i, i] for i in range(11)]
