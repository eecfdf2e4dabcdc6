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


# ---------------------------------------------------------------------------------------------------------
# BASELINE config 5's protocol on CPU: node-range sharding + all-reduce(min) of packed (cost, parent) keys,
# two gloo ranks, the per-rank "kernels" being the CPU oracle.  Must rebuild the unsharded oracle tree.
# ---------------------------------------------------------------------------------------------------------
def _sharded_oracle_plan(rank, world, init, it, iter_limit):
    import torch
    from oracle.gmt_oracle import OraclePort, _exclusive_scan, _obstacles
    o = OraclePort("libm")
    states = np.ascontiguousarray(init['states'], np.float32)
    n = states.shape[0]
    nbr = np.ascontiguousarray(init['neighbors'], np.int32)
    num = np.ascontiguousarray(init['num_neighbors'], np.int32)
    nbr_index = _exclusive_scan(num)
    obs, m = _obstacles(it['obstacles'], it['num_obs'])
    lo, hi = multi_gpu.shard_range(n, rank, world)
    cost = np.full(n, np.inf, np.float32); parent = np.full(n, -1, np.int32)
    opn = np.zeros(n, np.int32); unexp = np.ones(n, np.int32)
    cost[it['start']] = 0; opn[it['start']] = 1; unexp[it['start']] = 0
    thr = np.array([it['threshold']], np.float32); radius = np.array([it['radius']], np.float32)
    G = np.zeros(n, np.int32)
    iteration = 0
    while True:
        iteration += 1
        o.k_wavefront(G, opn, cost, thr, n, 128)
        o.k_grow_threshold(thr, radius)
        if iteration >= iter_limit:
            return 1, iteration, cost, parent
        if G[it['goal']] == 1:
            return 0, iteration, cost, parent
        g_list = np.flatnonzero(G == 1).astype(np.int32)
        if len(g_list) == 0:
            continue
        y_list = np.flatnonzero(opn == 1).astype(np.int32)
        xind = np.zeros(n, np.int32)
        o.k_neighbor_indicator(xind, g_list, unexp, nbr, num, nbr_index, len(g_list), 128)
        x_list = np.flatnonzero(xind == 1).astype(np.int32)
        if len(x_list) == 0:
            continue
        mine = np.ascontiguousarray(x_list[(x_list >= lo) & (x_list < hi)])          # my node range only
        scratch_open, scratch_unexp = opn.copy(), unexp.copy()
        if len(mine):
            o.k_dubin_connection(cost, parent, mine, y_list, states, scratch_open, scratch_unexp, len(mine), len(y_list),
                                 obs, m, radius, 128)
        keys = torch.from_numpy(multi_gpu.pack_keys(cost, parent))
        multi_gpu.allreduce_min(keys)                                                # the per-wavefront exchange
        cost, parent = (a.copy() for a in multi_gpu.unpack_keys(keys.numpy()))
        opn[y_list] = 0                                                              # kernel.py:445
        opn[x_list] = 1; unexp[x_list] = 0                                           # every x is connected (1e10 at worst)


def _sharded_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import synthetic_world as sw
    from oracle.world_oracle import build_neighbors, sample_states
    init, it, meta = sw.make_world(n=400, m=6, seed=18, sampler=sample_states, neighbor_builder=build_neighbors)
    res = _sharded_oracle_plan(rank, world, init, it, 40)
    dist.barrier()
    out.put((rank, res[0], res[1], res[2].tobytes(), res[3].tobytes()))
    dist.destroy_process_group()


def test_node_range_sharding_protocol_two_ranks_gloo():
    import synthetic_world as sw
    from oracle.gmt_oracle import OraclePort
    from oracle.world_oracle import build_neighbors, sample_states
    init, it, meta = sw.make_world(n=400, m=6, seed=18, sampler=sample_states, neighbor_builder=build_neighbors)
    ref = OraclePort("libm").plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'],
                                  it['radius'], it['threshold'], it['obstacles'], it['num_obs'], iter_limit=40, trace=False)
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_sharded_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    got = [out.get(timeout=180) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, status, iters, cost_b, parent_b in got:
        assert (status, iters) == (ref["status"], ref["iterations"])
        assert cost_b == ref["cost"].tobytes() and parent_b == ref["parent"].tobytes()     # bit for bit, on both ranks


def test_packed_keys_order_like_the_reference():
    cost = np.array([1.5, 1.5, np.inf, 0.0, 1e10], np.float32)
    parent = np.array([7, 3, -1, -1, 2], np.int32)
    k = multi_gpu.pack_keys(cost, parent)
    assert (k >= 0).all()                                    # signed order == unsigned order
    assert k[1] < k[0] < k[4] < k[2] and k[3] < k[1]         # lower cost first, ties -> lower parent, inf last
    c2, p2 = multi_gpu.unpack_keys(k)
    assert c2.tobytes() == cost.tobytes() and p2.tolist() == parent.tolist()
