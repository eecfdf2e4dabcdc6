"""Profiling target (run under ncu through gpurun): a few launches of ONE kernel on a bench workload, nothing else.
usage: prof_run.py batch [queries] | cfg2 | cfg3u | cfg5 | words6 | oriented | edge"""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import bench, synthetic_world as sw
from gmt_planner import GMT
mode = sys.argv[1] if len(sys.argv) > 1 else "batch"
if mode == "batch":
    q = int(sys.argv[2]) if len(sys.argv) > 2 else 1184
    init, it, meta = bench.make_named_world("cfg4")
    g = GMT(init)
    g.set_obstacles(it['obstacles'], it['num_obs'])
    starts, goals, _ = bench.batch_queries(g, meta)
    for rep in range(3):
        res, _r = sw.plan_batch(g, starts[:q], goals[:q])
    print("batch of %d: %.2f ms -> %.0f plans/s" % (q, res[0].device_ms, q / res[0].device_ms * 1e3))
elif mode == "edge":                    # the kernel-level entry on 2^18 neighbour pairs of the cfg4 map (bench.py's edge_kernel)
    import ctypes as C, _lib
    init, it, meta = bench.make_named_world("cfg4")
    g = GMT(init)
    g.set_obstacles(it['obstacles'], it['num_obs'])
    rng = np.random.default_rng(7)
    nbr, num = init['neighbors'], init['num_neighbors']
    rows = np.repeat(np.arange(meta["n"], dtype=np.int32), num)
    sel = rng.integers(0, len(nbr), 1 << 18)
    ey, ex = np.ascontiguousarray(rows[sel]), np.ascontiguousarray(nbr[sel].astype(np.int32))
    ms = C.c_float()
    for rep in range(3):
        _lib.check(g._lib.gmt_edge_costs(g._handle, ey.ctypes.data, ex.ctypes.data, 1 << 18, 2.0, None, None, None, C.byref(ms)), "edge")
    print("edge kernel: %.3f ms for %d pairs -> %.0f M edge checks/s" % (ms.value, 1 << 18, (1 << 18) / ms.value * 1e-3))
else:
    name = {"words6": "cfg2", "oriented": "cfg2"}.get(mode, mode)
    init, it, meta = bench.make_named_world(name)
    g = GMT(init, words=6 if mode == "words6" else 4)
    it, _c = bench.bench_query(name, init, it, meta)
    if mode == "oriented":      # the same boxes as oriented rectangles turned by seeded angles (as bench.py --oriented)
        obs = np.ascontiguousarray(it['obstacles'], np.float32)
        lo, hi = np.minimum(obs[:, :2], obs[:, 2:]), np.maximum(obs[:, :2], obs[:, 2:])
        rect = np.concatenate([(lo + hi) / 2, (hi - lo) / 2, np.random.default_rng(5).uniform(-np.pi, np.pi, (len(obs), 1))], 1)
        it = dict(it, obstacles=rect.astype(np.float32), obstacle_format='oriented')
        g.run_step(dict(it, goal=-1), iter_limit=1000)               # the turned rectangles change what is reachable
        it = dict(it, goal=int(g.last_result.farthest_frontier))
    for rep in range(3):
        g.run_step(it, iter_limit=1000)
    r = g.last_result
    print("%s: %.3f ms, %d wavefronts, status %d" % (mode, r.device_ms, r.iterations, r.status))
