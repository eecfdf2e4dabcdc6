"""Development aid: one line per workload (device ms, work counters, phase split) -- run under gpurun.
usage: quick_perf.py [cfg2 cfg3u cfg5 cfg4 ...] [opt=value ...]   (opt = a gmt_set_option knob)"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import bench, _lib, synthetic_world as sw
from gmt_planner import GMT
names = [a for a in sys.argv[1:] if '=' not in a] or ["cfg2", "cfg3u", "cfg5", "cfg4"]
opts = [a.split('=') for a in sys.argv[1:] if '=' in a]
lib = _lib.load()
for name in names:
    init, it, meta = bench.make_named_world(name)
    g = GMT(init)
    for k, v in opts:
        g.set_option(k, int(v))
    if name == "cfg4":
        g.set_obstacles(it['obstacles'], it['num_obs'])
        starts, goals, _ = bench.batch_queries(g, meta)
        best = 1e30
        for rep in range(4):
            res, _r = sw.plan_batch(g, starts, goals)
            best = min(best, res[0].device_ms)
        print("%-6s %9.3f ms -> %8.0f plans/s; pairs evaluated %d, words swept %d, reached %d"
              % (name, best, len(starts) / best * 1e3, sum(r.edge_evals for r in res), sum(r.word_checks for r in res),
                 sum(r.status == 0 for r in res)), flush=True)
        continue
    it, _c = bench.bench_query(name, init, it, meta)
    ms = []
    for rep in range(8):
        g.run_step(it, iter_limit=1000)
        ms.append(g.last_result.device_ms)
    r = g.last_result
    d = (C.c_ulonglong * 8)()
    lib.gmt_debug_counters(g._handle, d)
    w = max(r.iterations, 1)
    print("%-6s %9.3f ms (min of 8; median %.3f); %d wavefronts, pairs evaluated %d, words swept %d; per wavefront us: frontier %.1f "
          "select %.1f connect %.1f (fused step: list+keys in after %.1f, rows after %.1f)"
          % (name, min(ms[2:]), float(np.median(ms[2:])), r.iterations, r.edge_evals, r.word_checks,
             d[0] / 1e3 / w, d[1] / 1e3 / w, d[4] / 1e3 / w, d[5] / 1e3 / w, d[6] / 1e3 / w), flush=True)
    del g
