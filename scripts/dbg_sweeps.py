import ctypes as C, sys
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import bench, _lib
from gmt_planner import GMT
for name in ("cfg2","cfg3u","cfg5"):
    init, it, meta = bench.make_named_world(name); it,_=bench.bench_query(name, init, it, meta)
    g = GMT(init)
    for _ in range(2): g.run_step(it, iter_limit=1000)
    d = (C.c_ulonglong * 8)(); g._lib.gmt_debug_counters(g._handle, d)
    r=g.last_result
    print(name, "ms %.3f evals %d (words %d) sweeps %d: free %d, hit in round 1 %d, hit later %d" % (r.device_ms, r.edge_evals, 4*r.edge_evals, r.word_checks, d[5], d[6], d[7]))
