"""Development aid: where the select phase spends its time (needs a -DDBG_SELECT build, GMT_B200_LIB=...)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import _lib, synthetic_world as sw
from gmt_planner import GMT
if sys.argv[1].startswith('cfg'):                 # a bench workload by name
    import bench
    init, it, meta = bench.make_named_world(sys.argv[1])
    g = GMT(init)
    for a in sys.argv[2:]:
        g.set_option(a.split('=')[0], int(a.split('=')[1]))
    it, _c = bench.bench_query(sys.argv[1], init, it, meta)
else:
    if sys.argv[1] == '0':
        init, it, meta = sw.make_world(n=int(sys.argv[2]), m=int(sys.argv[3]), seed=23)
    else:
        init, it, meta = sw.make_world(int(sys.argv[1]))
    it = dict(it, goal=sw.pick_reachable_goal(init, it, meta["side"]))
    g = GMT(init)
for rep in range(2):
    g.run_step(it, iter_limit=1000)
res = g.last_result
out = (C.c_ulonglong * 16)()
g._lib.gmt_debug_select(g._handle, out)
v = list(out)
names = ["whole x", "pass 1 scan", "seed choice + word lengths", "seed sweeps", "masks", "pass 2 + flush"]
print("plan %.3f ms, %d wavefronts, %d x selected" % (res.device_ms, res.iterations, v[8]))
for k, nm in enumerate(names):
    print("  %-28s sum over wavefronts of the slowest warp: %8.3f ms   mean per x: %8.0f cycles"
          % (nm, v[k] / 1.965e6, (v[8 + k] / max(v[8], 1)) if k else 0))
o8 = (C.c_ulonglong * 8)()
g._lib.gmt_debug_counters(g._handle, o8)
w = list(o8)
print("leader: select phase %.3f ms = staging done after %.3f ms, own x done after %.3f ms, rest = waiting at the barrier"
      % (w[1] * 1e-6, w[5] * 1e-6, w[6] * 1e-6))
