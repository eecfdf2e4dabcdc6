"""Development aid: per-wavefront device time and work counters of one plan (run under gpurun)."""
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import _lib, synthetic_world as sw
from gmt_planner import GMT
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 2
if cfg == 0:      # explicit sizes: dbg_phases.py 0 <n> <m>
    init, it, meta = sw.make_world(n=int(sys.argv[2]), m=int(sys.argv[3]), seed=23)
    it = dict(it, goal=sw.pick_reachable_goal(init, it, meta["side"]))
else:
    init, it, meta = sw.make_world(cfg)
    if len(sys.argv) > 2:
        it = dict(it, goal=int(sys.argv[2]))
    else:
        it = dict(it, goal=sw.pick_reachable_goal(init, it, meta["side"]))
g = GMT(init)
for rep in range(3):
    route = g.run_step(it, iter_limit=1000, time=True)
res = g.last_result
print('goal', it['goal'], 'status', res.status, 'iters', res.iterations, 'ms', res.device_ms, 'pairs', res.edge_pairs,
      'evals', res.edge_evals, 'checks', res.word_checks)
el = np.array(g.time_data['elapsed'][-(res.iterations - 1):]) * 1e3
print('per-iteration ms: sum %.3f max %.3f median %.4f' % (el.sum(), el.max(), np.median(el)))
print('sizes (G,Y,X):', [(e['gSize'], e['ySize'], e['xSize']) for e in g.trace])
print('iter ms', np.round(el, 3).tolist())
out = (C.c_ulonglong * 8)()
g._lib.gmt_debug_counters(g._handle, out)
v = list(out)
print('leader time: frontier+barrier %.3f ms, select+barrier %.3f ms, (unused) %.3f ms, connect+barrier %.3f ms; pairs %d'
      % (v[0] * 1e-6, v[1] * 1e-6, v[2] * 1e-6, v[4] * 1e-6, v[3]))
for bps, thr in ((1, 128), (1, 256), (1, 512), (2, 256), (4, 128), (8, 64)):
    us = C.c_float()
    g._lib.gmt_debug_barrier_us(g._handle, bps, thr, 2000, C.byref(us))
    print('grid barrier: %d blocks/SM x %d threads: %.2f us' % (bps, thr, us.value))
