"""Development aid: throughput of gmt_plan_batch (BASELINE config 4 shape) vs back-to-back single plans."""
import sys, time, ctypes as C, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import _lib, synthetic_world as sw
from gmt_planner import GMT
q = int(sys.argv[1]) if len(sys.argv) > 1 else 512
init, it, meta = sw.make_world(2)
g = GMT(init)
lib = g._lib
obs = np.ascontiguousarray(it['obstacles'], np.float32)
_lib.check(lib.gmt_set_obstacles(g._handle, obs.ctypes.data, int(it['num_obs'][0])), "obs")
starts, goals = sw.make_reachable_queries(g, meta["n"], q, 123)
res = (_lib.GmtResult * q)()
stride = 256
routes = np.zeros((q, stride), np.int32)
for rep in range(2):
    t0 = time.perf_counter()
    _lib.check(lib.gmt_plan_batch(g._handle, starts.ctypes.data, goals.ctypes.data, q, 2.0, 1.0, 1000, res,
                                  routes.ctypes.data, stride), "batch")
    dt = time.perf_counter() - t0
    print('batch of %d: %.1f ms wall, device %.1f ms -> %.0f plans/s; goal reached %d, mean wavefronts %.1f'
          % (q, dt * 1e3, res[0].device_ms, q / (res[0].device_ms * 1e-3), sum(r.status == 0 for r in res),
             np.mean([r.iterations for r in res])))
single = _lib.GmtResult()
t0 = time.perf_counter()
ms = 0.0
nq = min(q, 64)
ok = 0
for k in range(nq):
    _lib.check(lib.gmt_plan_resident(g._handle, int(starts[k]), int(goals[k]), 2.0, 1.0, 1000, C.byref(single)), "single")
    ms += single.device_ms
    same = (single.status, single.iterations, single.route_len, np.float32(single.goal_cost).view(np.uint32)) == \
        (res[k].status, res[k].iterations, res[k].route_len, np.float32(res[k].goal_cost).view(np.uint32))
    ok += bool(same)
print('%d single plans: %.1f ms device -> %.0f plans/s; results identical to the batch: %d/%d'
      % (nq, ms, nq / (ms * 1e-3), ok, nq))
