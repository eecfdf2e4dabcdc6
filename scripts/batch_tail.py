"""Development aid: per-query durations of the cfg4 batch and what scheduling could gain at 512 queries per GPU."""
import ctypes as C, sys, heapq
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import bench, _lib, synthetic_world as sw
from gmt_planner import GMT
lib = _lib.load()
init, it, meta = bench.make_named_world("cfg4")
g = GMT(init)
g.set_obstacles(it['obstacles'], it['num_obs'])
starts, goals, _ = bench.batch_queries(g, meta)
q = int(sys.argv[1]) if len(sys.argv) > 1 else 512
s, gl = starts[:q], goals[:q]
for rep in range(3):
    res, _r = sw.plan_batch(g, s, gl)
be = (C.c_ulonglong * (2 * q))(); teams = C.c_int()
lib.gmt_debug_batch_times(g._handle, be, q, C.byref(teams))
b = np.array(be[0::2], np.float64) * 1e-6; e = np.array(be[1::2], np.float64) * 1e-6
d = e - b
st = init['states']
dist = np.hypot(st[s, 0] - st[gl, 0], st[s, 1] - st[gl, 1])
wf = np.array([r.iterations for r in res]); pairs = np.array([r.edge_pairs for r in res], np.float64)
T = teams.value
print("q %d teams %d: kernel %.2f ms; query duration mean %.3f ms, std %.3f, min %.3f, max %.3f" % (q, T, res[0].device_ms, d.mean(), d.std(), d.min(), d.max()))
print("correlation of duration with straight-line distance %.3f, with wavefronts %.3f, with edge checks %.3f" % (np.corrcoef(d, dist)[0, 1], np.corrcoef(d, wf)[0, 1], np.corrcoef(d, pairs)[0, 1]))
def lpt(key):
    order = np.argsort(-key, kind="stable"); h = [0.0] * T; heapq.heapify(h)
    for i in order: heapq.heappush(h, heapq.heappop(h) + d[i])
    return max(h)
print("ideal makespan (sum / teams) %.2f ms; greedy hand-out ordered by distance %.2f, by true duration %.2f, by edge checks %.2f, unordered %.2f"
      % (d.sum() / T, lpt(dist), lpt(d), lpt(pairs), lpt(np.zeros(q))))
print("actual span %.2f ms, busy fraction %.3f" % (e.max(), d.sum() / (e.max() * T)))
