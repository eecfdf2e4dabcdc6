"""Development aid: extension mode (six words, oriented rectangles) on the cfg2 world: plan time and edge-kernel rate."""
import sys, ctypes as C
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import _lib, synthetic_world as sw
from gmt_planner import GMT
init, it, meta = sw.make_world(2)
it = dict(it, goal=sw.pick_reachable_goal(init, it, meta["side"]))
n = meta["n"]
rng = np.random.default_rng(7)
nbr, num = init['neighbors'], init['num_neighbors']
rows = np.repeat(np.arange(n, dtype=np.int32), num)
sel = rng.integers(0, len(nbr), 1 << 18)
ey, ex = np.ascontiguousarray(rows[sel]), np.ascontiguousarray(nbr[sel].astype(np.int32))
obs = np.ascontiguousarray(it['obstacles'], np.float32)
lo, hi = np.minimum(obs[:, :2], obs[:, 2:]), np.maximum(obs[:, :2], obs[:, 2:])
rect5 = np.concatenate([(lo + hi) / 2, (hi - lo) / 2, rng.uniform(-np.pi, np.pi, (len(obs), 1))], 1).astype(np.float32)
for words in (4, 6):
    for oriented in (False, True):
        g = GMT(init, words=words)
        q = dict(it, obstacles=rect5, obstacle_format='oriented') if oriented else it
        for _ in range(5):
            g.run_step(q, iter_limit=1000)
        ms = [0.0] * 20
        for k in range(20):
            g.run_step(q, iter_limit=1000); ms[k] = g.last_result.device_ms
        r = g.last_result
        t = C.c_float(); best = 1e9
        for _ in range(4):
            _lib.check(g._lib.gmt_edge_costs(g._handle, ey.ctypes.data, ex.ctypes.data, len(ey), 2.0, None, None, None, C.byref(t)), "edge")
            best = min(best, t.value)
        print("words %d, %s obstacles: plan %.3f ms (status %d, %d wavefronts, %d pairs evaluated, %d words swept); "
              "edge kernel %.0f M edge checks/s (%d words each)"
              % (words, "oriented" if oriented else "corner-box", float(np.median(ms)), r.status, r.iterations, r.edge_evals,
                 r.word_checks, len(ey) / best * 1e-3, words))
