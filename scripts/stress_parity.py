"""Randomised parity stress: random worlds / radii / kernel instantiations / extension modes, CUDA path against
the CPU oracle (cuda flavour): status, iteration count, per-wavefront set sizes, parents and costs.
usage: stress_parity.py [seconds] [seed]"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import synthetic_world as sw
from gmt_planner import GMT, oriented_records
from oracle import build_oracle
from oracle.gmt_oracle import OraclePort
build_oracle.build_port()
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
oracle = OraclePort("cuda")
t0, cases, nodes, worst, batches = time.time(), 0, 0, 0.0, 0
while time.time() - t0 < budget:
    n = int(rng.integers(300, 3500)); m = int(rng.integers(0, 50)); seed = int(rng.integers(0, 1 << 30))
    mode = "lattice" if rng.random() < 0.2 else "uniform"
    if mode == "lattice":
        n = 7 * max(n // 7, 50)
    radius = float(rng.choice([1.0, 2.0, 2.0, 2.9, 3.0]))
    words = 6 if rng.random() < 0.3 else 4
    oriented = rng.random() < 0.25
    os.environ["GMT_WIDE"] = str(int(rng.random() < 0.5))
    if rng.random() < 0.4:
        os.environ["GMT_SELECT_K"] = str(int(rng.choice([2, 4, 8, 16])))
    else:
        os.environ.pop("GMT_SELECT_K", None)
    if rng.random() < 0.2:
        os.environ["GMT_PAIR_CAP"] = str(int(rng.choice([128, 1024])))
    else:
        os.environ.pop("GMT_PAIR_CAP", None)
    os.environ["GMT_MASK_MIN_Y"] = str(int(rng.choice([0, 512])))
    init, it, meta = sw.make_world(n=n, m=m, seed=seed, mode=mode)
    it = dict(it, radius=radius)
    obs_oracle = it['obstacles']
    if oriented and m > 0:
        side = meta["side"]
        rect5 = np.stack([rng.uniform(0, side, m), rng.uniform(0, side, m), rng.uniform(0.5, 4.0, m), rng.uniform(0.3, 2.0, m),
                          rng.uniform(-np.pi, np.pi, m)], 1).astype(np.float32)
        s0 = init['states'][it['start']]
        rect5 = rect5[np.hypot(rect5[:, 0] - s0[0], rect5[:, 1] - s0[1]) > 7.0]
        it = dict(it, obstacles=rect5 if len(rect5) else np.zeros((1, 5), np.float32), num_obs=np.array([len(rect5)], np.int32),
                  obstacle_format='oriented')
        obs_oracle = oriented_records(rect5) if len(rect5) else np.zeros((1, 6), np.float32)
    else:
        oriented = False
    g = GMT(init, words=words)
    g.run_step(dict(it, goal=-1), iter_limit=60)
    far = g.last_result.farthest_frontier
    it = dict(it, goal=int(far) if far >= 0 and rng.random() < 0.8 else int(rng.integers(0, g.n)))
    limit = int(rng.choice([30, 80]))
    route = list(g.run_step(it, iter_limit=limit, debug=True))
    oracle.set_extensions(words, oriented)
    r = oracle.plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'], radius,
                    it['threshold'], obs_oracle, it['num_obs'], iter_limit=limit)
    tag = "n=%d m=%d seed=%d %s r=%g words=%d oriented=%s env=%s" % (
        n, m, seed, mode, radius, words, oriented, {k: v for k, v in os.environ.items() if k.startswith("GMT_")})
    assert (g.status, g.iterations) == (r["status"], r["iterations"]), tag
    for a, b in zip(g.trace, r["trace"]):
        assert (a["gSize"], a["ySize"], a["xSize"]) == (int(b["gSize"]), int(b["ySize"]), int(b["xSize"])), tag
        for key in ("G", "Y", "X"):
            if key in b:
                assert np.array_equal(np.asarray(a[key]), np.asarray(b[key])), (tag, key)
    fin = np.isfinite(r["cost"])
    assert np.array_equal(np.isfinite(g.dev_cost), fin), tag
    rel = np.abs(g.dev_cost[fin] - r["cost"][fin]) / np.maximum(np.abs(r["cost"][fin]), 1e-3)
    assert (rel <= 1e-5).all(), (tag, float(rel.max()))
    worst = max(worst, float(rel.max()) if fin.any() else 0.0)
    diff = np.flatnonzero(g.parent != r["parent"])
    for i in diff:
        assert abs(g.dev_cost[i] - r["cost"][i]) <= 1e-5 * max(1.0, abs(r["cost"][i])), (tag, int(i))
    if g.status == 0 or r["route"]:
        assert route == r["route"] or len(diff) > 0, tag
    cases += 1; nodes += int(fin.sum())
    # the batched entry on the same world: 6 random queries in one launch == 6 single plans (corner boxes only)
    if not oriented and rng.random() < 0.5:
        import _lib
        obs = np.ascontiguousarray(it['obstacles'], np.float32)
        mm = int(it['num_obs'][0])
        _lib.check(g._lib.gmt_set_obstacles(g._handle, obs.ctypes.data if mm else None, mm), "set")
        qs = rng.integers(0, g.n, 6).astype(np.int32)
        res0, _ = sw.plan_batch(g, qs, np.full(6, -1, np.int32), radius=radius, iter_limit=limit)
        qg = np.array([r_.farthest_frontier if r_.farthest_frontier >= 0 else int(s_) for r_, s_ in zip(res0, qs)], np.int32)
        res, routes = sw.plan_batch(g, qs, qg, radius=radius, iter_limit=limit, route_stride=256)
        for k in range(6):
            single = GMT(init, words=words)
            rs = list(single.run_step(dict(it, start=int(qs[k]), goal=int(qg[k])), iter_limit=limit))
            assert (res[k].status, res[k].iterations) == (single.status, single.iterations), (tag, "batch", k)
            if res[k].status == 0:
                assert [int(v) for v in routes[k, :res[k].route_len]] == rs, (tag, "batch route", k)
        batches += 1
oracle.set_extensions(4, False)
print("stress parity: %d random plans (worlds of 300-3500 states, 0-50 obstacles, radii 1-3, 4 / 6 words, corner / oriented boxes, "
      "lean / wide kernels, warp groups, tiny pair lists, masks on / off) in %.0f s: every status, iteration count, wavefront set and "
      "parent equal to the CPU oracle; %d connected nodes, largest relative cost difference %.2e; %d batches of 6 queries "
      "equal to single plans" % (cases, time.time() - t0, nodes, worst, batches))
