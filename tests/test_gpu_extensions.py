"""GPU: extension mode of the CUDA path (six Dubins words, oriented rectangle obstacles).

BASELINE.json's north star names both; the reference implements neither, so their definition is
oracle/gmt_oracle.c (ccc_impl, check_col_obb; pinned against the closed-form Dubins distance in
tests/test_oracle_extensions.py) and parity here means: the CUDA path reproduces that oracle the same
way it reproduces the reference in parity mode -- verdicts, sets and parents bit-exact, costs within
1e-5 -- and switching the extensions off again gives the reference results back.
"""
import ctypes as C

import numpy as np
import pytest

from tests.test_gpu_parity import GMT, RTOL, _compare_plan, _ulp_diff, _world, report  # noqa: F401 (GMT: fixture)

pytestmark = pytest.mark.gpu


@pytest.fixture()
def ext(oracles):
    try:
        yield oracles["cuda"]
    finally:
        oracles["libm"].set_extensions(4, False)
        oracles["cuda"].set_extensions(4, False)


def _rects(rng, m, side):
    """m oriented rectangles [cx, cy, half_x, half_y, angle]."""
    return np.stack([rng.uniform(0, side, m), rng.uniform(0, side, m), rng.uniform(0.5, 4.0, m),
                     rng.uniform(0.3, 2.0, m), rng.uniform(-np.pi, np.pi, m)], 1).astype(np.float32)


def _edge_ext(gmt_lib, states, y, x, radius, words, corner=None, rects6=None):
    import _lib
    h = C.c_void_p()
    n = states.shape[0]
    num = np.zeros(n, np.int32)
    _lib.check(gmt_lib.gmt_create(states.ctypes.data, n, None, num.ctypes.data, 0, C.byref(h)), "gmt_create")
    try:
        _lib.check(gmt_lib.gmt_set_words(h, words), "gmt_set_words")
        if rects6 is not None:
            _lib.check(gmt_lib.gmt_set_obstacles_oriented(h, rects6.ctypes.data if len(rects6) else None, len(rects6)),
                       "gmt_set_obstacles_oriented")
        elif corner is not None:
            _lib.check(gmt_lib.gmt_set_obstacles(h, corner.ctypes.data if len(corner) else None, len(corner)),
                       "gmt_set_obstacles")
        P = len(y)
        cost = np.zeros((P, words), np.float32)
        bits = np.zeros(P, np.uint32)
        best = np.zeros(P, np.float32)
        _lib.check(gmt_lib.gmt_edge_costs(h, y.ctypes.data, x.ctypes.data, P, float(radius), cost.ctypes.data,
                                          bits.ctypes.data, best.ctypes.data, None), "gmt_edge_costs")
    finally:
        gmt_lib.gmt_destroy(h)
    return cost, bits, best


@pytest.mark.parametrize("words,oriented", [(6, False), (4, True), (6, True)])
def test_edge_kernel_extension_modes_match_oracle(gmt_lib, ext, words, oriented):
    from gmt_planner import oriented_records
    rng = np.random.default_rng(100 + words + int(oriented))
    side, n, m, radius = 45.0, 1200, 40, 2.0
    states = np.stack([rng.uniform(0, side, n), rng.uniform(0, side, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    P = 30000
    y = rng.integers(0, n, P).astype(np.int32)
    x = ((y + 1 + rng.integers(0, n - 1, P)) % n).astype(np.int32)
    if oriented:
        rec = oriented_records(_rects(rng, m, side))
        obs_oracle, corner = rec, None
    else:
        c, hh = rng.uniform(0, side, (m, 2)), rng.uniform(0.5, 3.0, (m, 2))
        corner = np.concatenate([c + hh, c - hh], 1).astype(np.float32)
        rec, obs_oracle = None, corner
    cost, bits, best = _edge_ext(gmt_lib, states, y, x, radius, words, corner=corner, rects6=rec)
    ext.set_extensions(words, oriented)
    ocost, obits, obest = ext.edge_costs(states, y, x, radius, obs_oracle, [m])
    assert np.array_equal(np.isnan(cost), np.isnan(ocost))
    ok = ~np.isnan(ocost)
    ulp = _ulp_diff(cost[ok], ocost[ok])
    cb = 4 if words == 4 else 8
    n_ccc = int(ok[:, 4:].sum()) if words == 6 else 0
    report("extension edge kernel (words %d, oriented %s): %d pairs, %d CCC words exist; cost words differing from the "
           "oracle: %d (max %d ulp); verdict words differing: %d; colliding words: %d"
           % (words, oriented, P, n_ccc, int((ulp > 0).sum()), int(ulp.max()), int((bits != obits).sum()),
              int(sum(bin(int(v) >> cb).count("1") for v in bits))))
    # extension mode has no reference pin (the C restatement is its definition, evaluated with glibc-emulating
    # arithmetic): verdicts must be identical; lengths may differ by 1 ulp on a handful of words (logged above)
    assert int(ulp.max()) <= 1 and int((ulp > 0).sum()) <= 32
    assert int((bits != obits).sum()) == 0
    same = bits == obits
    assert np.array_equal(best[same].view(np.uint32), obest[same].view(np.uint32)) or \
        np.mean(np.abs(best[same] - obest[same]) > RTOL * np.abs(obest[same])) < 2e-3
    if words == 6:
        assert n_ccc > 1000
    assert (bits >> cb).any()


@pytest.mark.parametrize("words,oriented", [(6, False), (4, True), (6, True)])
def test_plan_extension_modes_match_oracle(GMT, ext, words, oriented):
    from gmt_planner import oriented_records
    init, it, meta = _world(None, 1500, 30, 29, reachable=False)
    rng = np.random.default_rng(7)
    if oriented:
        rect5 = _rects(rng, 30, meta["side"])
        # keep the start clear
        s = init['states'][it['start']]
        rect5 = rect5[np.hypot(rect5[:, 0] - s[0], rect5[:, 1] - s[1]) > 7.0]
        it = dict(it, obstacles=rect5, num_obs=np.array([len(rect5)], np.int32), obstacle_format='oriented')
        obs_oracle = oriented_records(rect5)
    else:
        obs_oracle = it['obstacles']
    g = GMT(init, words=words)
    # a goal this mode can reach: farthest member of any wavefront of a goal-less run
    g.run_step(dict(it, goal=-1), iter_limit=60)
    goal = g.last_result.farthest_frontier
    assert goal >= 0
    it = dict(it, goal=int(goal))
    route = g.run_step(it, iter_limit=60, debug=True)
    ext.set_extensions(words, oriented)
    r = ext.plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'], it['radius'],
                 it['threshold'], obs_oracle, it['num_obs'], iter_limit=60)
    _compare_plan(g, r["status"], r["iterations"], r["parent"], r["cost"], r["trace"], init['states'],
                  "extension plan words=%d oriented=%s" % (words, oriented))
    assert g.status == 0 and len(route) == len(r["route"])
    if words == 6:
        # the extra words shorten some connections: the tree is not the 4-word tree
        ext.set_extensions(4, oriented)
        r4 = ext.plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'], it['radius'],
                      it['threshold'], obs_oracle, it['num_obs'], iter_limit=60, trace=False)
        fin = np.isfinite(r4["cost"]) & np.isfinite(r["cost"])
        assert (r["cost"][fin] <= r4["cost"][fin] * (1 + 1e-6) + 1e-6).mean() > 0.5
        assert (r["cost"][fin] < r4["cost"][fin] - 1e-3).any()
    # trajectory output in this mode
    ft, word, arcs = g.route_edges()
    oft, oword, oarcs = ext.set_extensions(words, oriented).route_edges(init['states'], route, it['radius'], obs_oracle,
                                                                        it['num_obs'])
    assert np.array_equal(ft, oft) and np.array_equal(word, oword)
    np.testing.assert_allclose(arcs, oarcs, rtol=RTOL, atol=1e-6)
    report("extension plan (words %d, oriented %s): status %d after %d wavefronts, %d nodes connected, route of %d "
           "nodes uses words %s" % (words, oriented, g.status, g.iterations, int((g.parent >= 0).sum()), len(route),
                                   sorted(set(int(w) for w in word))))


def test_switching_extensions_off_restores_the_reference_results(GMT, gmt_lib):
    import _lib
    init, it, meta = _world(None, 1500, 30, 29)
    ref = GMT(init)
    route = list(ref.run_step(it, iter_limit=200))
    parent, cost = ref.parent.copy(), ref.dev_cost.copy()
    g = GMT(init, words=6)
    rect5 = _rects(np.random.default_rng(3), 10, meta["side"])
    g.run_step(dict(it, obstacles=rect5, num_obs=np.array([10], np.int32), obstacle_format='oriented'), iter_limit=50)
    g.set_words(4)
    assert list(g.run_step(it, iter_limit=200)) == route
    assert np.array_equal(g.parent, parent) and np.array_equal(g.dev_cost.view(np.uint32), cost.view(np.uint32))
    with pytest.raises(_lib.GmtError):
        g.set_words(5)
    bad = np.array([[0, 0, 1, 1, 0, 0]], np.float32)               # (cos, sin) = 0
    assert gmt_lib.gmt_set_obstacles_oriented(g._handle, bad.ctypes.data, 1) != 0


def test_batch_in_extension_mode_equals_single_plans(GMT, gmt_lib):
    import _lib
    import synthetic_world as sw
    from gmt_planner import oriented_records
    init, it, meta = _world(None, 1500, 30, 29)
    rect5 = _rects(np.random.default_rng(5), 25, meta["side"])
    g = GMT(init, words=6)
    itx = dict(it, obstacles=rect5, num_obs=np.array([25], np.int32), obstacle_format='oriented')
    starts = np.random.default_rng(9).integers(0, g.n, 12).astype(np.int32)
    rec = oriented_records(rect5)
    _lib.check(gmt_lib.gmt_set_obstacles_oriented(g._handle, rec.ctypes.data, 25), "gmt_set_obstacles_oriented")
    res0, _ = sw.plan_batch(g, starts, np.full(12, -1, np.int32), iter_limit=80)
    goals = np.array([r.farthest_frontier if r.farthest_frontier >= 0 else int(s) for r, s in zip(res0, starts)], np.int32)
    res, routes = sw.plan_batch(g, starts, goals, iter_limit=80, route_stride=128)
    for k in range(12):
        single = list(GMT(init, words=6).run_step(dict(itx, start=int(starts[k]), goal=int(goals[k])), iter_limit=80))
        assert res[k].status == 0
        assert [int(v) for v in routes[k, :res[k].route_len]] == single
