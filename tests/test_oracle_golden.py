"""CPU: the oracle against the reference's golden vectors.

The pins, in order of strength:
 1. unit_tests.json / edge_vectors_*.npz were produced by the reference's own kernel source
    compiled for the host (tests/golden/make_golden.py); they reproduce the routes pictured in
    the reference's images/unitTest2.png, unitTest3.png and the vectors of SURVEY appendix B.
 2. oracle/gmt_oracle.c (libm flavour) must reproduce them BIT FOR BIT -- costs included.
 3. where oracle/_ref/libref_host.so exists (this container), port and reference are also
    compared bit for bit on fresh random pairs and whole random plans.
"""
import json
import os

import numpy as np
import pytest

from oracle.gmt_oracle import OracleGMT
from tests.golden.fixtures import FIXTURES

GOLD = os.path.join(os.path.dirname(__file__), "golden")
UNIT = json.load(open(os.path.join(GOLD, "unit_tests.json")))

# SURVEY.md appendix B (host-compiled reference): exit, iterations, route, parents
APPENDIX_B = {
    "unitTest1": (0, 4, [17, 11, 7, 1], [-1, -1, -1, 1, 1, 1, 1, 1, 1, 8, 7, 7, -1, -1, -1, 11, 11, 11]),
    "unitTest2": (0, 6, [16, 13, 10, 7, 4, 1], [5, -1, 3, 1, 1, 1, 4, 4, 4, 7, 7, 7, 10, 10, 10, 13, 13, 13]),
    "unitTest3": (0, 5, [16, 10, 8, 4, 1], [5, -1, 3, 1, 1, 1, 4, 4, 4, 6, 8, 8, 6, 6, 8, 10, 10, 10]),
}


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_golden_matches_survey_appendix_b(name):
    st, it, route, parent = APPENDIX_B[name]
    g = UNIT[name]
    assert (g["status"], g["iterations"], g["route"], g["parent"]) == (st, it, route, parent)


@pytest.mark.parametrize("name", sorted(FIXTURES))
@pytest.mark.parametrize("mode", ["kernels", "plan"])
def test_port_libm_reproduces_unit_fixtures_bit_exact(oracles, name, mode):
    init, it = FIXTURES[name]()
    gold = UNIT[name]
    port = oracles["libm"]
    if mode == "kernels":      # restated host loop over the port's kernel-level functions
        g = OracleGMT(init, port)
        route = g.run_step(it, iter_limit=20)
        status, iters, parent, cost, trace = g.status, g.iterations, g.parent, g.dev_cost, g.trace
    else:                      # the whole loop in C
        r = port.plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'],
                      it['radius'], it['threshold'], it['obstacles'], it['num_obs'], iter_limit=20)
        route, status, iters, parent, cost, trace = (r["route"], r["status"], r["iterations"], r["parent"],
                                                     r["cost"], r["trace"])
    assert status == gold["status"] and iters == gold["iterations"]
    assert [int(v) for v in route] == gold["route"]
    assert parent.tolist() == gold["parent"]
    assert cost.view(np.uint32).tolist() == gold["cost_bits"]
    assert len(trace) == len(gold["trace"])
    for a, b in zip(trace, gold["trace"]):
        for key in ("gSize", "ySize", "xSize"):
            assert int(a[key]) == int(b[key])
        for key in ("G", "Y", "X"):
            if key in b:
                assert np.asarray(a[key]).tolist() == b[key]


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_port_cuda_flavour_same_sets_and_parents(oracles, name):
    """The libdevice-emulating flavour may differ in the last bits of a cost, never in structure."""
    init, it = FIXTURES[name]()
    gold = UNIT[name]
    r = oracles["cuda"].plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], it['goal'],
                             it['radius'], it['threshold'], it['obstacles'], it['num_obs'], iter_limit=20)
    assert r["route"] == gold["route"] and r["parent"].tolist() == gold["parent"]
    gc = np.array(gold["cost_bits"], np.uint32).view(np.float32)
    fin = np.isfinite(gc)
    assert np.array_equal(np.isfinite(r["cost"]), fin)
    np.testing.assert_allclose(r["cost"][fin], gc[fin], rtol=1e-5)


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_port_libm_reproduces_edge_vectors_bit_exact(oracles, tag):
    e = np.load(os.path.join(GOLD, "edge_vectors_%s.npz" % tag))
    m = e["obstacles"].shape[0]
    cost4, bits, best = oracles["libm"].edge_costs(e["states"], e["y"], e["x"], float(e["radius"]),
                                                   e["obstacles"], [m])
    assert np.array_equal(bits, e["verdict_bits"])
    assert np.array_equal(cost4.view(np.uint32), e["cost4_bits"])
    assert np.array_equal(best.view(np.uint32), e["best_bits"])


@pytest.mark.parametrize("tag", ["r1", "r2"])
def test_port_cuda_flavour_edge_vectors_close(oracles, tag):
    e = np.load(os.path.join(GOLD, "edge_vectors_%s.npz" % tag))
    m = e["obstacles"].shape[0]
    cost4, bits, best = oracles["cuda"].edge_costs(e["states"], e["y"], e["x"], float(e["radius"]),
                                                   e["obstacles"], [m])
    gold = e["cost4_bits"].view(np.float32)
    assert np.array_equal(np.isnan(cost4), np.isnan(gold))
    ok = ~np.isnan(gold)
    rel = np.abs(cost4[ok] - gold[ok]) / np.maximum(np.abs(gold[ok]), 1e-3)
    # a wrap decision (theta within 1e-6 of zero, kernel.py:73) can flip between math libraries and
    # move a cost by 2*pi*r; everything else agrees to a few ulp
    assert np.mean(rel > 1e-5) < 2e-3
    assert np.mean(bits != e["verdict_bits"]) < 5e-3


def _random_world(rng, n, m, side):
    states = np.stack([rng.uniform(0, side, n), rng.uniform(0, side, n), rng.uniform(-np.pi, np.pi, n)],
                      1).astype(np.float32)
    c = rng.uniform(0, side, (max(m, 1), 2))
    h = rng.uniform(0.5, 3.0, (max(m, 1), 2))
    obs = np.concatenate([c - h, c + h], 1).astype(np.float32)
    return states, obs


@pytest.mark.parametrize("seed,n,m,side,radius", [(1, 400, 10, 30.0, 2.0), (2, 300, 0, 25.0, 1.0),
                                                  (3, 500, 25, 1400.0, 2.9), (4, 200, 6, 12.0, 0.5)])
def test_port_matches_reference_host_build_on_random_pairs(oracles, seed, n, m, side, radius):
    if oracles["ref"] is None:
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference)")
    rng = np.random.default_rng(seed)
    states, obs = _random_world(rng, n, m, side)
    y = rng.integers(0, n, 6000).astype(np.int32)
    x = rng.integers(0, n, 6000).astype(np.int32)
    a = oracles["ref"].edge_costs(states, y, x, radius, obs, [m])
    b = oracles["libm"].edge_costs(states, y, x, radius, obs, [m])
    assert np.array_equal(a[1], b[1])
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32))
    assert np.array_equal(a[2].view(np.uint32), b[2].view(np.uint32))


@pytest.mark.parametrize("seed,n,m", [(11, 350, 8), (12, 500, 0), (13, 420, 20)])
def test_port_plan_matches_reference_kernels_under_restated_loop(oracles, seed, n, m):
    """Whole plans: C loop over the port == Python restatement of GMT.run_step over the reference kernels."""
    if oracles["ref"] is None:
        pytest.skip("oracle/_ref/libref_host.so not built (needs /root/reference)")
    from oracle.world_oracle import build_neighbors
    rng = np.random.default_rng(seed)
    side = float(np.sqrt(n / 0.55))
    states, obs = _random_world(rng, n, m, side)
    nbr, num = build_neighbors(states)
    init = {'states': states, 'neighbors': nbr, 'num_neighbors': num, 'threadsPerBlock': 128}
    start = int(np.argmin(np.hypot(states[:, 0] - 0.1 * side, states[:, 1] - 0.1 * side)))
    goal = int(np.argmin(np.hypot(states[:, 0] - 0.9 * side, states[:, 1] - 0.9 * side)))
    it = {'start': start, 'goal': goal, 'radius': 2.0, 'threshold': 1.0, 'obstacles': obs,
          'num_obs': np.array([m], np.int32)}
    g = OracleGMT(init, oracles["ref"])
    route = g.run_step(it, iter_limit=60)
    r = oracles["libm"].plan(states, nbr, num, start, goal, 2.0, 1.0, obs, [m], iter_limit=60)
    assert (r["status"], r["iterations"]) == (g.status, g.iterations)
    assert r["parent"].tolist() == g.parent.tolist()
    assert np.array_equal(r["cost"].view(np.uint32), g.dev_cost.view(np.uint32))
    if g.status == 0 or g.parent[goal] != -1:
        assert r["route"] == [int(v) for v in route]
    for a, b in zip(r["trace"], g.trace):
        assert (a["gSize"], a["ySize"], a["xSize"]) == (b["gSize"], b["ySize"], b["xSize"])
        for key in ("G", "Y", "X"):
            if key in b:
                assert np.array_equal(a[key], b[key])


def test_euclidean_lower_bound_holds_with_the_planner_slack(oracles):
    """The CUDA planner skips y when cost[y] + 0.99999*|x-y| - slack > best.  That is exact only if
    every word length >= 0.99999*distance - slack; check it where rounding is worst (large
    coordinates) with the slack formula of gmt_kernels.cu (launch_plan)."""
    rng = np.random.default_rng(5)
    for side, radius in ((60.0, 2.0), (1400.0, 2.0), (1400.0, 2.9), (6000.0, 1.0)):
        n = 3000
        states = np.stack([rng.uniform(0, side, n), rng.uniform(0, side, n),
                           rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
        y = rng.integers(0, n, 40000).astype(np.int32)
        # neighbours within ~12 m: the regime where the bound actually prunes
        x = np.array([(yi + 1 + k) % n for yi, k in zip(y, rng.integers(0, 50, len(y)))], np.int32)
        order = np.argsort(states[:, 0])
        states = states[order]
        for flavour in ("libm", "cuda"):
            cost4, bits, _ = oracles[flavour].edge_costs(states, y, x, radius, np.zeros((1, 4), np.float32), [0])
            d = np.hypot(states[y, 0].astype(np.float64) - states[x, 0], states[y, 1].astype(np.float64) - states[x, 1])
            slack = (np.abs(states[:, :2]).max() + 16.0) / 65536.0
            wc = np.nanmin(np.where(np.isnan(cost4), np.inf, cost4), axis=1)
            ok = np.isfinite(wc) & (y != x)
            margin = wc[ok] - (0.99999 * d[ok] - slack)
            assert margin.min() > 0.25 * slack, (side, radius, flavour, margin.min(), slack)
