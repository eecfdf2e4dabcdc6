"""CPU: the extension-mode definitions of the oracle (six Dubins words, oriented rectangles).

The reference has neither ("parity unpinned"), so oracle/gmt_oracle.c is their
definition and what pins it here is mathematics: the shortest of the six words must be the closed-form
Dubins distance, the 150 sample points must trace a continuous curve from start to end, and an
oriented rectangle at angle 0 must give the verdicts of the equivalent corner box.
"""
import math

import numpy as np
import pytest

NAMES = ["RSR", "LSL", "LSR", "RSL", "RLR", "LRL"]


@pytest.fixture()
def ext(oracles):
    """Both flavours of the oracle with a guaranteed return to reference-parity mode."""
    try:
        yield oracles
    finally:
        oracles["libm"].set_extensions(4, False)
        oracles["cuda"].set_extensions(4, False)


def _mod2pi(a):
    return a - 2 * math.pi * math.floor(a / (2 * math.pi))


def closed_form_dubins(q0, q1, r):
    """All six word lengths from the classical closed forms (Shkel & Lumelsky 2001), double precision."""
    dx, dy = q1[0] - q0[0], q1[1] - q0[1]
    d = math.hypot(dx, dy) / r
    th = _mod2pi(math.atan2(dy, dx))
    a, b = _mod2pi(q0[2] - th), _mod2pi(q1[2] - th)
    sa, sb, ca, cb, cab = math.sin(a), math.sin(b), math.cos(a), math.cos(b), math.cos(a - b)
    out = {}
    t = 2 + d * d - 2 * cab + 2 * d * (sa - sb)
    if t >= 0:
        tmp = math.atan2(cb - ca, d + sa - sb)
        out["LSL"] = (_mod2pi(-a + tmp) + math.sqrt(t) + _mod2pi(b - tmp)) * r
    t = 2 + d * d - 2 * cab + 2 * d * (sb - sa)
    if t >= 0:
        tmp = math.atan2(ca - cb, d - sa + sb)
        out["RSR"] = (_mod2pi(a - tmp) + math.sqrt(t) + _mod2pi(-b + tmp)) * r
    t = -2 + d * d + 2 * cab + 2 * d * (sa + sb)
    if t >= 0:
        p = math.sqrt(t)
        tmp = math.atan2(-ca - cb, d + sa + sb) - math.atan2(-2.0, p)
        out["LSR"] = (_mod2pi(-a + tmp) + p + _mod2pi(-_mod2pi(b) + tmp)) * r
    t = d * d - 2 + 2 * cab - 2 * d * (sa + sb)
    if t >= 0:
        p = math.sqrt(t)
        tmp = math.atan2(ca + cb, d - sa - sb) - math.atan2(2.0, p)
        out["RSL"] = (_mod2pi(a - tmp) + p + _mod2pi(b - tmp)) * r
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sa - sb)) / 8.0
    if abs(tmp) <= 1:
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(a - math.atan2(ca - cb, d - sa + sb) + _mod2pi(p / 2.0))
        out["RLR"] = (t + p + _mod2pi(a - b - t + _mod2pi(p))) * r
    tmp = (6.0 - d * d + 2 * cab + 2 * d * (sb - sa)) / 8.0
    if abs(tmp) <= 1:
        p = _mod2pi(2 * math.pi - math.acos(tmp))
        t = _mod2pi(-a - math.atan2(ca - cb, d + sa - sb) + p / 2.0)
        out["LRL"] = (t + p + _mod2pi(_mod2pi(b) - a - t + _mod2pi(p))) * r
    return out


def _random_pair(rng, side):
    q = np.array([rng.uniform(0, side), rng.uniform(0, side), rng.uniform(-math.pi, math.pi)], np.float32)
    return q


@pytest.mark.parametrize("flavour", ["libm", "cuda"])
def test_shortest_of_six_words_is_the_dubins_distance(ext, flavour):
    o = ext[flavour].set_extensions(6, False)
    rng = np.random.default_rng(11)
    n_ccc_best = 0
    for k in range(3000):
        q0, q1 = _random_pair(rng, 9.0), _random_pair(rng, 9.0)
        r = int(rng.integers(1, 4))
        ref = closed_form_dubins(q0.astype(np.float64), q1.astype(np.float64), float(r))
        got = {}
        for w in range(6):
            fl, c, xs, ys = o.word(w, q0, q1, r)
            if fl & 1:
                got[NAMES[w]] = c
        # every word: same length as the closed form (tolerance: float32 trig on coordinates ~10)
        for name in ("RLR", "LRL"):
            if name in got and name in ref:
                assert abs(got[name] - ref[name]) <= 2e-4 * max(1.0, ref[name]), (k, name, q0, q1, r)
        best, rbest = min(got.values()), min(ref.values())
        if abs(best - rbest) > 2e-4 * max(1.0, rbest):
            # the reference's CSC words wrap angles within 1e-6 of zero the other way (kernel.py:73): allow a
            # full-turn difference only there
            assert abs(abs(best - rbest) - 2 * math.pi * r) < 1e-3, (k, q0, q1, r, got, ref)
        if min(got, key=got.get) in ("RLR", "LRL"):
            n_ccc_best += 1
    assert n_ccc_best > 30          # the CCC words do matter at these distances


@pytest.mark.parametrize("flavour", ["libm", "cuda"])
def test_ccc_sample_points_trace_the_curve(ext, flavour):
    o = ext[flavour].set_extensions(6, False)
    rng = np.random.default_rng(12)
    seen = 0
    for k in range(400):
        q0, q1 = _random_pair(rng, 6.0), _random_pair(rng, 6.0)
        r = 2
        for w in (4, 5):
            fl, c, xs, ys = o.word(w, q0, q1, r)
            if not fl & 1:
                continue
            seen += 1
            assert abs(xs[0] - q0[0]) < 1e-3 and abs(ys[0] - q0[1]) < 1e-3
            assert abs(xs[149] - q1[0]) < 1e-3 and abs(ys[149] - q1[1]) < 1e-3
            # the arcs join: point 49 = point 50, point 99 = point 100
            assert math.hypot(xs[49] - xs[50], ys[49] - ys[50]) < 1e-3
            assert math.hypot(xs[99] - xs[100], ys[99] - ys[100]) < 1e-3
            # polyline length approaches the word length from below
            poly = float(np.sum(np.hypot(np.diff(xs.astype(np.float64)), np.diff(ys.astype(np.float64)))))
            assert poly <= c + 1e-3 and poly >= 0.995 * c - 1e-3
            # heading at the first step = start heading, at the last step = end heading
            h0 = math.atan2(ys[1] - ys[0], xs[1] - xs[0]) if c > 1e-2 and math.hypot(xs[1] - xs[0], ys[1] - ys[0]) > 1e-4 else None
            if h0 is not None:
                assert abs(_mod2pi(h0 - q0[2] + math.pi) - math.pi) < 0.15
    assert seen > 200


def test_existence_rule_and_default_mode_is_untouched(ext):
    o = ext["libm"]
    q0 = np.array([0, 0, 0], np.float32)
    far = np.array([20, 0, 0], np.float32)          # circle centres 20 m apart > 4 r
    o.set_extensions(6, False)
    assert o.word(4, q0, far, 2)[0] == 0 and o.word(5, q0, far, 2)[0] == 0
    same = o.word(4, q0, q0, 2)
    assert same[0] == 0                             # coincident centres: no CCC word (d = 0)
    c6, b6, best6 = o.edge_costs(np.stack([q0, far]), [0], [1], 2.0, [], [0])
    assert c6.shape == (1, 6) and np.isnan(c6[0, 4]) and np.isnan(c6[0, 5])
    o.set_extensions(4, False)
    c4, b4, best4 = o.edge_costs(np.stack([q0, far]), [0], [1], 2.0, [], [0])
    assert c4.shape == (1, 4) and np.array_equal(c4[0], c6[0, :4]) and best4[0] == best6[0]


def test_oriented_rectangles(ext):
    o = ext["libm"]
    rng = np.random.default_rng(13)
    n = 300
    states = np.stack([rng.uniform(0, 30, n), rng.uniform(0, 30, n), rng.uniform(-np.pi, np.pi, n)], 1).astype(np.float32)
    y = rng.integers(0, n, 3000).astype(np.int32)
    x = ((y + 1 + rng.integers(0, n - 1, 3000)) % n).astype(np.int32)
    # angle 0, coordinates exact in float: identical verdicts to the equivalent corner boxes
    c = rng.integers(2, 28, (12, 2)).astype(np.float32)
    hh = rng.integers(1, 4, (12, 2)).astype(np.float32)
    corner = np.concatenate([c - hh, c + hh], 1).astype(np.float32)
    rect = np.concatenate([c, hh, np.ones((12, 1), np.float32), np.zeros((12, 1), np.float32)], 1).astype(np.float32)
    o.set_extensions(4, False)
    c4, bits_corner, best_corner = o.edge_costs(states, y, x, 2.0, corner, [12])
    o.set_extensions(4, True)
    c4o, bits_rect, best_rect = o.edge_costs(states, y, x, 2.0, rect, [12])
    assert np.array_equal(bits_corner, bits_rect) and np.array_equal(best_corner, best_rect)
    assert ((bits_rect >> 4) & 15).any()
    # a thin rectangle turned by 45 degrees blocks the diagonal it lies across, not the one it lies along
    s = math.sqrt(0.5)
    wall = np.array([[10, 10, 6.0, 0.2, s, s]], np.float32)         # long axis along (1, 1)
    a = np.array([6, 14, -math.pi / 4], np.float32)                 # heading (1, -1): crosses the wall
    b = np.array([14, 6, -math.pi / 4], np.float32)
    fl, cost, xs, ys = o.word(0, a, b, 2, wall, 1)
    assert fl == 3
    a2 = np.array([5, 7, math.pi / 4], np.float32)                  # parallel to the wall, 1.4 m beside it
    b2 = np.array([13, 15, math.pi / 4], np.float32)
    fl2 = o.word(0, a2, b2, 2, wall, 1)[0]
    assert fl2 == 1
    # the bounding box of that wall would have blocked the second path: orientation matters
    o.set_extensions(4, False)
    bb = np.array([[10 - 6.2 * s * 1.0 - 0.2, 10 - 6.2 * s - 0.2, 10 + 6.2 * s + 0.2, 10 + 6.2 * s + 0.2]], np.float32)
    assert o.word(0, a2, b2, 2, bb, 1)[0] == 3
