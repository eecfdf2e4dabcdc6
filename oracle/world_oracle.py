"""oracle/world_oracle.py -- TEST INFRASTRUCTURE ONLY.

NumPy restatement of the roadmap builder that sits in front of the planner:

* ``philox4x32_10``      -- the counter-based generator the CUDA sampler uses (Random123's
                            published Philox4x32-10; pinned by its known-answer vectors in
                            tests/test_world_oracle.py).
* ``sample_states``      -- synthetic replacement of CARLA's ``generate_waypoints`` +
                            yaw up-sampling (reference: carla/cuda_agent.py:56-113).
* ``build_neighbors``    -- the reference's neighbour rule (cuda_agent.py:71-91), restated on
                            states: j is a neighbour of i iff their locations differ and the
                            float32 distance sqrtf(dx*dx + dy*dy) <= disk_radius (compared as
                            double, inclusive); every yaw state of a neighbouring location is
                            listed, in ascending state index.  CARLA's Location.distance is
                            binary-only, so this restatement IS the definition ("parity unpinned",
                            SURVEY.md section 8c).

The product's CUDA sampler / cell-grid neighbour search (gmt_sample_states,
gmt_build_neighbors in include/gmt_c_api.h) are checked against these bit for bit.
"""
import math

import numpy as np

PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(counter, key):
    """counter: uint32[..., 4], key: (k0, k1).  Returns uint32[..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c[0]
        p1 = PHILOX_M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & _MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & _MASK
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def _u01(r):
    """24-bit uniform in [0,1): (r >> 8) * 2^-24, exact in float32."""
    return (r >> np.uint32(8)).astype(np.float32) * np.float32(2.0 ** -24)


def sample_states(seed, n, side, mode="uniform", first=0):
    """n Dubins states (x, y, theta) as float32[n,3].

    uniform: state i draws Philox(counter=(first+i,0,0,0), key=(seed, 0x474D5421)):
             x = u0*side, y = u1*side, theta = u2*2pi - pi   (each op rounded to float32).
    lattice: the reference's layout (cuda_agent.py:95-109) on a 4 m grid: location l = i // 7 at
             ((l % W)*4, (l // W)*4) with W = floor(side/4)+1, base yaw 0 and the 7 yaw offsets
             k*45 deg, k != 4, wrapped the way cuda_agent.py:102-107 wraps them.
    """
    n = int(n)
    if mode == "uniform":
        ctr = np.zeros((n, 4), np.uint32)
        ctr[:, 0] = (np.arange(n, dtype=np.uint64) + np.uint64(first)).astype(np.uint32)
        r = philox4x32_10(ctr, (int(seed) & 0xFFFFFFFF, 0x474D5421))
        side32 = np.float32(side)
        x = _u01(r[:, 0]) * side32
        y = _u01(r[:, 1]) * side32
        th = _u01(r[:, 2]) * np.float32(6.2831855) - np.float32(3.1415927)
        return np.stack([x, y, th], axis=1).astype(np.float32)
    if mode == "lattice":
        W = int(math.floor(float(np.float32(side)) / 4.0)) + 1
        i = np.arange(n)
        loc, k7 = i // 7, i % 7
        k = np.where(k7 >= 4, k7 + 1, k7)                  # skip k == 4 (cuda_agent.py:96-97)
        theta = k * 360.0 / 8                              # base yaw 0
        theta = np.where(theta >= 180, theta - 360, theta)  # cuda_agent.py:103-106
        x = (loc % W) * 4.0
        y = (loc // W) * 4.0
        return np.stack([x, y, theta * np.pi / 180], axis=1).astype(np.float32)
    raise ValueError(mode)


def build_neighbors(states, disk_radius=4 * math.sqrt(2)):
    """CSR neighbour lists (neighbors int32[sum deg], num_neighbors int32[n]) by the rule above.

    Cell-binned NumPy implementation (exact float32 arithmetic, no FMA): candidates come from
    the 3x3 block of cells of size >= disk_radius around each state.
    """
    states = np.ascontiguousarray(states, np.float32)
    n = states.shape[0]
    x, y = states[:, 0], states[:, 1]
    r = float(disk_radius)
    h = np.float32(r * 1.001 + 1e-3)
    x0, y0 = x.min() if n else np.float32(0), y.min() if n else np.float32(0)
    cx = np.floor((x - x0) / h).astype(np.int64)
    cy = np.floor((y - y0) / h).astype(np.int64)
    nx = int(cx.max()) + 1 if n else 1
    cell = cy * nx + cx
    order = np.argsort(cell, kind="stable")
    sorted_cell = cell[order]
    ncell = (int(cy.max()) + 1 if n else 1) * nx
    starts = np.searchsorted(sorted_cell, np.arange(ncell + 1))
    rows = [None] * n
    for c in np.unique(cell):
        members = order[starts[c]:starts[c + 1]]
        ccx, ccy = c % nx, c // nx
        cand = []
        for dy in (-1, 0, 1):
            yy = ccy + dy
            if yy < 0 or yy * nx >= ncell:
                continue
            for dx in (-1, 0, 1):
                xx = ccx + dx
                if xx < 0 or xx >= nx:
                    continue
                cc = yy * nx + xx
                cand.append(order[starts[cc]:starts[cc + 1]])
        cand = np.sort(np.concatenate(cand))
        dxm = x[members][:, None] - x[cand][None, :]
        dym = y[members][:, None] - y[cand][None, :]
        d = np.sqrt((dxm * dxm + dym * dym).astype(np.float32)).astype(np.float32)
        same = (x[members][:, None] == x[cand][None, :]) & (y[members][:, None] == y[cand][None, :])
        ok = (d.astype(np.float64) <= r) & ~same
        for a, i in enumerate(members):
            rows[i] = cand[ok[a]]
    num = np.array([len(rw) for rw in rows], np.int32) if n else np.zeros(0, np.int32)
    nbr = np.concatenate(rows).astype(np.int32) if n and num.sum() else np.zeros(0, np.int32)
    return nbr, num


def build_neighbors_bruteforce(states, disk_radius=4 * math.sqrt(2)):
    """O(n^2) version of the same rule, for cross-checking the binned one on small inputs."""
    states = np.ascontiguousarray(states, np.float32)
    x, y = states[:, 0], states[:, 1]
    dx = x[:, None] - x[None, :]
    dy = y[:, None] - y[None, :]
    d = np.sqrt((dx * dx + dy * dy).astype(np.float32)).astype(np.float32)
    same = (x[:, None] == x[None, :]) & (y[:, None] == y[None, :])
    ok = (d.astype(np.float64) <= float(disk_radius)) & ~same
    num = ok.sum(1).astype(np.int32)
    nbr = np.nonzero(ok)[1].astype(np.int32)
    return nbr, num
