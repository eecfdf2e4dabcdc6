"""Generates the golden vectors in this directory from the reference's OWN kernel source.

Run in the build container (needs /root/reference, which oracle/build_oracle.py compiles for the
host into oracle/_ref/libref_host.so):   python tests/golden/make_golden.py

* unit_tests.json   -- the reference's three planner fixtures (carla/pycuda_example.py:854-956)
                       run through the reference kernels under the restated GMT.run_step loop
                       (oracle/gmt_oracle.py:OracleGMT): exit status, iteration count, route,
                       parent[], cost[] (as float32 bit patterns) and the per-iteration G/Y/X sets.
                       They reproduce SURVEY.md appendix B.
* edge_vectors.npz  -- 4096 random (y -> x) pairs on a small obstacle map through the reference's
                       RSR/LSL/LSR/RSLcost + check_col + computeDubinsCost: per-word lengths,
                       exists / collides bits, min over free words.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import build_oracle  # noqa: E402
from oracle.gmt_oracle import OracleGMT, RefHost  # noqa: E402
from tests.golden.fixtures import FIXTURES  # noqa: E402


def edge_world(seed=7, n=600, m=14, side=40.0):
    rng = np.random.default_rng(seed)
    states = np.stack([rng.uniform(0, side, n), rng.uniform(0, side, n), rng.uniform(-np.pi, np.pi, n)],
                      1).astype(np.float32)
    c = rng.uniform(0, side, (m, 2))
    h = rng.uniform(0.5, 2.5, (m, 2))
    obs = np.concatenate([c + h, c - h], 1).astype(np.float32)      # corners in "reversed" order
    y = rng.integers(0, n, 4096).astype(np.int32)
    x = rng.integers(0, n, 4096).astype(np.int32)
    x[x == y] = (x[x == y] + 1) % n
    # half of the pairs close together (the planner's regime), half anywhere
    near = np.argsort(np.hypot(states[:, None, 0] - states[None, :, 0],
                               states[:, None, 1] - states[None, :, 1]), axis=1)[:, 1:24]
    x[:2048] = near[y[:2048], rng.integers(0, near.shape[1], 2048)]
    return states, obs, y, x


def main():
    build_oracle.build_all()
    ref = RefHost()
    out = {}
    for name, fx in FIXTURES.items():
        init, it = fx()
        g = OracleGMT(init, ref)
        route = g.run_step(it, iter_limit=20)
        out[name] = {
            "status": int(g.status), "iterations": int(g.iterations), "route": [int(r) for r in route],
            "parent": [int(p) for p in g.parent],
            "cost_bits": [int(b) for b in g.dev_cost.view(np.uint32)],
            "trace": [{k: (v.tolist() if hasattr(v, "tolist") else int(v)) for k, v in e.items()}
                      for e in g.trace],
        }
    with open(os.path.join(HERE, "unit_tests.json"), "w") as f:
        json.dump(out, f, indent=1)
    states, obs, y, x = edge_world()
    for radius, tag in ((2.0, "r2"), (1.0, "r1")):
        cost4, bits, best = ref.edge_costs(states, y, x, radius, obs, [obs.shape[0]])
        out_e = dict(states=states, obstacles=obs, y=y, x=x, radius=np.float32(radius),
                     cost4_bits=cost4.view(np.uint32), verdict_bits=bits, best_bits=best.view(np.uint32))
        np.savez_compressed(os.path.join(HERE, "edge_vectors_%s.npz" % tag), **out_e)
    print("golden vectors written")


if __name__ == "__main__":
    main()
