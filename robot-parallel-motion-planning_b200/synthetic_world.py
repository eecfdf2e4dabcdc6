"""synthetic_world.py -- the 2-D world that stands in for CARLA (host side, NumPy).

Produces exactly the two dicts the planner consumes (reference: carla/cuda_agent.py:115,215):
``init_parameters`` {'states','neighbors','num_neighbors','threadsPerBlock'} and
``iter_parameters`` {'start','goal','radius','threshold','obstacles','num_obs'}.

World definition: square of side
S = sqrt(n / 0.55) m, neighbour radius 4*sqrt(2) m (cuda_agent.py:56), AABB obstacles with
half-extents U[1,5] m given as two opposite corners (cuda_agent.py:196-206), start / goal = the
states nearest (0.1 S, 0.1 S) / (0.9 S, 0.9 S), radius 2.0, threshold 1.0
(test_agent.py:199-200).  States come from ``sampler(seed, n, side, mode)`` and the CSR lists
from ``neighbor_builder(states, radius)``: by default the CUDA builders of the library
(gmt_sample_states / gmt_build_neighbors); CPU-only tests pass the NumPy restatements of
oracle/world_oracle.py instead.
"""
import ctypes as C
import math

import numpy as np

try:
    from . import _lib
except ImportError:
    import _lib

DENSITY = 0.55
DISK_RADIUS = 4 * math.sqrt(2)

# BASELINE.json configs 1..5
CONFIGS = {
    1: dict(n=2000, m=20, mode="uniform", layout="uniform"),
    2: dict(n=10000, m=100, mode="uniform", layout="uniform"),
    3: dict(n=100000, m=2000, mode="uniform", layout="roads"),
    4: dict(n=10000, m=100, mode="uniform", layout="uniform", queries=4096),
    5: dict(n=1000000, m=2000, mode="uniform", layout="uniform"),
}


def cuda_sampler(seed, n, side, mode="uniform", device=0):
    """gmt_sample_states: Philox4x32-10 sampler on the device."""
    lib = _lib.load()
    out = np.zeros((int(n), 3), np.float32)
    rc = lib.gmt_sample_states(int(seed) & 0xFFFFFFFF, int(n), float(np.float32(side)),
                               0 if mode == "uniform" else 1, int(device), out.ctypes.data)
    _lib.check(rc, "gmt_sample_states")
    return out


def cuda_neighbor_builder(states, disk_radius=DISK_RADIUS, device=0):
    """gmt_build_neighbors: cell-grid radius search on the device -> (neighbors, num_neighbors)."""
    lib = _lib.load()
    states = np.ascontiguousarray(states, np.float32)
    n = states.shape[0]
    vals, counts, total = C.c_void_p(), C.c_void_p(), C.c_longlong()
    rc = lib.gmt_build_neighbors(states.ctypes.data, n, float(disk_radius), int(device),
                                 C.byref(vals), C.byref(counts), C.byref(total))
    _lib.check(rc, "gmt_build_neighbors")
    try:
        num = np.ctypeslib.as_array(C.cast(counts, C.POINTER(C.c_int)), shape=(n,)).copy()
        nbr = (np.ctypeslib.as_array(C.cast(vals, C.POINTER(C.c_int)), shape=(max(total.value, 1),))
               [:total.value].copy())
    finally:
        lib.gmt_free(vals)
        lib.gmt_free(counts)
    return nbr.astype(np.int32), num.astype(np.int32)


def world_side(n):
    return math.sqrt(n / DENSITY)


def make_obstacles(rng, m, side, layout):
    """m AABBs [xa, ya, xb, yb] with half-extents U[1,5] m, stored as two opposite corners in
    arbitrary order like the reference does (cuda_agent.py:197)."""
    boxes = []
    while len(boxes) < m:
        c = rng.uniform(0.0, side, 2)
        if layout == "roads":
            # streets every 40 m, 12 m wide: keep centres on a street
            on_street = (abs((c[0] + 20.0) % 40.0 - 20.0) < 6.0) or (abs((c[1] + 20.0) % 40.0 - 20.0) < 6.0)
            if not on_street:
                continue
        h = rng.uniform(1.0, 5.0, 2)
        lo, hi = c - h, c + h
        if rng.random() < 0.5:
            boxes.append([hi[0], hi[1], lo[0], lo[1]])
        else:
            boxes.append([lo[0], lo[1], hi[0], hi[1]])
    if m == 0:
        return np.array([[-1, -1, -1, -1]]).astype(np.float32), np.array([0]).astype(np.int32)
    return np.array(boxes).astype(np.float32), np.array([m]).astype(np.int32)


def clearance(points, obstacles, m):
    """Distance from each (x, y) to the nearest of the first m boxes (0 inside a box)."""
    pts = np.asarray(points, np.float64).reshape(-1, 2)
    if m == 0:
        return np.full(len(pts), np.inf)
    obs = np.asarray(obstacles, np.float64)[:m]
    lo = np.minimum(obs[:, :2], obs[:, 2:])
    hi = np.maximum(obs[:, :2], obs[:, 2:])
    d = np.maximum(np.maximum(lo[None] - pts[:, None], pts[:, None] - hi[None]), 0.0)
    return np.hypot(d[..., 0], d[..., 1]).min(axis=1)


def nearest_clear_state(states, x, y, obstacles, m, clear=4.0, tries=2000):
    """The state nearest (x, y) that keeps `clear` metres from every box (the reference starts from a
    driving car, never from inside a parked one); falls back to the plain nearest state."""
    d = (states[:, 0].astype(np.float64) - x) ** 2 + (states[:, 1].astype(np.float64) - y) ** 2
    order = np.argsort(d, kind="stable")[:tries]
    ok = clearance(states[order, :2], obstacles, m) >= clear
    return int(order[np.argmax(ok)]) if ok.any() else int(order[0])


def nearest_state(states, x, y):
    d = (states[:, 0].astype(np.float64) - x) ** 2 + (states[:, 1].astype(np.float64) - y) ** 2
    return int(np.argmin(d))


def make_world(config=None, n=None, m=None, mode=None, layout=None, seed=None, sampler=None,
               neighbor_builder=None, radius=2.0, threshold=1.0, side=None, with_neighbors=True):
    """Returns (init_parameters, iter_parameters, meta) for a BASELINE config or explicit sizes.
    with_neighbors=False leaves the lists out (init_parameters['neighbors'] = None) for GMT.from_states, which
    builds them on the device."""
    cfg = dict(CONFIGS[config]) if config is not None else {}
    n = int(n if n is not None else cfg["n"])
    m = int(m if m is not None else cfg["m"])
    mode = mode or cfg.get("mode", "uniform")
    layout = layout or cfg.get("layout", "uniform")
    seed = int(seed if seed is not None else 18 + (config or 0))       # seed base 18, run.py:45
    sampler = sampler or cuda_sampler
    neighbor_builder = neighbor_builder or cuda_neighbor_builder
    side = float(side if side is not None else world_side(n))

    states = np.ascontiguousarray(sampler(seed, n, side, mode), np.float32)
    n = states.shape[0]
    neighbors, num_neighbors = neighbor_builder(states, DISK_RADIUS) if with_neighbors else (None, None)
    rng = np.random.default_rng(seed)
    obstacles, num_obs = make_obstacles(rng, m, side, layout)
    start = nearest_clear_state(states, 0.1 * side, 0.1 * side, obstacles, m)
    goal = nearest_clear_state(states, 0.9 * side, 0.9 * side, obstacles, m)

    init_parameters = {'states': states, 'neighbors': neighbors, 'num_neighbors': num_neighbors,
                       'threadsPerBlock': 128}
    iter_parameters = {'start': start, 'goal': goal, 'radius': radius, 'threshold': threshold,
                       'obstacles': obstacles, 'num_obs': num_obs}
    meta = dict(n=n, m=m, side=side, seed=seed, mode=mode, layout=layout,
                mean_degree=float(num_neighbors.mean()) if (n and num_neighbors is not None) else None)
    return init_parameters, iter_parameters, meta


def explore_cuda(init_parameters, iter_parameters, iter_limit=1000, device=0):
    """Runs the planner without a goal and returns the ids that ever entered a wavefront G."""
    try:
        from .gmt_planner import GMT
    except ImportError:
        from gmt_planner import GMT
    g = GMT(init_parameters, device=device)
    g.run_step(dict(iter_parameters, goal=-1), iter_limit=iter_limit, debug=False, time=True)
    members = [e["G"] for e in g.trace if e["gSize"] > 0]
    return np.unique(np.concatenate(members)) if members else np.zeros(0, np.int32)


def pick_reachable_goal(init_parameters, iter_parameters, side, explore=None, frac=0.9):
    """The reference's GMT variant closes every open node above the cost threshold unexpanded
    (kernel.py:445), so on a random roadmap many states can never enter a
    wavefront and a plan to them spins to iter_limit.  Benchmark queries therefore use as goal the
    state nearest (frac*S, frac*S) among those that DO enter a wavefront from this start."""
    reach = (explore or explore_cuda)(init_parameters, iter_parameters)
    states = init_parameters['states']
    if len(reach) == 0:
        return int(iter_parameters['goal'])
    d = (states[reach, 0].astype(np.float64) - frac * side) ** 2 + (states[reach, 1].astype(np.float64) - frac * side) ** 2
    return int(reach[int(np.argmin(d))])


def make_queries(states, side, q, seed):
    """q (start, goal) pairs drawn uniformly with |start - goal| >= side/2 (BASELINE config 4)."""
    rng = np.random.default_rng(seed)
    n = states.shape[0]
    starts, goals = [], []
    while len(starts) < q:
        a, b = int(rng.integers(0, n)), int(rng.integers(0, n))
        if np.hypot(*(states[a, :2] - states[b, :2])) >= side / 2:
            starts.append(a)
            goals.append(b)
    return np.array(starts, np.int32), np.array(goals, np.int32)


def plan_batch(handle_owner, starts, goals, radius=2.0, threshold=1.0, iter_limit=1000, route_stride=0):
    """gmt_plan_batch on the GMT instance's handle (obstacles must be resident: gmt_set_obstacles).
    Returns (results ctypes array, routes int32[q, route_stride] or None)."""
    lib = _lib.load()
    starts = np.ascontiguousarray(starts, np.int32)
    goals = np.ascontiguousarray(goals, np.int32)
    q = len(starts)
    res = (_lib.GmtResult * q)()
    routes = np.zeros((q, route_stride), np.int32) if route_stride > 0 else None
    rc = lib.gmt_plan_batch(handle_owner._handle, starts.ctypes.data, goals.ctypes.data, q, float(np.float32(radius)),
                            float(np.float32(threshold)), int(iter_limit), res,
                            routes.ctypes.data if routes is not None else None, int(route_stride))
    _lib.check(rc, "gmt_plan_batch")
    return res, routes


def make_reachable_queries(gmt, n, q, seed, radius=2.0, threshold=1.0, iter_limit=1000):
    """q independent queries for BASELINE config 4: starts drawn uniformly; each goal is the state that
    enters a wavefront farthest (straight line) from its start -- found with one goal-less batch
    (gmt_result.farthest_frontier), so that every query ends with "goal reached"."""
    rng = np.random.default_rng(seed)
    starts = rng.integers(0, n, q).astype(np.int32)
    res, _ = plan_batch(gmt, starts, np.full(q, -1, np.int32), radius, threshold, iter_limit)
    goals = np.array([r.farthest_frontier if r.farthest_frontier >= 0 else int(s) for r, s in zip(res, starts)],
                     np.int32)
    return starts, goals


# ------------------------------------------------------------------------------------------------------
# Obstacle ingestion (reference: CudaAgent._trace_route, cuda_agent.py:127-206)
# ------------------------------------------------------------------------------------------------------
def vehicles_to_obstacles(vehicles, ego_xy, margin=2.5, max_dist=30.0, oriented=False):
    """Vehicle boxes -> the planner's obstacle array, the way the reference builds it every re-plan.

    vehicles: rows [x, y, yaw_deg, extent_x, extent_y, box_offset_x, box_offset_y] (vehicle pose, half
    extents and centre offset of its bounding box in the vehicle frame: what CARLA's actor transform and
    bounding_box give; planar world, so pitch = roll = 0 and z plays no role).  ego_xy: centre of the ego
    vehicle's box.  Kept like the reference (cuda_agent.py:186-198):
      * the two diagonal corners (+-(extent + 2.5 m)) of the inflated box are taken to the world frame and
        used AS the two opposite corners [xa, ya, xb, yb] of an axis-aligned box -- for a turned vehicle
        that is not the bounding box of the rotated rectangle (at 45 degrees it degenerates to a sliver);
      * only vehicles whose box centre is within 0 < d <= 30 m of the ego's are obstacles (which also drops
        the ego itself);
      * no obstacle -> the placeholder [[-1, -1, -1, -1]] with num_obs = [0].
    oriented=True (extension, not in the reference) returns instead the geometrically faithful inflated
    rectangles as rows [centre_x, centre_y, half_x, half_y, angle_rad] for
    iter_parameters['obstacle_format'] = 'oriented'.
    Returns (obstacles float32[m, 4 or 5], num_obs int32[1])."""
    v = np.asarray(vehicles, dtype=np.float64).reshape(-1, 7)
    ego = np.asarray(ego_xy, dtype=np.float64).reshape(2)
    yaw = np.radians(v[:, 2])
    c, s = np.cos(yaw), np.sin(yaw)

    def to_world(px, py):          # bounding-box frame -> vehicle frame (offset) -> world (yaw, position)
        lx, ly = v[:, 5] + px, v[:, 6] + py
        return v[:, 0] + c * lx - s * ly, v[:, 1] + s * lx + c * ly

    ex, ey = v[:, 3] + margin, v[:, 4] + margin
    xa, ya = to_world(ex, ey)
    xb, yb = to_world(-ex, -ey)
    cx, cy = to_world(np.zeros(len(v)), np.zeros(len(v)))
    d = np.hypot(cx - ego[0], cy - ego[1])
    near = (d > 0) & (d <= max_dist)
    m = int(near.sum())
    if m == 0:
        width = 5 if oriented else 4
        return np.full((1, width), -1, np.float32), np.array([0], np.int32)
    if oriented:
        obs = np.stack([cx, cy, ex, ey, yaw], 1)[near]
    else:
        obs = np.stack([xa, ya, xb, yb], 1)[near]
    return obs.astype(np.float32), np.array([m], np.int32)
