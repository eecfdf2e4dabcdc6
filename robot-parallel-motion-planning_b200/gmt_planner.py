"""gmt_planner.py -- drop-in for the reference module of the same name.

The reference's callers do ``from gmt_planner import *`` and use the name ``GMT`` only
(/root/reference/carla/cuda_agent.py:11,122,218; carla/test_agent.py:30,141,232).  This module
keeps that surface -- ``GMT(init_parameters, debug=False)`` and
``run_step(iter_parameters, iter_limit=1000, debug=False, time=False) -> route`` with the same
dict keys, the same goal-first route list, the same stale-route behaviour on failure and the
same public attributes (``route, start, goal, n, parent, time_data`` ...) as
/root/reference/carla/gmt_planner.py:22-239 -- while the planning itself is one persistent
CUDA kernel for sm_100a behind the C ABI of include/gmt_c_api.h (ctypes, no PyCuda, no CPU
fallback: importing this module without the built library raises).
"""
import ctypes as C
import math

import numpy as np

try:
    from . import _lib
except ImportError:  # imported as a top-level module, the way the reference's callers do
    import _lib

__all__ = ["GMT"]


def oriented_records(rects):
    """Extension mode: rows [centre_x, centre_y, half_x, half_y, angle_rad] -> the float32 records
    [cx, cy, half_x, half_y, cos_a, sin_a] that gmt_set_obstacles_oriented takes (cos / sin from libm
    in double, rounded once)."""
    r = np.asarray(rects, dtype=np.float64).reshape(-1, 5)
    out = np.zeros((r.shape[0], 6), np.float32)
    out[:, :4] = r[:, :4]
    out[:, 4] = [math.cos(a) for a in r[:, 4]]
    out[:, 5] = [math.sin(a) for a in r[:, 4]]
    return out


class GMT(object):
    """GMT* planner over a fixed Dubins roadmap (reference: carla/gmt_planner.py:22-239)."""

    # the reference prints its exit reason (gmt_planner.py:144,151); set True to get the same lines
    verbose = False

    def __init__(self, init_parameters, debug=False, device=0, words=4):
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        self._cpu_init(init_parameters, debug)
        self._gpu_init(debug, device)
        self._common_init(words)

    @classmethod
    def from_states(cls, states, disk_radius=4 * math.sqrt(2), debug=False, device=0, words=4):
        """Extension: a planner for the roadmap over `states` (float32[n, 3]) whose neighbour lists
        -- the rule of CudaAgent.create_samples, cuda_agent.py:71-91, as gmt_build_neighbors states it -- are built
        ON THE DEVICE and stay there (gmt_create_from_states): the lists never visit the host.  Same plans as
        GMT({'states': states, 'neighbors': ..., 'num_neighbors': ...}) with the lists gmt_build_neighbors returns;
        .neighbors / .num_neighbors download them on first use."""
        self = cls.__new__(cls)
        self._lib = _lib.load()
        self._handle = C.c_void_p()
        self.states = np.ascontiguousarray(states, dtype=np.float32).reshape(-1, 3)
        self.n = self.states.shape[0]
        self.waypoints = np.arange(self.n).astype(np.int32)
        self._neighbors = self._num_neighbors = None
        self.cost = np.full(self.n, np.inf).astype(np.float32)
        self.Vunexplored = np.full(self.n, 1).astype(np.int32)
        self.Vopen = np.zeros_like(self.Vunexplored).astype(np.int32)
        self.threadsPerBlock = 128
        rc = self._lib.gmt_create_from_states(self.states.ctypes.data, self.n, float(disk_radius), int(device),
                                              C.byref(self._handle))
        _lib.check(rc, "gmt_create_from_states")
        self._n_base = self.n
        self._common_init(words)
        return self

    def _common_init(self, words):
        self.words = 4
        if words != 4:
            self.set_words(words)

        self.route = []
        self.start = 0
        self.time_data = {"wavefront": [], "wavefront_compact": [], "open_compact": [], "neighbors": [],
                          "neighbors_compact": [], "connection": [], "elapsed": [], "threshold": [],
                          "goal": [], "iteration": []}
        self.last_result = None

    # ---- gmt_planner.py:31-44 ---------------------------------------------------------------
    def _cpu_init(self, init_parameters, debug):
        self.states = init_parameters['states']
        self.n = self.states.shape[0]
        self.waypoints = np.arange(self.n).astype(np.int32)
        self._neighbors = init_parameters['neighbors']
        self._num_neighbors = init_parameters['num_neighbors']
        self._n_base = self.n
        # kept for callers that poke at them; the device owns the live copies
        self.cost = np.full(self.n, np.inf).astype(np.float32)
        self.Vunexplored = np.full(self.n, 1).astype(np.int32)
        self.Vopen = np.zeros_like(self.Vunexplored).astype(np.int32)
        self.threadsPerBlock = init_parameters.get('threadsPerBlock', 128)  # accepted, not needed
        if debug:
            print('neighbors: ', self.neighbors)
            print('number neighbors: ', self.num_neighbors)

    # ---- gmt_planner.py:50-61: roadmap to the device, once ---------------------------------------
    def _gpu_init(self, debug, device=0):
        states = np.ascontiguousarray(self.states, dtype=np.float32).reshape(self.n, 3)
        nbr = np.ascontiguousarray(self.neighbors, dtype=np.int32).reshape(-1)
        cnt = np.ascontiguousarray(self.num_neighbors, dtype=np.int32).reshape(-1)
        if cnt.shape[0] != self.n:
            raise ValueError("num_neighbors must have one entry per state")
        if int(cnt.sum()) != nbr.shape[0]:
            raise ValueError("neighbors does not hold sum(num_neighbors) entries")
        rc = self._lib.gmt_create(states.ctypes.data, self.n, nbr.ctypes.data if nbr.size else None,
                                  cnt.ctypes.data, int(device), C.byref(self._handle))
        _lib.check(rc, "gmt_create")

    def _download_neighbors(self):
        total = C.c_longlong()
        _lib.check(self._lib.gmt_get_neighbors(self._handle, None, None, 0, C.byref(total)), "gmt_get_neighbors")
        counts = np.zeros(self._n_base, np.int32)
        vals = np.zeros(max(total.value, 1), np.int32)
        _lib.check(self._lib.gmt_get_neighbors(self._handle, counts.ctypes.data, vals.ctypes.data, int(vals.shape[0]),
                                               C.byref(total)), "gmt_get_neighbors")
        self._num_neighbors, self._neighbors = counts, vals[:total.value]

    @property
    def neighbors(self):
        if self._neighbors is None:
            self._download_neighbors()
        return self._neighbors

    @neighbors.setter
    def neighbors(self, value):
        self._neighbors = value

    @property
    def num_neighbors(self):
        if self._num_neighbors is None:
            self._download_neighbors()
        return self._num_neighbors

    @num_neighbors.setter
    def num_neighbors(self, value):
        self._num_neighbors = value

    def set_words(self, words):
        """Extension mode (not in the reference, which tries RSR, LSL, LSR, RSL: kernel.py:421-424):
        words=6 also tries the CCC words RLR and LRL in every later plan; words=4 restores parity mode."""
        _lib.check(self._lib.gmt_set_words(self._handle, int(words)), "gmt_set_words")
        self.words = int(words)

    def set_option(self, name, value):
        """Tuning / test knob of the library (gmt_set_option): never changes a result."""
        _lib.check(self._lib.gmt_set_option(self._handle, str(name).encode(), int(value)), "gmt_set_option")

    def set_obstacles(self, obstacles, num_obs=None):
        """Extension: make an obstacle set resident on the device; later run_step calls whose
        iter_parameters['obstacles'] is None plan against it without any upload."""
        obs = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
        m = obs.shape[0] if num_obs is None else int(np.asarray(num_obs).reshape(-1)[0])
        _lib.check(self._lib.gmt_set_obstacles(self._handle, obs.ctypes.data if m > 0 else None, m), "gmt_set_obstacles")

    def set_obstacles_from_vehicles(self, vehicles, ego_xy, margin=2.5, max_dist=30.0, oriented=False):
        """Extension: the reference's per-re-plan obstacle pipeline (cuda_agent.py:127-206: vehicle
        boxes -> inflated boxes within 30 m of the ego) run ON THE DEVICE; the result becomes the resident obstacle
        set, so ``run_step(dict(iter_parameters, obstacles=None))`` plans against it.  vehicles: rows [x, y, yaw_deg,
        extent_x, extent_y, box_offset_x, box_offset_y] as synthetic_world.vehicles_to_obstacles takes them (the
        NumPy statement of the same pipeline, which this reproduces bit for bit).  Returns the number of obstacles."""
        v = np.asarray(vehicles, dtype=np.float64).reshape(-1, 7)
        ego = np.asarray(ego_xy, dtype=np.float64).reshape(2)
        yaw = np.radians(v[:, 2])
        yaw32 = yaw.astype(np.float32)
        rows = np.stack([v[:, 0], v[:, 1], np.cos(yaw), np.sin(yaw), v[:, 3], v[:, 4], v[:, 5], v[:, 6],
                         np.array([math.cos(a) for a in yaw32], np.float64),
                         np.array([math.sin(a) for a in yaw32], np.float64)], 1) if len(v) else np.zeros((0, 10))
        rows = np.ascontiguousarray(rows, np.float64)
        m = C.c_int()
        rc = self._lib.gmt_set_obstacles_from_vehicles(self._handle, rows.ctypes.data if len(v) else None, len(v),
                                                       float(ego[0]), float(ego[1]), float(margin), float(max_dist),
                                                       1 if oriented else 0, C.byref(m))
        _lib.check(rc, "gmt_set_obstacles_from_vehicles")
        return m.value

    def get_obstacles(self):
        """The resident obstacle set: (corner boxes float32[m, 4], oriented records float32[m, 6] or None)."""
        m = C.c_int()
        _lib.check(self._lib.gmt_get_obstacles(self._handle, None, None, 0, C.byref(m)), "gmt_get_obstacles")
        xyxy = np.zeros((max(m.value, 1), 4), np.float32)
        obb = np.zeros((max(m.value, 1), 6), np.float32)
        _lib.check(self._lib.gmt_get_obstacles(self._handle, xyxy.ctypes.data, obb.ctypes.data, m.value, C.byref(m)),
                   "gmt_get_obstacles")
        return xyxy[:m.value], (obb[:m.value] if obb[:m.value].any() else None)

    def update_obstacles(self, index, obstacles):
        """Replace the boxes `index` of the resident set in place (only those cross PCIe)."""
        idx = np.ascontiguousarray(index, dtype=np.int32).reshape(-1)
        obs = np.ascontiguousarray(obstacles, dtype=np.float32).reshape(-1, 4)
        if obs.shape[0] != idx.shape[0]:
            raise ValueError("one box per index")
        _lib.check(self._lib.gmt_update_obstacles(self._handle, idx.ctypes.data, obs.ctypes.data, idx.shape[0]),
                   "gmt_update_obstacles")

    def insert_states(self, states_xyt, disk_radius=4 * math.sqrt(2)):
        """Extension: append states (e.g. the vehicle's current pose as a new start) to the resident
        roadmap without rebuilding the neighbour lists; returns their ids.  The reference rebuilds the whole
        roadmap with start and goal appended (cuda_agent.py:64-66).  Plans equal plans on the rebuilt roadmap."""
        new = np.ascontiguousarray(states_xyt, dtype=np.float32).reshape(-1, 3)
        first = C.c_int()
        _lib.check(self._lib.gmt_insert_states(self._handle, new.ctypes.data, new.shape[0], float(disk_radius),
                                               C.byref(first)), "gmt_insert_states")
        self.states = np.concatenate([np.asarray(self.states, np.float32).reshape(-1, 3), new], 0)
        self.n = self.states.shape[0]
        self.waypoints = np.arange(self.n).astype(np.int32)
        self._parent = self._cost = None
        return list(range(first.value, first.value + new.shape[0]))

    def reset_inserted(self):
        """Drops every state appended by insert_states."""
        _lib.check(self._lib.gmt_reset_inserted(self._handle), "gmt_reset_inserted")
        self.n = int(self._n_base)
        self.states = np.asarray(self.states).reshape(-1, 3)[:self.n]
        self.waypoints = np.arange(self.n).astype(np.int32)
        self._parent = self._cost = None

    def __del__(self):
        h = getattr(self, "_handle", None)
        if h is not None and h.value:
            try:
                self._lib.gmt_destroy(h)
            except Exception:
                pass
            self._handle = C.c_void_p()

    # ---- gmt_planner.py:63-101 -------------------------------------------------------------------
    def step_init(self, iter_parameters, debug):
        if self.start != iter_parameters['start'] and len(self.route) > 2:
            del self.route[-1]                                 # gmt_planner.py:68-69

        self.obstacles = iter_parameters.get('obstacles')
        self.num_obs = iter_parameters.get('num_obs')
        self.start = iter_parameters['start']
        self.goal = iter_parameters['goal']
        self.radius = iter_parameters['radius']
        self.threshold = np.array([iter_parameters['threshold']]).astype(np.float32)
        self._parent = None
        self._cost = None

    def get_path(self):
        """gmt_planner.py:103-107 -- the walk goal -> start happened on the device."""
        res = getattr(self, "last_result", None)
        cap = (res.route_len if res is not None and res.route_len >= 0 else self.n) + 1   # no n-sized buffer per plan
        buf = np.empty(cap, np.int32)
        ln = C.c_int()
        _lib.check(self._lib.gmt_get_route(self._handle, buf.ctypes.data, cap, C.byref(ln)), "gmt_get_route")
        self.route += buf[:ln.value].tolist()

    def route_edges(self):
        """Trajectory of the last route, start -> goal: arrays (from_to int32[E,2], word uint8[E],
        arcs float32[E,3] = first arc, straight segment, second arc lengths).  What a controller needs
        to track the Dubins curve instead of chasing waypoints (reference README.md:257)."""
        cap = max(len(self.route) - 1, 0)
        ft = np.zeros((cap, 2), np.int32)
        word = np.zeros(cap, np.uint8)
        arcs = np.zeros((cap, 3), np.float32)
        ln = C.c_int()
        _lib.check(self._lib.gmt_get_route_edges(self._handle, ft.ctypes.data, word.ctypes.data, arcs.ctypes.data,
                                                 cap, C.byref(ln)), "gmt_get_route_edges")
        return ft[:ln.value], word[:ln.value], arcs[:ln.value]

    # ---- results the reference downloads (gmt_planner.py:145,152) plus the cost array -----------
    @property
    def parent(self):
        if getattr(self, "_parent", None) is None:
            out = np.zeros(self.n, np.int32)
            _lib.check(self._lib.gmt_get_parent(self._handle, out.ctypes.data), "gmt_get_parent")
            self._parent = out
        return self._parent

    @parent.setter
    def parent(self, value):
        self._parent = value

    @property
    def dev_cost(self):
        """cost[] after the last plan (the reference never downloads it; parity tests need it)."""
        if getattr(self, "_cost", None) is None:
            out = np.zeros(self.n, np.float32)
            _lib.check(self._lib.gmt_get_cost(self._handle, out.ctypes.data), "gmt_get_cost")
            self._cost = out
        return self._cost

    # ---- gmt_planner.py:109-239 ------------------------------------------------------------------
    def run_step(self, iter_parameters, iter_limit=1000, debug=False, time=False):
        self.step_init(iter_parameters, debug)
        if self.obstacles is None:                       # extension: plan against the resident obstacle set
            want_trace = bool(debug or time)
            _lib.check(self._lib.gmt_set_trace(self._handle, 1 if want_trace else 0), "gmt_set_trace")
            res = _lib.GmtResult()
            rc = self._lib.gmt_plan_resident(self._handle, int(self.start), int(self.goal), float(np.float32(self.radius)),
                                             float(self.threshold[0]), int(iter_limit), C.byref(res))
            _lib.check(rc, "gmt_plan_resident")
            return self._finish(res, want_trace, debug, time)
        num_obs = int(np.asarray(self.num_obs).reshape(-1)[0])
        # extension (not in the reference): iter_parameters['obstacle_format'] = 'oriented' -> rows
        # [centre_x, centre_y, half_x, half_y, angle_rad]; default: the reference's [xa, ya, xb, yb]
        oriented = iter_parameters.get('obstacle_format', 'corners') == 'oriented'
        obs = np.ascontiguousarray(self.obstacles, dtype=np.float32).reshape(-1, 5 if oriented else 4)
        if num_obs > obs.shape[0]:
            raise ValueError("num_obs exceeds the obstacle array")
        want_trace = bool(debug or time)
        _lib.check(self._lib.gmt_set_trace(self._handle, 1 if want_trace else 0), "gmt_set_trace")
        res = _lib.GmtResult()
        if oriented:
            rec = oriented_records(obs[:num_obs])
            _lib.check(self._lib.gmt_set_obstacles_oriented(self._handle, rec.ctypes.data if num_obs > 0 else None,
                                                            num_obs), "gmt_set_obstacles_oriented")
            rc = self._lib.gmt_plan_resident(self._handle, int(self.start), int(self.goal),
                                             float(np.float32(self.radius)), float(self.threshold[0]),
                                             int(iter_limit), C.byref(res))
        else:
            rc = self._lib.gmt_plan(self._handle, int(self.start), int(self.goal),
                                    float(np.float32(self.radius)), float(self.threshold[0]),
                                    obs.ctypes.data if num_obs > 0 else None, num_obs, int(iter_limit),
                                    C.byref(res))
        _lib.check(rc, "gmt_plan")
        return self._finish(res, want_trace, debug, time)

    def _finish(self, res, want_trace, debug, time):
        self.last_result = res
        self.iterations = res.iterations
        self.status = res.status

        if res.status == _lib.GMT_STATUS_LIMIT:
            if self.verbose:
                print('### iteration limit ###', res.iterations)
            if res.route_len >= 0:                             # parent[goal] != -1, gmt_planner.py:146
                self.route = []
                self.get_path()
        else:
            if self.verbose:
                print('### goal reached ### ', res.iterations)
            self.route = []
            self.get_path()

        if want_trace:
            self._collect_trace(res.iterations, debug, time)
        return self.route

    def _collect_trace(self, iterations, debug, time):
        sizes = np.zeros((max(iterations, 1), 3), np.int32)
        ms = np.zeros(max(iterations, 1), np.float32)
        ln = C.c_longlong()
        _lib.check(self._lib.gmt_get_trace(self._handle, None, 0, None, 0, C.byref(ln), None), "gmt_get_trace")
        cap = max(int(ln.value), 1)
        sets = np.zeros(cap, np.int32)
        _lib.check(self._lib.gmt_get_trace(self._handle, sizes.ctypes.data, iterations, sets.ctypes.data,
                                           cap, C.byref(ln), ms.ctypes.data), "gmt_get_trace")
        ph = np.zeros((max(iterations, 1), 4), np.float32)
        _lib.check(self._lib.gmt_get_phase_times(self._handle, ph.ctypes.data, iterations), "gmt_get_phase_times")
        self.trace = []
        pos = 0
        for k in range(iterations):
            g, y, x = (int(v) for v in sizes[k])
            ent = {"gSize": g, "ySize": y, "xSize": x}
            ent["G"] = sets[pos:pos + g].copy(); pos += g
            if y > 0:
                ent["Y"] = sets[pos:pos + y].copy(); pos += y
            if x > 0:
                ent["X"] = sets[pos:pos + x].copy(); pos += x
            self.trace.append(ent)
            if debug:
                print('y size: ', y, 'G size: ', g, 'x size: ', x)
                print(f'######### iteration: {k + 1} iteration time: {ms[k] * 1e-3}')
            # the reference only times turns that reach the end of the loop body (gmt_planner.py:229).
            # Its host timers bracket launches; here each key gets the DEVICE time of the phase that does
            # that work inside the persistent kernel (seconds, like timeit's):
            #   wavefront          <- frontier phase (wavefront + growThreshold + goal test + neighborIndicator)
            #   neighbors_compact  <- select phase (which open nodes can still be a parent of which x)
            #   connection         <- word lengths + swept paths (dubinConnection)
            # the scans/compactions of the reference have no counterpart (lists are built in place): 0
            if time and y > 0 and x > 0:
                a, s, e, w = (float(v) * 1e-3 for v in ph[k])
                self.time_data["wavefront"].append(a)
                self.time_data["neighbors_compact"].append(s)
                self.time_data["connection"].append(e + w)
                for key in ("wavefront_compact", "open_compact", "neighbors", "threshold", "goal"):
                    self.time_data[key].append(0.0)
                self.time_data["elapsed"].append(a + s + e + w)
                self.time_data["iteration"].append(k + 1)
