"""oracle/gmt_oracle.py -- TEST INFRASTRUCTURE ONLY (ctypes front-end of the CPU oracles).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module; the product never does.

* ``OraclePort(flavour)``  -- the C restatement oracle/gmt_oracle.c
  (flavour "libm" = what g++ makes of the reference source, "cuda" = what nvcc 12.9
  makes of it for sm_100a; see the header of gmt_oracle.c).
* ``RefHost()``            -- the reference's own kernel source compiled for the host
  (oracle/_ref/libref_host.so, present only where oracle/build_oracle.py found
  /root/reference).
* ``OracleGMT``            -- restatement of the host class ``GMT`` of
  /root/reference/carla/gmt_planner.py:22-239, line for line in behaviour, running its
  five kernels + three scans through either backend.  It also records the per-iteration
  G / Y / X index lists that the parity tests compare.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

_f32p = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")


def _obstacles(obstacles, num_obs, width=4):
    """Reference convention (cuda_agent.py:201-206): a [[-1,-1,-1,-1]] placeholder with num_obs=[0].
    width 6 = extension-mode oriented rectangles [cx, cy, half_x, half_y, cos_a, sin_a]."""
    m = int(np.asarray(num_obs).reshape(-1)[0]) if num_obs is not None else len(obstacles)
    obs = np.ascontiguousarray(np.asarray(obstacles, dtype=np.float32).reshape(-1, width))
    if obs.shape[0] == 0:
        obs = np.full((1, width), -1, dtype=np.float32)
    return obs, m


class OraclePort:
    """oracle/gmt_oracle.c through ctypes."""

    def __init__(self, flavour="libm"):
        path = os.path.join(HERE, "libgmt_oracle_%s.so" % flavour)
        if not os.path.exists(path):
            raise FileNotFoundError(path + " -- run `python oracle/build_oracle.py`")
        self.flavour = flavour
        L = self.lib = C.CDLL(path)
        L.gmto_flavour.restype = C.c_char_p
        L.gmto_word.restype = C.c_int
        L.gmto_word.argtypes = [C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.POINTER(C.c_float),
                                C.c_void_p, C.c_void_p]
        L.gmto_word_arcs.restype = C.c_int
        L.gmto_word_arcs.argtypes = [C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.POINTER(C.c_float), _f32p]
        L.gmto_compute_dubins_cost.restype = C.c_int
        L.gmto_compute_dubins_cost.argtypes = [C.POINTER(C.c_float), C.c_float, _f32p, _f32p, C.c_float,
                                               _f32p, C.c_int]
        L.gmto_dubin_connection.restype = None
        L.gmto_dubin_connection.argtypes = [_f32p, _i32p, _i32p, _i32p, _f32p, _i32p, _i32p, C.c_int,
                                            C.c_int, _f32p, C.c_int, C.c_float]
        L.gmto_wavefront.restype = None
        L.gmto_wavefront.argtypes = [_i32p, _i32p, _f32p, C.c_float, C.c_int]
        L.gmto_grow_threshold.restype = C.c_float
        L.gmto_grow_threshold.argtypes = [C.c_float, C.c_float]
        L.gmto_neighbor_indicator.restype = None
        L.gmto_neighbor_indicator.argtypes = [_i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int]
        L.gmto_plan.restype = C.c_int
        L.gmto_plan.argtypes = [_f32p, C.c_int, _i32p, _i32p, C.c_int, C.c_int, C.c_float, C.c_float,
                                _f32p, C.c_int, C.c_int, _f32p, _i32p, _i32p, _i32p,
                                C.POINTER(C.c_int), C.c_void_p, C.c_int, C.c_void_p, C.c_long,
                                C.POINTER(C.c_long)]
        L.gmto_edge_costs.restype = None
        L.gmto_edge_costs.argtypes = [_f32p, _i32p, _i32p, C.c_long, C.c_float, _f32p, C.c_int, _f32p,
                                      _u32p, _f32p]
        L.gmto_get_path.restype = C.c_int
        L.gmto_get_path.argtypes = [_i32p, C.c_int, C.c_int, _i32p, C.c_int]
        L.gmto_set_num_threads.restype = C.c_int
        L.gmto_set_num_threads.argtypes = [C.c_int]
        L.gmto_set_extensions.restype = None
        L.gmto_set_extensions.argtypes = [C.c_int, C.c_int]
        self.words, self.oriented = 4, False

    def set_num_threads(self, n):
        """OpenMP threads of the x loops (returns the team size in effect)."""
        return int(self.lib.gmto_set_num_threads(int(n)))

    def num_threads(self):
        return int(self.lib.gmto_set_num_threads(0))

    def set_extensions(self, words=4, oriented=False):
        """Extension mode (parity unpinned, see gmt_oracle.c): 6 words and / or oriented rectangles
        given as rows [cx, cy, half_x, half_y, cos_a, sin_a].  (4, False) = reference-parity mode."""
        self.words, self.oriented = (6 if int(words) == 6 else 4), bool(oriented)
        self.lib.gmto_set_extensions(self.words, int(self.oriented))
        return self

    def _obs(self, obstacles, num_obs):
        return _obstacles(obstacles, num_obs, 6 if self.oriented else 4)

    # ---- single word: flags, cost, 150 sample points -------------------------------------
    def word(self, word, start, end, r_min, obstacles=None, num_obs=0):
        obs, m = self._obs(obstacles if obstacles is not None else [], [num_obs])
        cost = C.c_float()
        xs = np.zeros(150, np.float32)
        ys = np.zeros(150, np.float32)
        fl = self.lib.gmto_word(word, np.ascontiguousarray(start, np.float32),
                                np.ascontiguousarray(end, np.float32), int(r_min), obs, m,
                                C.byref(cost), xs.ctypes.data, ys.ctypes.data)
        return fl, cost.value, xs, ys

    def route_edges(self, states, route, radius, obstacles, num_obs):
        """Trajectory of a goal-first route, start -> goal: (from_to, word, arcs3) like gmt_get_route_edges."""
        obs, m = self._obs(obstacles, num_obs)
        states = np.ascontiguousarray(states, np.float32)
        path = list(route)[::-1]
        ft, words, arcs = [], [], []
        for a, b in zip(path[:-1], path[1:]):
            bw, bc, ba = 255, np.inf, [np.nan] * 3
            for w in range(self.words):
                cost = C.c_float()
                a3 = np.zeros(3, np.float32)
                fl = self.lib.gmto_word_arcs(w, states[a], states[b], int(radius), obs, m, C.byref(cost), a3)
                if (fl & 1) and not (fl & 2) and cost.value < bc:
                    bw, bc, ba = w, cost.value, a3.tolist()
            ft.append((a, b)); words.append(bw); arcs.append(ba)
        return (np.array(ft, np.int32).reshape(-1, 2), np.array(words, np.uint8),
                np.array(arcs, np.float32).reshape(-1, 3))

    def edge_costs(self, states, y, x, radius, obstacles, num_obs):
        obs, m = self._obs(obstacles, num_obs)
        states = np.ascontiguousarray(states, np.float32)
        y = np.ascontiguousarray(y, np.int32)
        x = np.ascontiguousarray(x, np.int32)
        cost4 = np.zeros((len(y), self.words), np.float32)
        bits = np.zeros(len(y), np.uint32)
        best = np.zeros(len(y), np.float32)
        self.lib.gmto_edge_costs(states, y, x, len(y), float(radius), obs, m, cost4, bits, best)
        return cost4, bits, best

    def plan(self, states, neighbors, num_neighbors, start, goal, radius, threshold, obstacles, num_obs,
             iter_limit=1000, trace=True):
        """Whole plan in C.  Returns a dict with status/iterations/cost/parent/route and the trace."""
        states = np.ascontiguousarray(states, np.float32)
        neighbors = np.ascontiguousarray(neighbors, np.int32)
        num_neighbors = np.ascontiguousarray(num_neighbors, np.int32)
        obs, m = self._obs(obstacles, num_obs)
        n = states.shape[0]
        cost = np.zeros(n, np.float32)
        parent = np.zeros(n, np.int32)
        opn = np.zeros(n, np.int32)
        unexp = np.zeros(n, np.int32)
        iters = C.c_int()
        cap_it = int(iter_limit) if trace else 0
        sizes = np.full((max(cap_it, 1), 3), -2, np.int32)
        sets_cap = 0
        sets = np.zeros(1, np.int32)
        if trace:
            sets_cap = int(4 * n + 64)
            sets = np.zeros(sets_cap, np.int32)
        sets_len = C.c_long()
        status = self.lib.gmto_plan(states, n, neighbors, num_neighbors, int(start), int(goal),
                                    float(np.float32(radius)), float(np.float32(threshold)), obs, m,
                                    int(iter_limit), cost, parent, opn, unexp, C.byref(iters),
                                    sizes.ctypes.data if trace else None, cap_it,
                                    sets.ctypes.data if trace else None, sets_cap, C.byref(sets_len))
        out = dict(status=status, iterations=iters.value, cost=cost, parent=parent, open=opn,
                   unexplored=unexp)
        route = np.zeros(n + 1, np.int32)
        out["route"] = []
        if int(goal) >= 0 and (parent[int(goal)] != -1 or status == 0):
            ln = self.lib.gmto_get_path(parent, n, int(goal), route, n + 1)
            out["route"] = [int(v) for v in route[:ln]]
        if trace:
            out["trace"] = _split_trace(sizes[:iters.value], sets[:sets_len.value])
        return out

    # ---- kernel-level backend used by OracleGMT -------------------------------------------
    def k_wavefront(self, G, opn, cost, threshold, n, tpb):
        self.lib.gmto_wavefront(G, opn, cost, float(threshold[0]), n)

    def k_grow_threshold(self, threshold, radius):
        threshold[0] = self.lib.gmto_grow_threshold(float(threshold[0]), float(radius[0]))

    def k_neighbor_indicator(self, xind, G, unexplored, neighbors, num_neighbors, neighbors_index, gSize, tpb):
        self.lib.gmto_neighbor_indicator(xind, G, unexplored, neighbors, num_neighbors, neighbors_index, gSize)

    def k_compact(self, out, scan, indicator, waypoints, n, tpb):
        sel = indicator == 1
        out[scan[sel]] = waypoints[sel]

    def k_dubin_connection(self, cost, parent, x, y, states, opn, unexplored, xSize, ySize, obstacles,
                           num_obs, radius, tpb):
        self.lib.gmto_dubin_connection(cost, parent, x, y, states, opn, unexplored, xSize, ySize,
                                       obstacles, num_obs, float(radius[0]))


def _split_trace(sizes, sets):
    trace, pos = [], 0
    full = True
    for g, y, x in sizes:
        ent = {"gSize": int(g), "ySize": int(y), "xSize": int(x)}
        need = int(g) + max(int(y), 0) + max(int(x), 0)
        if full and pos + need <= len(sets):
            ent["G"] = sets[pos:pos + g].copy(); pos += int(g)
            if y > 0:
                ent["Y"] = sets[pos:pos + y].copy(); pos += int(y)
            if x > 0:
                ent["X"] = sets[pos:pos + x].copy(); pos += int(x)
        else:
            full = False
        trace.append(ent)
    return trace


class RefHost:
    """The reference's own kernels compiled for the host (oracle/_ref/libref_host.so)."""

    PATH = os.path.join(HERE, "_ref", "libref_host.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        if not self.available():
            raise FileNotFoundError(self.PATH)
        L = self.lib = C.CDLL(self.PATH)
        L.ref_num_threads.restype = C.c_int
        L.ref_dubin_connection.restype = None
        L.ref_dubin_connection.argtypes = [_f32p, _i32p, _i32p, _i32p, _f32p, _i32p, _i32p, C.c_int,
                                           C.c_int, _f32p, C.c_int, C.c_float, C.c_int]
        L.ref_wavefront.restype = None
        L.ref_wavefront.argtypes = [_i32p, _i32p, _f32p, _f32p, C.c_int, C.c_int]
        L.ref_grow_threshold.restype = None
        L.ref_grow_threshold.argtypes = [_f32p, _f32p]
        L.ref_neighbor_indicator.restype = None
        L.ref_neighbor_indicator.argtypes = [_i32p, _i32p, _i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int]
        L.ref_compact.restype = None
        L.ref_compact.argtypes = [_i32p, _i32p, _i32p, _i32p, C.c_int, C.c_int]
        L.ref_word.restype = C.c_int
        L.ref_word.argtypes = [C.c_int, _f32p, _f32p, C.c_int, _f32p, C.c_int, C.POINTER(C.c_float)]
        L.ref_edge_costs.restype = None
        L.ref_edge_costs.argtypes = [_f32p, _i32p, _i32p, C.c_long, C.c_float, _f32p, C.c_int, _f32p,
                                     _u32p, _f32p]

    def num_threads(self):
        return int(self.lib.ref_num_threads())

    def set_num_threads(self, n):
        self.lib.ref_set_num_threads(int(n))
        return self.num_threads()

    def word(self, word, start, end, r_min, obstacles=None, num_obs=0):
        obs, m = _obstacles(obstacles if obstacles is not None else [], [num_obs])
        cost = C.c_float()
        fl = self.lib.ref_word(word, np.ascontiguousarray(start, np.float32),
                               np.ascontiguousarray(end, np.float32), int(r_min), obs, m, C.byref(cost))
        return fl, cost.value

    def edge_costs(self, states, y, x, radius, obstacles, num_obs):
        obs, m = _obstacles(obstacles, num_obs)
        states = np.ascontiguousarray(states, np.float32)
        y = np.ascontiguousarray(y, np.int32)
        x = np.ascontiguousarray(x, np.int32)
        cost4 = np.zeros((len(y), 4), np.float32)
        bits = np.zeros(len(y), np.uint32)
        best = np.zeros(len(y), np.float32)
        self.lib.ref_edge_costs(states, y, x, len(y), float(radius), obs, m, cost4, bits, best)
        return cost4, bits, best

    def k_wavefront(self, G, opn, cost, threshold, n, tpb):
        self.lib.ref_wavefront(G, opn, cost, threshold, n, tpb)

    def k_grow_threshold(self, threshold, radius):
        self.lib.ref_grow_threshold(threshold, radius)

    def k_neighbor_indicator(self, xind, G, unexplored, neighbors, num_neighbors, neighbors_index, gSize, tpb):
        self.lib.ref_neighbor_indicator(xind, G, unexplored, neighbors, num_neighbors, neighbors_index, gSize, tpb)

    def k_compact(self, out, scan, indicator, waypoints, n, tpb):
        self.lib.ref_compact(out, scan, indicator, waypoints, n, tpb)

    def k_dubin_connection(self, cost, parent, x, y, states, opn, unexplored, xSize, ySize, obstacles,
                           num_obs, radius, tpb):
        self.lib.ref_dubin_connection(cost, parent, x, y, states, opn, unexplored, xSize, ySize,
                                      obstacles, num_obs, float(radius[0]), tpb)


def _exclusive_scan(a):
    """pycuda.scan.ExclusiveScanKernel(np.int32, "a+b", 0) (gmt_planner.py:17): exact int prefix sum."""
    out = np.zeros_like(a)
    if len(a) > 1:
        np.cumsum(a[:-1], out=out[1:])
    return out


class OracleGMT(object):
    """Behavioural restatement of reference class GMT (carla/gmt_planner.py:22-239).

    "Device" arrays are host NumPy arrays; every kernel launch of the reference becomes one
    backend call with the same arguments.  ``trace`` collects, per loop turn, the G / Y / X
    lists and sizes (what ``debug=True`` prints in the reference, gmt_planner.py:206-216).
    """

    def __init__(self, init_parameters, backend, debug=False):
        self.backend = backend
        # _cpu_init, gmt_planner.py:31-44
        self.states = np.ascontiguousarray(init_parameters['states'], np.float32)
        self.n = self.states.shape[0]
        self.waypoints = np.arange(self.n).astype(np.int32)
        self.neighbors = np.ascontiguousarray(init_parameters['neighbors'], np.int32)
        self.num_neighbors = np.ascontiguousarray(init_parameters['num_neighbors'], np.int32)
        self.cost = np.full(self.n, np.inf).astype(np.float32)
        self.Vunexplored = np.full(self.n, 1).astype(np.int32)
        self.Vopen = np.zeros_like(self.Vunexplored).astype(np.int32)
        self.threadsPerBlock = init_parameters.get('threadsPerBlock', 128)
        # _gpu_init, gmt_planner.py:50-61
        self.neighbors_index = _exclusive_scan(self.num_neighbors)
        self.dev_Gindicator = np.zeros(self.n, np.int32)
        self.route = []
        self.start = 0
        self.trace = []

    def step_init(self, iter_parameters):
        # gmt_planner.py:63-101
        self.cost[self.start] = np.inf
        self.Vunexplored[self.start] = 1
        self.Vopen[self.start] = 0
        if self.start != iter_parameters['start'] and len(self.route) > 2:
            del self.route[-1]
        self.obstacles, self.num_obs = _obstacles(iter_parameters['obstacles'], iter_parameters['num_obs'])
        self.parent = np.full(self.n, -1).astype(np.int32)
        self.start = iter_parameters['start']
        self.goal = iter_parameters['goal']
        self.radius = iter_parameters['radius']
        self.threshold = np.array([iter_parameters['threshold']]).astype(np.float32)
        self.cost[self.start] = 0
        self.Vunexplored[self.start] = 0
        self.Vopen[self.start] = 1
        self.dev_radius = np.array([self.radius]).astype(np.float32)
        self.dev_threshold = self.threshold.copy()
        self.dev_parent = self.parent.copy()
        self.dev_cost = self.cost.copy()
        self.dev_open = self.Vopen.copy()
        self.dev_unexplored = self.Vunexplored.copy()

    def get_path(self):
        # gmt_planner.py:103-107
        p = self.goal
        while p != -1:
            self.route.append(int(p))
            p = self.parent[p]

    def run_step(self, iter_parameters, iter_limit=1000, debug=False, time=False, budget_s=None):
        # gmt_planner.py:109-204.  budget_s (bench only): stop after that many seconds of
        # dubinConnection work with status 2 -- a bounded SAMPLE of the plan, not a plan.
        import time as _time
        b, n, tpb = self.backend, self.n, self.threadsPerBlock
        self.step_init(iter_parameters)
        self.trace = []
        self.status = None
        self.edge_pairs = 0
        self.connect_seconds = 0.0
        iteration = 0
        while True:
            if budget_s is not None and self.connect_seconds > budget_s:
                self.status = 2
                self.iterations = iteration
                return self.route
            iteration += 1
            b.k_wavefront(self.dev_Gindicator, self.dev_open, self.dev_cost, self.dev_threshold, n, tpb)
            b.k_grow_threshold(self.dev_threshold, self.dev_radius)
            goal_reached = self.dev_Gindicator[self.goal] == 1
            dev_Gscan = _exclusive_scan(self.dev_Gindicator)
            gSize = int(dev_Gscan[-1] + self.dev_Gindicator[-1])
            ent = {"gSize": gSize, "ySize": -1, "xSize": -1,
                   "G": np.flatnonzero(self.dev_Gindicator == 1).astype(np.int32)}
            self.trace.append(ent)
            self.iterations = iteration
            if iteration >= iter_limit:
                self.status = 1
                self.parent = self.dev_parent.copy()
                if self.parent[self.goal] != -1:
                    self.route = []
                    self.get_path()
                return self.route
            elif goal_reached:
                self.status = 0
                self.parent = self.dev_parent.copy()
                self.route = []
                self.get_path()
                return self.route
            elif gSize == 0:
                continue
            dev_G = np.zeros(gSize, np.int32)
            b.k_compact(dev_G, dev_Gscan, self.dev_Gindicator, self.waypoints, n, tpb)
            dev_yscan = _exclusive_scan(self.dev_open)
            ySize = int(dev_yscan[-1] + self.dev_open[-1])
            dev_y = np.zeros(ySize, np.int32)
            b.k_compact(dev_y, dev_yscan, self.dev_open, self.waypoints, n, tpb)
            ent["ySize"], ent["Y"] = ySize, dev_y.copy()
            dev_xindicator = np.zeros(n, np.int32)
            b.k_neighbor_indicator(dev_xindicator, dev_G, self.dev_unexplored, self.neighbors,
                                   self.num_neighbors, self.neighbors_index, gSize, tpb)
            dev_xscan = _exclusive_scan(dev_xindicator)
            xSize = int(dev_xscan[-1] + dev_xindicator[-1])
            ent["xSize"] = xSize
            if xSize == 0:
                continue
            dev_x = np.zeros(xSize, np.int32)
            b.k_compact(dev_x, dev_xscan, dev_xindicator, self.waypoints, n, tpb)
            ent["X"] = dev_x.copy()
            t0 = _time.perf_counter()
            b.k_dubin_connection(self.dev_cost, self.dev_parent, dev_x, dev_y, self.states, self.dev_open,
                                 self.dev_unexplored, xSize, ySize, self.obstacles, self.num_obs,
                                 self.dev_radius, tpb)
            self.connect_seconds += _time.perf_counter() - t0
            self.edge_pairs += xSize * ySize
