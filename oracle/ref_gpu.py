"""oracle/ref_gpu.py -- TEST INFRASTRUCTURE ONLY.

Runs the reference's own kernels (carla/kernel.py, compiled unchanged for sm_100a into
oracle/_ref/ref_kernel.cubin by oracle/build_oracle.py -- the build PyCuda's SourceModule would
do) on the GPU through cuda-python's driver bindings, since PyCuda itself is not installable
here.  It is the bit-exact oracle of the GPU parity tests: same libdevice, same hardware.

``RefGpu`` offers the kernel-level backend interface of oracle/gmt_oracle.py:OracleGMT (each call
uploads the NumPy "device" arrays, launches the reference kernel with the reference's launch
shape, and downloads them again) plus ``edge_costs`` through the per-word harness kernel.
"""
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
CUBIN = os.path.join(HERE, "_ref", "ref_kernel.cubin")


class RefGpu:
    @staticmethod
    def available():
        return os.path.exists(CUBIN)

    def __init__(self):
        import torch
        from cuda.bindings import driver as drv
        self.torch, self.drv = torch, drv
        torch.cuda.init()
        torch.zeros(1, device="cuda")                      # makes the primary context current
        err, self.mod = drv.cuModuleLoadData(open(CUBIN, "rb").read())
        self._ck(err, "cuModuleLoadData")
        self.fn = {}
        for name in ("dubinConnection", "wavefront", "neighborIndicator", "compact", "growThreshold",
                     "refh_edge_costs"):
            err, f = drv.cuModuleGetFunction(self.mod, name.encode())
            self._ck(err, "cuModuleGetFunction " + name)
            self.fn[name] = f

    def _ck(self, err, what):
        if int(err) != 0:
            raise RuntimeError("%s failed: %s" % (what, err))

    def _launch(self, name, grid, block, args):
        """args: list of torch tensors (passed as device pointers) or numpy scalars (by value)."""
        holders = []
        for a in args:
            if hasattr(a, "data_ptr"):
                holders.append(np.array([a.data_ptr()], dtype=np.uint64))
            else:
                holders.append(np.atleast_1d(a))
        ptrs = np.array([h.ctypes.data for h in holders], dtype=np.uint64)
        err, = self.drv.cuLaunchKernel(self.fn[name], int(grid), 1, 1, int(block), 1, 1, 0, 0,
                                       ptrs.ctypes.data, 0)
        self._ck(err, "cuLaunchKernel " + name)
        self.torch.cuda.synchronize()

    def _dev(self, a):
        return self.torch.from_numpy(np.ascontiguousarray(a)).cuda()

    # ---- per-word harness ------------------------------------------------------------------------
    def edge_costs(self, states, y, x, radius, obstacles, num_obs):
        t = self.torch
        m = int(np.asarray(num_obs).reshape(-1)[0])
        obs = np.ascontiguousarray(np.asarray(obstacles, np.float32).reshape(-1, 4))
        if obs.shape[0] == 0:
            obs = np.full((1, 4), -1, np.float32)
        P = len(y)
        d_s, d_y, d_x, d_o = self._dev(np.asarray(states, np.float32)), self._dev(np.asarray(y, np.int32)), \
            self._dev(np.asarray(x, np.int32)), self._dev(obs)
        c4 = t.zeros(P * 4, dtype=t.float32, device="cuda")
        bits = t.zeros(P, dtype=t.int32, device="cuda")
        best = t.zeros(P, dtype=t.float32, device="cuda")
        self._launch("refh_edge_costs", (P + 127) // 128, 128,
                     [d_s, d_y, d_x, np.int32(P), np.float32(radius), d_o, np.int32(m), c4, bits, best])
        return (c4.cpu().numpy().reshape(P, 4), bits.cpu().numpy().view(np.uint32), best.cpu().numpy())

    # ---- OracleGMT backend (launch shapes as gmt_planner.py:124,129,161,180,201) -----------------
    def k_wavefront(self, G, opn, cost, threshold, n, tpb):
        d = [self._dev(G), self._dev(opn), self._dev(cost), self._dev(threshold), self._dev(np.array([n], np.int32))]
        self._launch("wavefront", (n + tpb - 1) // tpb, tpb, d)
        G[:] = d[0].cpu().numpy()

    def k_grow_threshold(self, threshold, radius):
        d = [self._dev(threshold), self._dev(radius)]
        self._launch("growThreshold", 1, 1, d)
        threshold[:] = d[0].cpu().numpy()

    def k_neighbor_indicator(self, xind, G, unexplored, neighbors, num_neighbors, neighbors_index, gSize, tpb):
        d = [self._dev(xind), self._dev(G), self._dev(unexplored), self._dev(neighbors), self._dev(num_neighbors),
             self._dev(neighbors_index), self._dev(np.array([gSize], np.int32))]
        self._launch("neighborIndicator", (gSize + tpb - 1) // tpb, tpb, d)
        xind[:] = d[0].cpu().numpy()

    def k_compact(self, out, scan, indicator, waypoints, n, tpb):
        d = [self._dev(out), self._dev(scan), self._dev(indicator), self._dev(waypoints),
             self._dev(np.array([n], np.int32))]
        self._launch("compact", (n + tpb - 1) // tpb, tpb, d)
        out[:] = d[0].cpu().numpy()

    def k_dubin_connection(self, cost, parent, x, y, states, opn, unexplored, xSize, ySize, obstacles,
                           num_obs, radius, tpb):
        d = [self._dev(cost), self._dev(parent), self._dev(x), self._dev(y), self._dev(states), self._dev(opn),
             self._dev(unexplored), self._dev(np.array([xSize], np.int32)), self._dev(np.array([ySize], np.int32)),
             self._dev(obstacles), self._dev(np.array([num_obs], np.int32)),
             self._dev(np.asarray(radius, np.float32).reshape(1))]
        self._launch("dubinConnection", (xSize + tpb - 1) // tpb, tpb, d)
        cost[:] = d[0].cpu().numpy()
        parent[:] = d[1].cpu().numpy()
        opn[:] = d[5].cpu().numpy()
        unexplored[:] = d[6].cpu().numpy()


class RefGpuPlanner:
    """The reference's GMT.run_step (carla/gmt_planner.py:109-204) with the reference's own kernels
    (ref_kernel.cubin) and DEVICE-RESIDENT arrays, the way PyCuda runs it: every array stays on the
    GPU, the three scans are exact int32 prefix sums (pycuda ExclusiveScanKernel -> torch.cumsum), and
    the host reads back exactly what the reference's `.get()` calls read (G[goal], gSize, ySize, xSize).
    Two uses: the bit-exact oracle of the full-size parity tests (parents, cost bit patterns, iteration
    count of whole cfg2 / cfg3 / cfg4 plans) and the "reference kernels on this B200" baseline of bench.py.
    TEST / BENCH INFRASTRUCTURE: nothing of the product imports it."""

    def __init__(self, init_parameters, ref=None):
        self.ref = ref or RefGpu()
        t = self.t = self.ref.torch
        self.dev = t.device("cuda", t.cuda.current_device())
        self.states_np = np.ascontiguousarray(init_parameters['states'], np.float32)
        self.n = n = self.states_np.shape[0]
        self.tpb = int(init_parameters.get('threadsPerBlock', 128))
        num = np.ascontiguousarray(init_parameters['num_neighbors'], np.int32)
        idx = np.zeros(n, np.int32)
        if n > 1:
            np.cumsum(num[:-1], out=idx[1:])                    # ExclusiveScanKernel, gmt_planner.py:59
        d = self._dev
        self.d_states = d(self.states_np.reshape(-1))
        self.d_neighbors = d(np.ascontiguousarray(init_parameters['neighbors'], np.int32))
        self.d_num = d(num)
        self.d_nidx = d(idx)
        self.d_way = d(np.arange(n, dtype=np.int32))
        self.d_n = d(np.array([n], np.int32))
        self.d_G = t.zeros(n, dtype=t.int32, device=self.dev)

    def _dev(self, a):
        return self.t.from_numpy(np.ascontiguousarray(a)).to(self.dev)

    def _launch(self, name, count, args):
        r = self.ref
        holders = [np.array([a.data_ptr()], dtype=np.uint64) for a in args]
        ptrs = np.array([h.ctypes.data for h in holders], dtype=np.uint64)
        err, = r.drv.cuLaunchKernel(r.fn[name], int((count + self.tpb - 1) // self.tpb), 1, 1, self.tpb, 1, 1, 0, 0,
                                    ptrs.ctypes.data, 0)
        r._ck(err, "cuLaunchKernel " + name)

    def _scan(self, ind):
        """exclusive int32 prefix sum + total, like `exclusiveScan(scan); scan[-1] + ind[-1]` (gmt_planner.py:138-141)"""
        inc = self.t.cumsum(ind, 0, dtype=self.t.int32)
        return (inc - ind).contiguous(), inc[-1:].contiguous()

    def run_step(self, iter_parameters, iter_limit=1000, budget_s=None, keep_sets=False):
        """Returns a dict: status (0 goal, 1 limit, 2 budget exhausted), iterations, parent, cost, route,
        edge_pairs (sum |X| * |Y|), seconds; with keep_sets the per-iteration G / Y / X id arrays as well."""
        import time as _time
        t, n = self.t, self.n
        obs = np.ascontiguousarray(np.asarray(iter_parameters['obstacles'], np.float32).reshape(-1, 4))
        if obs.shape[0] == 0:
            obs = np.full((1, 4), -1, np.float32)
        m = int(np.asarray(iter_parameters['num_obs']).reshape(-1)[0])
        start, goal = int(iter_parameters['start']), int(iter_parameters['goal'])
        # step_init, gmt_planner.py:63-101
        cost = np.full(n, np.inf, np.float32); cost[start] = 0
        unexp = np.full(n, 1, np.int32); unexp[start] = 0
        opn = np.zeros(n, np.int32); opn[start] = 1
        d_cost, d_unexp, d_open = self._dev(cost), self._dev(unexp), self._dev(opn)
        d_parent = self._dev(np.full(n, -1, np.int32))
        d_obs, d_m = self._dev(obs.reshape(-1)), self._dev(np.array([m], np.int32))
        d_radius = self._dev(np.array([iter_parameters['radius']], np.float32))
        d_thr = self._dev(np.array([iter_parameters['threshold']], np.float32))
        out = dict(status=None, iterations=0, edge_pairs=0, sets=[])
        t.cuda.synchronize()
        t0 = _time.perf_counter()
        iteration = 0
        while True:
            if budget_s is not None and _time.perf_counter() - t0 > budget_s:
                out["status"] = 2
                break
            iteration += 1
            self._launch("wavefront", n, [self.d_G, d_open, d_cost, d_thr, self.d_n])
            r = self.ref
            holders = [np.array([d_thr.data_ptr()], np.uint64), np.array([d_radius.data_ptr()], np.uint64)]
            ptrs = np.array([h.ctypes.data for h in holders], dtype=np.uint64)
            err, = r.drv.cuLaunchKernel(r.fn["growThreshold"], 1, 1, 1, 1, 1, 1, 0, 0, ptrs.ctypes.data, 0)
            r._ck(err, "cuLaunchKernel growThreshold")
            goal_reached = goal >= 0 and int(self.d_G[goal].item()) == 1
            Gscan, d_gSize = self._scan(self.d_G)
            gSize = int(d_gSize.item())
            ent = None
            if keep_sets:
                ent = {"gSize": gSize, "ySize": -1, "xSize": -1, "G": t.nonzero(self.d_G == 1).flatten().int().cpu().numpy()}
                out["sets"].append(ent)
            if iteration >= iter_limit:
                out["status"] = 1
                break
            if goal_reached:
                out["status"] = 0
                break
            if gSize == 0:
                continue
            d_Gl = t.zeros(gSize, dtype=t.int32, device=self.dev)
            self._launch("compact", n, [d_Gl, Gscan, self.d_G, self.d_way, self.d_n])
            yscan, d_ySize = self._scan(d_open)
            ySize = int(d_ySize.item())
            d_y = t.zeros(ySize, dtype=t.int32, device=self.dev)
            self._launch("compact", n, [d_y, yscan, d_open, self.d_way, self.d_n])
            d_xind = t.zeros(n, dtype=t.int32, device=self.dev)
            self._launch("neighborIndicator", gSize, [d_xind, d_Gl, d_unexp, self.d_neighbors, self.d_num, self.d_nidx, d_gSize])
            xscan, d_xSize = self._scan(d_xind)
            xSize = int(d_xSize.item())
            if ent is not None:
                ent["ySize"], ent["Y"], ent["xSize"] = ySize, d_y.cpu().numpy(), xSize
            if xSize == 0:
                continue
            d_x = t.zeros(xSize, dtype=t.int32, device=self.dev)
            self._launch("compact", n, [d_x, xscan, d_xind, self.d_way, self.d_n])
            if ent is not None:
                ent["X"] = d_x.cpu().numpy()
            self._launch("dubinConnection", xSize, [d_cost, d_parent, d_x, d_y, self.d_states, d_open, d_unexp, d_xSize,
                                                    d_ySize, d_obs, d_m, d_radius])
            out["edge_pairs"] += xSize * ySize
        t.cuda.synchronize()
        out["seconds"] = _time.perf_counter() - t0
        out["iterations"] = iteration
        out["parent"] = d_parent.cpu().numpy()
        out["cost"] = d_cost.cpu().numpy()
        route = []
        if out["status"] in (0, 1) and goal >= 0 and (out["status"] == 0 or out["parent"][goal] != -1):
            p = goal
            while p != -1 and len(route) <= n:
                route.append(int(p))
                p = int(out["parent"][p])
        out["route"] = route
        return out
