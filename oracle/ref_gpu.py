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
