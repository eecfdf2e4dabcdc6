"""Development aid: roadmap builders (Philox sampler + cell-grid neighbour search) at BASELINE sizes; run under
ncu to get per-kernel times and DRAM bytes.  usage: roadmap_bench.py [n]"""
import sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import synthetic_world as sw
from gmt_planner import GMT
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
side = sw.world_side(n)
for rep in range(2):
    t0 = time.perf_counter(); st = sw.cuda_sampler(23, n, side); t1 = time.perf_counter()
    nbr, num = sw.cuda_neighbor_builder(st); t2 = time.perf_counter()
print("n=%d: sampler call %.1f ms, neighbour search call %.1f ms (host copies included), %d neighbour entries, mean degree %.1f"
      % (n, (t1 - t0) * 1e3, (t2 - t1) * 1e3, len(nbr), num.mean()))
init = {'states': st, 'neighbors': nbr, 'num_neighbors': num}
g = GMT(init)
obs, nobs = sw.make_obstacles(np.random.default_rng(23), 2000, side, "uniform")
g.run_step({'start': 0, 'goal': 1, 'radius': 2.0, 'threshold': 1.0, 'obstacles': obs, 'num_obs': nobs}, iter_limit=2)
print("prepared one plan (obstacle grid + node preparation kernels ran)")
