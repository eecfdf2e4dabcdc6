"""Run under torchrun on N GPUs: one roadmap sharded by node range, NCCL all-reduce(min) per wavefront;
rank 0 compares with the single-GPU plan and prints timings.  usage: sharded_check.py [n] [m]"""
import os, sys, time
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import torch, torch.distributed as dist
import multi_gpu, synthetic_world as sw
from gmt_planner import GMT
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 500
init, it, meta = sw.make_world(n=n, m=m, seed=23, sampler=lambda *a: sw.cuda_sampler(*a, device=local),
                               neighbor_builder=lambda s, r: sw.cuda_neighbor_builder(s, r, device=local))
g = GMT(init, device=local)
if rank == 0:
    goal = sw.pick_reachable_goal(init, it, meta["side"], explore=lambda a, b: sw.explore_cuda(a, b, device=local))
else:
    goal = 0
gt = torch.tensor([goal], device="cuda")
if world > 1:
    dist.broadcast(gt, 0)
it = dict(it, goal=int(gt.item()))
single = list(g.run_step(it, iter_limit=1000))
parent, cost, res1 = g.parent.copy(), g.dev_cost.copy(), g.last_result
t_single = res1.device_ms
for rep in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    route, res, steps = multi_gpu.run_step_sharded(g, it, iter_limit=1000, rank=rank, world=world)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
same = (list(route) == single) and np.array_equal(g.parent, parent) and np.array_equal(g.dev_cost.view(np.uint32), cost.view(np.uint32))
# fused form: one launch per GPU, exchange over NVLink peer memory inside the kernel
multi_gpu.open_peers(g, rank, world)
for rep in range(3):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    route2, res2 = multi_gpu.run_step_sharded_p2p(g, it, iter_limit=1000, rank=rank, world=world)
    torch.cuda.synchronize()
    dt2 = time.perf_counter() - t0
same2 = (list(route2) == single) and np.array_equal(g.parent, parent) and np.array_equal(g.dev_cost.view(np.uint32), cost.view(np.uint32))
multi_gpu.close_peers(g)
flags = torch.tensor([int(same), int(same2)], device="cuda")
if world > 1:
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
dev2 = multi_gpu.max_over_ranks([res2.device_ms], device="cuda")[0]
if rank == 0:
    print("  fused peer-memory exchange: %.3f ms wall, %.3f ms on device (max over ranks), %d wavefronts; identical tree on every rank: %s (NCCL path: %s)"
          % (dt2 * 1e3, dev2, res2.iterations, bool(flags[1].item()), bool(flags[0].item())))
if rank == 0:
    print("n=%d m=%d world=%d: single persistent plan %.3f ms (%d wavefronts, %d edge checks); sharded %d launches in %.3f ms wall; identical tree: %s"
          % (n, m, world, t_single, res1.iterations, res1.edge_pairs, steps, dt * 1e3, same))
if world > 1:
    dist.destroy_process_group()
