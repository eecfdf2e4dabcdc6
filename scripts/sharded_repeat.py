"""Run under torchrun on N GPUs: the cfg5-shaped sharded plan repeated many times; EVERY repetition's tree (route,
parents, cost bit patterns) is compared with the single-GPU tree on every rank, and a mismatch is described (how many
nodes, at which cost) -- a race in the exchange shows up as an occasional difference, which one run cannot exclude.
usage: sharded_repeat.py [n] [m] [fused reps] [nccl reps]"""
import os, sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'robot-parallel-motion-planning_b200')
import torch, torch.distributed as dist
import multi_gpu, synthetic_world as sw
from gmt_planner import GMT
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
m = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
reps_fused = int(sys.argv[3]) if len(sys.argv) > 3 else 30
reps_nccl = int(sys.argv[4]) if len(sys.argv) > 4 else 3
init, it, meta = sw.make_world(n=n, m=m, seed=23, sampler=lambda *a: sw.cuda_sampler(*a, device=local),
                               neighbor_builder=lambda s, r: sw.cuda_neighbor_builder(s, r, device=local))
g = GMT(init, device=local)
g.run_step(dict(it, goal=-1), iter_limit=1000)
far = torch.tensor([g.last_result.farthest_frontier], device="cuda")
if world > 1:
    dist.broadcast(far, 0)
it = dict(it, goal=int(far.item()))
single = list(g.run_step(it, iter_limit=1000))
parent, cost, res1 = g.parent.copy(), g.dev_cost.copy(), g.last_result


def describe(tag, rep, route):
    dp = np.flatnonzero(g.parent != parent)
    dc = np.flatnonzero(g.dev_cost.view(np.uint32) != cost.view(np.uint32))
    if len(dp) or len(dc) or list(route) != single:
        lo = float(np.min(cost[dc])) if len(dc) else float('nan')
        print("MISMATCH rank %d %s rep %d: %d parents, %d costs differ (lowest reference cost among them %.3f), route equal %s, "
              "iterations %d vs %d" % (rank, tag, rep, len(dp), len(dc), lo, list(route) == single, g.iterations, res1.iterations), flush=True)
        return 0
    return 1


ok = 1
times = []
multi_gpu.open_peers(g, rank, world)
for rep in range(reps_fused):
    if world > 1:
        dist.barrier()
    route, res = multi_gpu.run_step_sharded_p2p(g, it, iter_limit=1000, rank=rank, world=world)
    times.append(res.device_ms)
    ok &= describe("fused", rep, route)
multi_gpu.close_peers(g)
for rep in range(reps_nccl):
    if world > 1:
        dist.barrier()
    route, res, steps = multi_gpu.run_step_sharded(g, it, iter_limit=1000, rank=rank, world=world)
    ok &= describe("nccl", rep, route)
flag = torch.tensor([ok], device="cuda")
if world > 1:
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("n=%d m=%d world=%d: single %.3f ms (%d wavefronts); %d fused + %d NCCL-form plans, every tree identical on every rank: %s; "
          "fused ms: first %.2f, median %.2f, min %.2f" % (n, m, world, res1.device_ms, res1.iterations, reps_fused, reps_nccl,
                                                          bool(flag.item()), times[0], float(np.median(times)), min(times)))
if world > 1:
    dist.destroy_process_group()
