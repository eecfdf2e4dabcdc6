"""Randomised check of the sharded plan under torchrun (one process per GPU): for random worlds the NCCL form
(one wavefront per launch + all-reduce(min)) and the fused form (in-kernel NVLink exchange) must rebuild the
single-GPU tree bit for bit on every rank.  usage: torchrun ... stress_sharded.py [seconds] [seed]"""
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
budget = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)      # same stream on every rank
t0, cases, ok_all = time.time(), 0, 1
while True:
    go = torch.tensor([int(time.time() - t0 < budget)], device="cuda")
    if world > 1:
        dist.broadcast(go, 0)                                                       # ranks stop together
    if not go.item():
        break
    n = int(rng.integers(2000, 30000)); m = int(rng.integers(0, 200)); seed = int(rng.integers(0, 1 << 30))
    k_env = int(rng.choice([0, 0, 2, 4]))
    if k_env:
        os.environ["GMT_SELECT_K"] = str(k_env)
    else:
        os.environ.pop("GMT_SELECT_K", None)
    init, it, meta = sw.make_world(n=n, m=m, seed=seed, sampler=lambda *a: sw.cuda_sampler(*a, device=local),
                                   neighbor_builder=lambda s, r: sw.cuda_neighbor_builder(s, r, device=local))
    g = GMT(init, device=local)
    g.run_step(dict(it, goal=-1), iter_limit=150)
    far = g.last_result.farthest_frontier
    it = dict(it, goal=int(far) if far >= 0 else it['goal'])
    single = list(g.run_step(it, iter_limit=150))
    parent, cost = g.parent.copy(), g.dev_cost.copy()
    r1, res1, steps = multi_gpu.run_step_sharded(g, it, iter_limit=150, rank=rank, world=world)
    same = list(r1) == single and np.array_equal(g.parent, parent) and np.array_equal(g.dev_cost.view(np.uint32), cost.view(np.uint32))
    multi_gpu.open_peers(g, rank, world)
    for _ in range(2):
        r2, res2 = multi_gpu.run_step_sharded_p2p(g, it, iter_limit=150, rank=rank, world=world)
        same = same and list(r2) == single and np.array_equal(g.parent, parent) and \
            np.array_equal(g.dev_cost.view(np.uint32), cost.view(np.uint32))
    multi_gpu.close_peers(g)
    flag = torch.tensor([int(same)], device="cuda")
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok_all &= int(flag.item())
    if not flag.item() and rank == 0:
        print("MISMATCH n=%d m=%d seed=%d K=%d" % (n, m, seed, k_env))
    cases += 1
    del g
if rank == 0:
    print("sharded stress, %d GPUs: %d random worlds (2k-30k states, 0-200 boxes) in %.0f s; NCCL form and fused NVLink form "
          "(twice each) identical to the single-GPU tree on every rank: %s" % (world, cases, time.time() - t0, bool(ok_all)))
if world > 1:
    dist.destroy_process_group()
