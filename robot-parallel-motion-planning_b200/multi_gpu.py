"""multi_gpu.py -- host-side plumbing for the path's multi-GPU shape.

Independent start/goal queries shard naturally: the roadmap and the obstacle set are replicated on
every GPU, the query list is split contiguously by rank, and there is NO collective on the data
path -- routes are gathered on the host afterwards.  torch.distributed (NCCL on GPUs, gloo in the
CPU tests) is used only for the barrier, the max-over-ranks of the timings and the final gather.
"""
import numpy as np


def shard_range(q, rank, world):
    """Contiguous [lo, hi) slice of q queries for `rank` of `world`; sizes differ by at most one."""
    q, rank, world = int(q), int(rank), int(world)
    if world < 1 or not (0 <= rank < world):
        raise ValueError("rank %d of world %d" % (rank, world))
    base, extra = divmod(q, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_queries(starts, goals, rank, world):
    lo, hi = shard_range(len(starts), rank, world)
    return np.asarray(starts)[lo:hi], np.asarray(goals)[lo:hi]


def max_over_ranks(values, device=None):
    """Element-wise max of a small float vector over all ranks (timings are reported as the slowest rank)."""
    import torch
    import torch.distributed as dist
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_routes(routes):
    """All ranks' route lists, in query order, on every rank (host objects: no device collective)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(routes)
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, list(routes))
    return [r for part in out for r in part]


def plan_queries(gmt, iter_parameters, starts, goals, iter_limit=1000):
    """Plans this rank's queries back to back on one GMT instance (obstacles as in iter_parameters)."""
    routes = []
    for s, g in zip(starts, goals):
        routes.append(list(gmt.run_step(dict(iter_parameters, start=int(s), goal=int(g)), iter_limit=iter_limit)))
    return routes
