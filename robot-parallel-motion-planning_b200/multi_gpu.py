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


# ------------------------------------------------------------------------------------------------------
# Single very large roadmap, sharded by node range (BASELINE config 5)
# ------------------------------------------------------------------------------------------------------
class _DeviceArray(object):
    """Minimal __cuda_array_interface__ holder so torch can alias memory the library owns."""

    def __init__(self, ptr, count, typestr):
        self.__cuda_array_interface__ = {"shape": (int(count),), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def best_tensor(gmt):
    """int64 torch tensor aliasing the planner's packed (cost bits << 32 | parent) array on the device.
    All keys are < 2^63 (cost >= 0), so the signed integer order is the unsigned one the kernels use."""
    import ctypes as C

    import torch
    try:
        from . import _lib
    except ImportError:
        import _lib
    ptr, count = C.c_void_p(), C.c_longlong()
    _lib.check(gmt._lib.gmt_best_dev(gmt._handle, C.byref(ptr), C.byref(count)), "gmt_best_dev")
    return torch.as_tensor(_DeviceArray(ptr.value, count.value, "<i8"), device="cuda")


def pack_keys(cost, parent):
    """The exchange format of the sharded plan as int64: float32 bits of cost (>= 0, +inf = unconnected) in the
    high word, parent id in the low word (-1 -> 0xFFFFFFFF).  Integer MIN over ranks = lowest cost, ties -> lowest
    parent id: the reference's strict-less scan over ascending y (kernel.py:427,440-443); an unconnected node
    loses to any connected one."""
    c = np.ascontiguousarray(cost, np.float32).view(np.uint32).astype(np.int64)
    q = np.ascontiguousarray(parent, np.int32).view(np.uint32).astype(np.int64)
    return (c << 32) | q


def unpack_keys(keys):
    k = np.asarray(keys, np.int64)
    cost = (k >> 32).astype(np.uint32).view(np.float32)
    parent = (k & 0xFFFFFFFF).astype(np.uint32).view(np.int32)
    return cost, parent


def allreduce_min(t):
    """The per-wavefront exchange: NCCL all-reduce(min) of the packed cost/parent array over NVLink.
    Returns only when the reduction has COMPLETED on the device: the next gmt_sharded_step launches on the
    handle's own stream, which has no dependency on the stream the collective ran on (include/gmt_c_api.h,
    "STREAM ORDERING")."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        if t.is_cuda:
            torch.cuda.current_stream(t.device).synchronize()


def run_step_sharded(gmt, iter_parameters, iter_limit=1000, rank=0, world=1, exchange=allreduce_min):
    """GMT.run_step for one roadmap sharded over `world` GPUs: this rank connects the candidate nodes with
    ids in its contiguous range; after every wavefront the packed cost/parent arrays are min-reduced.
    Every rank returns the same route.  Returns (route, gmt_result, wavefront launches)."""
    import ctypes as C

    import numpy as np
    try:
        from . import _lib
    except ImportError:
        import _lib
    lib = gmt._lib
    gmt.step_init(iter_parameters, False)
    num_obs = int(np.asarray(iter_parameters['num_obs']).reshape(-1)[0])
    obs = np.ascontiguousarray(iter_parameters['obstacles'], dtype=np.float32).reshape(-1, 4)
    _lib.check(lib.gmt_set_obstacles(gmt._handle, obs.ctypes.data if num_obs > 0 else None, num_obs), "gmt_set_obstacles")
    lo, hi = shard_range(gmt.n, rank, world)
    _lib.check(lib.gmt_sharded_begin(gmt._handle, int(gmt.start), int(gmt.goal), float(np.float32(gmt.radius)),
                                     float(gmt.threshold[0]), int(iter_limit), lo, hi), "gmt_sharded_begin")
    best = best_tensor(gmt)
    status, steps = C.c_int(-2), 0
    while True:
        _lib.check(lib.gmt_sharded_step(gmt._handle, C.byref(status)), "gmt_sharded_step")
        steps += 1
        if status.value != -2:
            break
        exchange(best)
    res = _lib.GmtResult()
    _lib.check(lib.gmt_sharded_finish(gmt._handle, C.byref(res)), "gmt_sharded_finish")
    gmt.last_result, gmt.iterations, gmt.status = res, res.iterations, res.status
    if res.status == _lib.GMT_STATUS_GOAL or res.route_len >= 0:
        gmt.route = []
        gmt.get_path()
    return gmt.route, res, steps


def open_peers(gmt, rank, world):
    """Maps every rank's cost/parent array and flag block into this rank's address space (CUDA IPC handles
    swapped with torch.distributed.all_gather_object; NVLink peer access) for run_step_sharded_p2p."""
    import ctypes as C

    import torch.distributed as dist
    try:
        from . import _lib
    except ImportError:
        import _lib
    blob = (C.c_ubyte * 128)()
    _lib.check(gmt._lib.gmt_peer_export(gmt._handle, blob), "gmt_peer_export")
    blobs = [None] * world
    if world > 1:
        dist.all_gather_object(blobs, bytes(blob))
    else:
        blobs = [bytes(blob)]
    packed = (C.c_ubyte * (128 * world)).from_buffer_copy(b"".join(blobs))
    _lib.check(gmt._lib.gmt_peer_open(gmt._handle, int(rank), int(world), packed), "gmt_peer_open")
    if world > 1:
        dist.barrier()          # nobody plans before everybody has mapped everybody


def close_peers(gmt):
    gmt._lib.gmt_peer_close(gmt._handle)


def run_step_sharded_p2p(gmt, iter_parameters, iter_limit=1000, rank=0, world=1):
    """The sharded plan as ONE kernel launch per GPU: the per-wavefront exchange of the packed cost/parent keys
    goes over NVLink peer memory inside the kernel (gmt_plan_sharded_p2p; open_peers first).  Every rank must
    call this with the same query.  Returns (route, gmt_result)."""
    import ctypes as C

    try:
        from . import _lib
    except ImportError:
        import _lib
    lib = gmt._lib
    gmt.step_init(iter_parameters, False)
    num_obs = int(np.asarray(iter_parameters['num_obs']).reshape(-1)[0])
    obs = np.ascontiguousarray(iter_parameters['obstacles'], dtype=np.float32).reshape(-1, 4)
    _lib.check(lib.gmt_set_obstacles(gmt._handle, obs.ctypes.data if num_obs > 0 else None, num_obs), "gmt_set_obstacles")
    lo, hi = shard_range(gmt.n, rank, world)
    res = _lib.GmtResult()
    _lib.check(lib.gmt_plan_sharded_p2p(gmt._handle, int(gmt.start), int(gmt.goal), float(np.float32(gmt.radius)),
                                        float(gmt.threshold[0]), int(iter_limit), lo, hi, C.byref(res)),
               "gmt_plan_sharded_p2p")
    gmt.last_result, gmt.iterations, gmt.status = res, res.iterations, res.status
    if res.status == _lib.GMT_STATUS_GOAL or res.route_len >= 0:
        gmt.route = []
        gmt.get_path()
    return gmt.route, res


def run_step_sharded_emulated(gmts, iter_parameters, iter_limit=1000):
    """The same protocol with all ranks emulated in ONE process on ONE GPU (len(gmts) handles on the same
    roadmap): used by the GPU tests, where more ranks than GPUs must not spin on each other."""
    import ctypes as C

    import numpy as np
    import torch
    try:
        from . import _lib
    except ImportError:
        import _lib
    world = len(gmts)
    bests = []
    for r, g in enumerate(gmts):
        g.step_init(iter_parameters, False)
        num_obs = int(np.asarray(iter_parameters['num_obs']).reshape(-1)[0])
        obs = np.ascontiguousarray(iter_parameters['obstacles'], dtype=np.float32).reshape(-1, 4)
        _lib.check(g._lib.gmt_set_obstacles(g._handle, obs.ctypes.data if num_obs > 0 else None, num_obs), "obs")
        lo, hi = shard_range(g.n, r, world)
        _lib.check(g._lib.gmt_sharded_begin(g._handle, int(g.start), int(g.goal), float(np.float32(g.radius)),
                                            float(g.threshold[0]), int(iter_limit), lo, hi), "begin")
        bests.append(best_tensor(g))
    steps = 0
    while True:
        stats = []
        for g in gmts:
            st = C.c_int()
            _lib.check(g._lib.gmt_sharded_step(g._handle, C.byref(st)), "step")
            stats.append(st.value)
        steps += 1
        assert len(set(stats)) == 1, "ranks disagree on the plan state: %r" % (stats,)
        if stats[0] != -2:
            break
        torch.cuda.synchronize()
        m = bests[0].clone()
        for b in bests[1:]:
            m = torch.minimum(m, b)
        for b in bests:
            b.copy_(m)
        torch.cuda.synchronize()
    out = []
    for g in gmts:
        res = _lib.GmtResult()
        _lib.check(g._lib.gmt_sharded_finish(g._handle, C.byref(res)), "finish")
        g.last_result, g.iterations, g.status = res, res.iterations, res.status
        if res.status == _lib.GMT_STATUS_GOAL or res.route_len >= 0:
            g.route = []
            g.get_path()
        out.append((list(g.route), res))
    return out, steps
