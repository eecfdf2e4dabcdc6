#!/usr/bin/env python
"""bench.py -- GMT* plans/sec on B200 (BASELINE.json metric), ONE JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--only NAME] [--words 6] [--oriented]

Headline workload (the same at every N, so the driver's 1/2/4/8 run is a strong-scaling curve): BASELINE
config 4 -- a batch of 4096 independent start/goal queries on the 10k-sample / 100-obstacle map, the query list
split contiguously over the ranks, no collective on the data path.  It is the largest configuration of the metric
that fits one GPU.  A "step" is one pass of the hot path over this rank's share of the batch (ONE launch of the
persistent plan kernel in team mode).

* value     plans/s with roadmap, obstacle set AND query list resident in HBM (gmt_set_queries once,
            gmt_plan_queries per step), CUDA events on the launching stream (an explicit torch stream the handle
            is told to use), L2 flushed before every timed step, max over ranks.
* e2e       the same through the host-buffer call gmt_plan_batch (numpy start/goal arrays in, result structs and
            routes out), wall clock around the call, copies inside.
* sub       records of the other BASELINE configurations, measured in the same run:
            N = 1: cfg2 / cfg3 single-query latency (resident and through GMT.run_step with host buffers), "cfg3u"
            (100k uniform samples, 500 obstacles, a goal ~150 wavefronts away: the driver-timed form of the
            "100k-sample plan < 10 ms" target), the edge kernel (Dubins edge checks/s), the reference's own kernels
            on this GPU (gpu_reference, with a bit-exact comparison of the two trees) and on the host cores;
            every N: cfg5, ONE plan on the 1M-sample roadmap -- single GPU at N = 1, node-range sharding with the
            in-kernel NVLink exchange at N > 1 (the NCCL all-reduce(min) form is timed beside it).
* roofline  dominant kernel gmt_plan_kernel: algorithmic FP32 work (reference edge checks x F_edge(m))
            over its CUDA-event time against the FP32 FMA peak measured live; `executed_frac` = the same with the
            work the kernel actually did (pairs evaluated, words swept); `ncu` = issue-slot / pipe utilisation of
            the committed ncu capture of this kernel (profiles/).
* cpu_baseline  the reference's kernels compiled for the host (oracle/_ref, kind "reference"; else the C port) on
            this box's cores, on a bounded sample of the batch.
--impl reference times that CPU implementation alone (rank 0 only), same config dict, metric and unit.
--only cfg1|cfg2|cfg3|cfg3u|cfg4|cfg5 restricts the run to one workload (development).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

# stdout carries ONE JSON line: keep NCCL's version banner (printed to stdout at NCCL_DEBUG=VERSION) out of it
if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "robot-parallel-motion-planning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "gmt_plans_per_sec"
UNIT = "plans/s"
Q_BATCH = 4096
ROUTE_STRIDE = 128
WORKLOADS = {
    "cfg1": "cfg1: single query, 2k samples, 20 obstacles",
    "cfg2": "cfg2: single query, 10k samples, 100 obstacles",
    "cfg3": "cfg3: single query, 100k samples, road lattice, 2k obstacles",
    "cfg3u": "cfg3u: single query, 100k uniform samples, 500 obstacles, goal ~150 wavefronts away",
    "cfg4": "cfg4: batch of 4096 independent start/goal queries on a 10k-sample / 100-obstacle map, split over the "
            "GPUs by rank, no collective",
    "cfg5": "cfg5: single query on ONE 1M-sample roadmap (2k obstacles), edge evaluation sharded by node range over "
            "the GPUs, per-wavefront exchange of the packed cost/parent keys over NVLink",
}
PHASES = ("frontier", "select", "connect", "exchange", "exchange_push", "exchange_grid_barrier", "exchange_wait_for_peers")
SINGLE_CONFIG = {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg5": 5}


def f_edge(m):
    """Algorithmic FP32 op-equivalents per reference edge check: 4*(1142 + 604 m)."""
    return 4.0 * (1142.0 + 604.0 * m)


def f_executed(evals, checks, nw=4):
    """Op-equivalents of the work a plan actually did, in SURVEY 8d's units: every evaluated pair pays the
    tangent construction + cost of its words (1142 - 401 - 401 - 254 = 86 per word without the sample points);
    every swept word pays its 150 sample points (401 + 401 + 254) and, against the obstacle grid, about one cell
    lookup (~12) per point instead of 4 compares x m boxes."""
    return float(evals) * nw * 86.0 + float(checks) * (1056.0 + 150.0 * 12.0)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def headline_config():
    """The `config` object of the JSON line: static, identical in both arms."""
    return {"workload": WORKLOADS["cfg4"], "n": 10000, "m": 100, "seed": 22, "queries": Q_BATCH, "iter_limit": 1000,
            "radius": 2.0, "threshold": 1.0, "goals": "farthest state that enters a wavefront from each start",
            "sub_workloads": [WORKLOADS[k] for k in ("cfg2", "cfg3", "cfg3u", "cfg5")]}


# --------------------------------------------------------------------------------------------------
# clocks sampling (B200_PROFILING.md): nvidia-smi polled during the timed region
# --------------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")
    NAMES = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")

    def __init__(self, gpu_id):
        self.gpu_id, self.proc, self.path, self.skip = gpu_id, None, None, 0

    def start(self):
        """Starts polling (50 ms) and waits until the first sample is there (nvidia-smi needs a moment)."""
        try:
            fd, self.path = tempfile.mkstemp(prefix="gmt_clocks_", suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_id), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 3.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples taken before this call (warm-up) are not part of the timed region."""
        try:
            self.skip = sum(1 for _ in open(self.path))
        except Exception:
            self.skip = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        try:
            self.proc.terminate()
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        try:
            lines = open(self.path).read().splitlines()
            if len(lines) > self.skip:        # keep only the timed region unless it was too short to be sampled
                lines = lines[self.skip:]
            for line in lines:
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for name, val in zip(self.NAMES, parts[3:7]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=float(np.median(sm)), sm_max_mhz=float(max(mx)), reasons=sorted(reasons),
                       samples=len(sm))
        return out


def gpu_uuid(torch, dev, local):
    props = torch.cuda.get_device_properties(dev)
    return "GPU-" + str(props.uuid) if hasattr(props, "uuid") else str(local)


def profile_json(name):
    path = os.path.join(ROOT, "profiles", name)
    if os.path.exists(path):
        try:
            return json.load(open(path))
        except Exception:
            return {}
    return {}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------------
# worlds
# --------------------------------------------------------------------------------------------------
def make_named_world(name, cpu=False, device=0, with_neighbors=True):
    """(init, it, meta) of a named workload; cpu=True builds it with the NumPy restatements (bit-identical to the
    CUDA builders, tests/test_world_oracle.py + tests/test_gpu_parity.py) so that the CPU arm needs no GPU."""
    import synthetic_world as sw
    kw = {}
    if cpu:
        from oracle.world_oracle import build_neighbors, sample_states
        kw = dict(sampler=sample_states, neighbor_builder=build_neighbors)
    else:
        kw = dict(sampler=lambda *a: sw.cuda_sampler(*a, device=device),
                  neighbor_builder=lambda s_, r_: sw.cuda_neighbor_builder(s_, r_, device=device))
    kw["with_neighbors"] = with_neighbors
    if name == "cfg3u":
        return sw.make_world(n=100000, m=500, seed=31, mode="uniform", layout="uniform", **kw)
    if name == "cfg4":
        return sw.make_world(4, **kw)
    return sw.make_world(SINGLE_CONFIG[name], **kw)


def cached_goal(name, it, meta):
    c = profile_json("workload_%s.json" % name)
    if c.get("n") == meta["n"] and c.get("m") == meta["m"] and c.get("seed") == meta["seed"] and \
            c.get("start") == int(it['start']) and "goal" in c:
        return int(c["goal"]), c
    return None, c


def bench_query(name, init, it, meta, device=0):
    """Benchmark query of a single-plan workload: start = state nearest (0.1 S, 0.1 S); goal = the state nearest
    (0.9 S, 0.9 S) among those the reference's GMT variant reaches from it (synthetic_world.pick_reachable_goal),
    for cfg3u the reachable state farthest from the start.  Cached in profiles/workload_<name>.json."""
    import synthetic_world as sw
    goal, c = cached_goal(name, it, meta)
    if goal is not None and (c.get("reach_fraction") is not None or meta["n"] > 200000):
        return dict(it, goal=goal), c
    reach = sw.explore_cuda(init, it, device=device)
    if goal is not None:                 # cached goal, reach fraction not yet recorded
        return dict(it, goal=goal), dict(c, reach_fraction=len(reach) / float(meta["n"]))
    st = init['states']
    if len(reach) == 0:
        return it, {}
    if name == "cfg3u":
        d = (st[reach, 0].astype(np.float64) - st[it['start'], 0]) ** 2 + (st[reach, 1].astype(np.float64) - st[it['start'], 1]) ** 2
        goal = int(reach[int(np.argmax(d))])
    else:
        goal = sw.pick_reachable_goal(init, it, meta["side"], explore=lambda a, b: reach)
    log("bench query for %s: start %d goal %d (not cached in profiles/), %d reachable states" % (name, it['start'], goal, len(reach)))
    return dict(it, goal=goal), {"reach_fraction": len(reach) / float(meta["n"])}


def save_workload(name, rec):
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):          # workload facts to cache under profiles/ (committed by hand)
        with open(os.path.join(out_dir, "workload_%s.json" % name), "w") as f:
            json.dump(rec, f)


def batch_queries(g, meta, device=0):
    """The 4096 (start, goal) pairs of cfg4 (same list on every rank and in both arms): cached in
    profiles/workload_cfg4.json (the goals need one goal-less GPU batch), else computed."""
    import synthetic_world as sw
    c = profile_json("workload_cfg4.json")
    if c.get("n") == meta["n"] and c.get("seed") == meta["seed"] and len(c.get("starts", [])) == Q_BATCH:
        return np.array(c["starts"], np.int32), np.array(c["goals"], np.int32), c
    if g is None:
        return None, None, {}
    starts, goals = sw.make_reachable_queries(g, meta["n"], Q_BATCH, meta["seed"] + 7)
    return starts, goals, {}


# --------------------------------------------------------------------------------------------------
# the reference's CPU implementation on a bounded sample (reference arm and cpu_baseline leg)
# --------------------------------------------------------------------------------------------------
def cpu_backend():
    from oracle import build_oracle
    from oracle.gmt_oracle import OraclePort, RefHost
    if RefHost.available():
        return RefHost(), "reference"
    build_oracle.build_port()
    return OraclePort("libm"), "port"


def cpu_batch_sample(init, it, backend, starts, goals, pairs_total, first, count, budget_s, threads):
    """Plans queries first .. first+count-1 of the batch with the restated GMT.run_step over the reference kernels,
    `threads` plans at a time, one host thread per plan (independent queries are the natural unit of CPU
    parallelism for a batch; the library call releases the GIL).  A plan that exceeds budget_s seconds of
    dubinConnection work is cut there and counted as the fraction of its edge checks it got through.
    Returns (plans per second, description)."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.gmt_oracle import OracleGMT

    def one(i):
        if hasattr(backend, "set_num_threads"):
            backend.set_num_threads(1)          # (OpenMP's thread-count setting is per calling thread)
        g = OracleGMT(init, backend)
        g.run_step(dict(it, start=int(starts[i]), goal=int(goals[i])), iter_limit=1000, budget_s=budget_s)
        if g.status in (0, 1):
            return 1.0, g.edge_pairs, True
        tot = pairs_total[i] if pairs_total is not None else 0
        return (min(1.0, g.edge_pairs / float(tot)) if tot else 0.0), g.edge_pairs, False

    idx = [(first + k) % len(starts) for k in range(count)]
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=threads) as ex:
        res = list(ex.map(one, idx))
    wall = time.perf_counter() - t0
    done = sum(r[0] for r in res)
    whole = sum(1 for r in res if r[2])
    pairs = sum(r[1] for r in res)
    desc = ("%d queries of the batch on %d host threads (one plan per thread, brute-force dubinConnection): %d whole "
            "plans, %d cut after %.0f s and counted by edge-check fraction; %d edge checks in %.1f s"
            % (count, threads, whole, count - whole, budget_s, pairs, wall))
    return done / wall, desc


def cpu_edge_checks(init, it, backend, pairs, threads):
    """Dubins edge checks/s of the reference's computeDubinsCost on the host: `pairs` neighbour pairs, all 4 words,
    against the obstacle set (kernel.py:418-431 per pair)."""
    rng = np.random.default_rng(7)
    nbr, num = init['neighbors'], init['num_neighbors']
    rows = np.repeat(np.arange(len(num), dtype=np.int32), num)
    sel = rng.integers(0, len(nbr), pairs)
    y, x = np.ascontiguousarray(rows[sel]), np.ascontiguousarray(nbr[sel].astype(np.int32))
    if hasattr(backend, "set_num_threads"):
        backend.set_num_threads(threads)
    t0 = time.perf_counter()
    backend.edge_costs(init['states'], y, x, it['radius'], it['obstacles'], it['num_obs'])
    return pairs / (time.perf_counter() - t0)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    cores = host_cores()
    init, it, meta = make_named_world("cfg4", cpu=True)
    backend, kind = cpu_backend()
    starts, goals, cache = batch_queries(None, meta)
    note = ""
    if starts is None:       # no cached list (it needs one GPU pass): uniform pairs at least S/2 apart instead
        import synthetic_world as sw
        starts, goals = sw.make_queries(init['states'], meta["side"], Q_BATCH, meta["seed"] + 7)
        note = " (profiles/workload_cfg4.json missing: uniform start/goal pairs)"
    pairs_total = cache.get("edge_pairs")
    steps = args.warmup + args.steps
    budget = min(args.cpu_budget, 150.0 / max(steps, 1))
    vals, walls, desc = [], [], ""
    for s in range(steps):
        t0 = time.perf_counter()
        v, desc = cpu_batch_sample(init, it, backend, starts, goals, pairs_total, s * cores, cores, budget, cores)
        if s >= args.warmup:
            vals.append(v)
            walls.append(time.perf_counter() - t0)
        log("reference step %d: %.6g plans/s (%s)" % (s, v, desc))
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": float(np.mean(walls)) * 1e3,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": headline_config(),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc + note},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
class Ctx:
    """Per-process GPU context of the own arm: device, explicit stream, library, flush buffer, barrier."""

    def __init__(self):
        import torch
        import _lib
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device -- the planner has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=self.dev)
            self.dist = dist
        self.lib = _lib.load()
        self._lib = _lib
        # an EXPLICIT stream: the library treats a NULL stream as "use the handle's own", so torch's default stream
        # (handle 0) would leave the timing events and the L2 flush on another, unordered stream
        self.stream = torch.cuda.Stream(self.dev)
        assert self.stream.cuda_stream != 0
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)        # > 126 MB L2

    def bind(self, g):
        import ctypes as C
        self._lib.check(self.lib.gmt_set_stream(g._handle, C.c_void_p(self.stream.cuda_stream)), "gmt_set_stream")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def events(self, k):
        return [(self.torch.cuda.Event(enable_timing=True), self.torch.cuda.Event(enable_timing=True)) for _ in range(k)]

    def timed(self, K, fn, pre=None):
        """K timed calls of fn() on the explicit stream, L2 flushed before each: (device seconds by CUDA events,
        host wall seconds with a synchronize on both sides of every call)."""
        torch = self.torch
        ev = self.events(K)
        wall = 0.0
        with torch.cuda.stream(self.stream):
            for k in range(K):
                self.flush.fill_(k & 0xFF)                                       # evict L2 (same stream: ordered)
                if pre:
                    pre(k)
                self.stream.synchronize()
                ev[k][0].record(self.stream)
                t0 = time.perf_counter()
                fn()
                self.stream.synchronize()
                wall += time.perf_counter() - t0
                ev[k][1].record(self.stream)
            self.stream.synchronize()
        self.last_steps_ms = [a.elapsed_time(b) for a, b in ev]
        return sum(self.last_steps_ms) * 1e-3, wall


def time_single(ctx, name, K=30, words=4, oriented=False):
    """Single-query workload: resident latency (gmt_plan_resident, CUDA events) and e2e latency (GMT.run_step with
    host obstacle buffers).  Returns the sub-record."""
    import ctypes as C
    from gmt_planner import GMT
    _lib, lib = ctx._lib, ctx.lib
    init, it, meta = make_named_world(name, device=ctx.local)
    it, cache = bench_query(name, init, it, meta, device=ctx.local)
    n, m = meta["n"], meta["m"]
    g = GMT(init, device=ctx.local, words=words)
    ctx.bind(g)
    start, goal = int(it['start']), int(it['goal'])
    radius, thr0 = float(np.float32(it['radius'])), float(np.float32(it['threshold']))
    obs = np.ascontiguousarray(it['obstacles'], np.float32)
    if oriented:
        # the same boxes as oriented rectangles turned by seeded angles (extension mode, parity unpinned)
        rng = np.random.default_rng(5)
        lo, hi = np.minimum(obs[:, :2], obs[:, 2:]), np.maximum(obs[:, :2], obs[:, 2:])
        rect = np.concatenate([(lo + hi) / 2, (hi - lo) / 2, rng.uniform(-np.pi, np.pi, (len(obs), 1))], 1).astype(np.float32)
        it = dict(it, obstacles=rect, obstacle_format='oriented')
        g.run_step(dict(it, goal=-1), iter_limit=1000)           # the turned rectangles change what is reachable:
        goal = int(g.last_result.farthest_frontier)              # plan to the farthest state the wavefront enters
        it = dict(it, goal=goal)
        g.run_step(it, iter_limit=1000)
    else:
        _lib.check(lib.gmt_set_obstacles(g._handle, obs.ctypes.data if m else None, m), "gmt_set_obstacles")
    res = _lib.GmtResult()

    def plan_resident():
        _lib.check(lib.gmt_plan_resident(g._handle, start, goal, radius, thr0, 1000, C.byref(res)), "gmt_plan_resident")

    for _ in range(3):
        plan_resident()
        g.run_step(it, iter_limit=1000)
    kern = [0.0]

    def timed_resident():
        plan_resident()
        kern[0] += res.device_ms
    t_dev, _ = ctx.timed(K, timed_resident)
    stats = dict(status=res.status, wavefronts=res.iterations, goal_cost=res.goal_cost, route_len=res.route_len,
                 edge_checks=res.edge_pairs, pairs_evaluated=res.edge_evals, words_swept=res.word_checks)
    dbg = (C.c_ulonglong * 8)()
    lib.gmt_debug_counters(g._handle, dbg)
    _, wall = ctx.timed(K, lambda: g.run_step(it, iter_limit=1000))
    assert kern[0] * 1e-3 <= t_dev * 1.0005 + 1e-6, "plan kernels (%.4f ms) cannot outlast the events around them (%.4f ms)" % (kern[0], t_dev * 1e3)
    st = init['states']
    gd = float(np.hypot(st[goal, 0] - st[start, 0], st[goal, 1] - st[start, 1])) if goal >= 0 else None
    rec = {"workload": WORKLOADS.get(name, name), "n": n, "m": m, "seed": meta["seed"], "words": words,
           "obstacles": "oriented rectangles" if oriented else "corner boxes",
           "resident_ms": t_dev / K * 1e3, "e2e_ms": wall / K * 1e3, "kernel_ms": kern[0] / K,
           "plans_per_s_resident": K / t_dev, "plans_per_s_e2e": K / wall, "steps": K, "plan": stats,
           "goal_distance_over_side": gd / meta["side"] if gd is not None else None,
           "reach_fraction": cache.get("reach_fraction"),
           "phase_us_per_wavefront": {"frontier": dbg[0] / 1e3 / max(res.iterations, 1), "select": dbg[1] / 1e3 / max(res.iterations, 1),
                                      "connect": dbg[4] / 1e3 / max(res.iterations, 1)},
           "h2d_bytes_per_step": int(16 * m), "d2h_bytes_per_step": int(512 + 4 * min(n + 1, 1024))}
    rec["algorithmic_tflops"] = res.edge_pairs * f_edge(m) / (kern[0] / K * 1e-3) * 1e-12
    rec["executed_tflops"] = f_executed(res.edge_evals, res.word_checks, words) / (kern[0] / K * 1e-3) * 1e-12
    if not oriented and words == 4:
        save_workload(name, {"n": n, "m": m, "seed": meta["seed"], "start": start, "goal": goal,
                             "edge_pairs": res.edge_pairs, "wavefronts": res.iterations, "status": res.status,
                             "goal_cost": res.goal_cost, "reach_fraction": cache.get("reach_fraction")})
    return rec, (g, init, it, meta)


def gpu_reference_leg(ctx, g, init, it, tag, budget_s=20.0):
    """The reference's own kernels (ref_kernel.cubin) on THIS GPU under the restated GMT.run_step loop with
    device-resident arrays -- a reported baseline and, because the same query was just planned by the CUDA path,
    a bit-exact comparison of the two trees (status, iterations, parents, cost bit patterns)."""
    from oracle.ref_gpu import RefGpu, RefGpuPlanner
    if not RefGpu.available():
        return {"unavailable": "oracle/_ref/ref_kernel.cubin did not travel to this box"}
    ref = RefGpuPlanner(init)
    ref.run_step(it, iter_limit=2, budget_s=budget_s)                     # warm-up (module load, allocator)
    r = ref.run_step(it, iter_limit=1000, budget_s=budget_s)
    out = {"workload": tag, "kind": "reference kernels (carla/kernel.py compiled for sm_100a) + restated host loop, "
                                    "device-resident arrays, torch.cumsum for the three scans",
           "status": r["status"], "wavefronts": r["iterations"], "seconds": r["seconds"], "edge_checks": r["edge_pairs"]}
    if r["status"] in (0, 1):
        out["plans_per_s"] = 1.0 / r["seconds"]
        out["edge_checks_per_s"] = r["edge_pairs"] / r["seconds"]
        g.run_step(it, iter_limit=1000)
        parent, cost = g.parent, g.dev_cost
        out["parity"] = {"status_equal": bool(g.status == r["status"]), "iterations_equal": bool(g.iterations == r["iterations"]),
                         "parents_differing": int((parent != r["parent"]).sum()),
                         "cost_bit_patterns_differing": int((cost.view(np.uint32) != r["cost"].view(np.uint32)).sum()),
                         "nodes_connected": int((r["parent"] >= 0).sum()), "route_equal": bool(list(g.route) == r["route"])}
    else:
        out["note"] = "cut after %.0f s" % budget_s
    return out


def time_sharded(ctx, K=8):
    K = int(os.environ.get("GMT_BENCH_CFG5_STEPS", K))
    """cfg5: ONE plan on the 1M-sample roadmap.  N = 1: the plain persistent plan.  N > 1: node-range sharding,
    fused form (one launch per GPU, per-wavefront exchange over NVLink peer memory inside the kernel) timed with
    CUDA events, max over ranks; the NCCL all-reduce(min) form once beside it; every rank's route must be the
    single-GPU route."""
    import ctypes as C
    import multi_gpu
    from gmt_planner import GMT
    torch, _lib, lib = ctx.torch, ctx._lib, ctx.lib
    rank, world = ctx.rank, ctx.world
    init, it, meta = make_named_world("cfg5", device=ctx.local, with_neighbors=False)   # lists are built on the device
    n, m = meta["n"], meta["m"]
    goal, cache = cached_goal("cfg5", it, meta)
    if goal is None:
        if rank == 0:
            init_full, _it, _meta = make_named_world("cfg5", device=ctx.local)
            it2, cache = bench_query("cfg5", init_full, it, meta, device=ctx.local)
            goal = int(it2['goal'])
            del init_full
        gt = torch.tensor([goal or 0], device=ctx.dev)
        if ctx.dist is not None:
            ctx.dist.broadcast(gt, 0)
        goal = int(gt.item())
    it = dict(it, goal=goal)
    t0 = time.perf_counter()
    g = GMT.from_states(init['states'], device=ctx.local)
    create_ms = (time.perf_counter() - t0) * 1e3
    ctx.bind(g)
    with torch.cuda.stream(ctx.stream):
        single_ms = 1e30
        for _ in range(4 if world > 1 else 1):                           # warm (N = 1 times this plan below anyway)
            single = list(g.run_step(it, iter_limit=1000))               # also the check value for the sharded forms
            if _ > 0 or world == 1:
                single_ms = min(single_ms, g.last_result.device_ms)
    res1 = g.last_result
    if world > 1:
        multi_gpu.open_peers(g, rank, world)

        def step():
            return multi_gpu.run_step_sharded_p2p(g, it, iter_limit=1000, rank=rank, world=world)
    else:
        def step():
            return g.run_step(it, iter_limit=1000), g.last_result
    out = {}
    for _ in range(3):
        out["route"], out["res"] = step()
    ctx.barrier()

    kern_steps = []

    def timed_step():
        out["route"], out["res"] = step()
        kern_steps.append(round(out["res"].device_ms, 3))
        out["all_same"] = out.get("all_same", True) and list(out["route"]) == single      # EVERY timed plan is compared
    t_dev, wall = ctx.timed(K, timed_step, pre=(lambda k: ctx.barrier()) if world > 1 else None)
    steps_ms = [round(v, 3) for v in ctx.last_steps_ms]
    kernel_ms_last = out["res"].device_ms
    if world > 1:
        allk = [None] * world
        ctx.dist.all_gather_object(allk, kern_steps)
        kern_steps = allk
    ctx.barrier()
    same = int(list(out["route"]) == single and out.get("all_same", True))
    same_fused = same
    dbg = (C.c_ulonglong * 8)()
    lib.gmt_debug_counters(g._handle, dbg)
    phases = [dbg[0] / 1e3, dbg[1] / 1e3, dbg[4] / 1e3, dbg[2] / 1e3, dbg[5] / 1e3, dbg[6] / 1e3, dbg[7] / 1e3]
    nccl_ms = 0.0
    if world > 1:
        multi_gpu.close_peers(g)
        ctx.barrier()
        with torch.cuda.stream(ctx.stream):
            t0 = time.perf_counter()
            r2, _res2, _steps2 = multi_gpu.run_step_sharded(g, it, iter_limit=1000, rank=rank, world=world)
            torch.cuda.synchronize(ctx.dev)
            nccl_ms = (time.perf_counter() - t0) * 1e3
        same = int(same and list(r2) == single)
    t_dev, wall, nccl_ms, single_ms = multi_gpu.max_over_ranks([t_dev, wall, nccl_ms, single_ms], device=ctx.dev)
    flags = torch.tensor([same, same_fused], device=ctx.dev)
    ph = torch.tensor(phases, dtype=torch.float64, device=ctx.dev)
    ph_max, ph_min = ph.clone(), ph.clone()
    if ctx.dist is not None:
        ctx.dist.all_reduce(flags, op=ctx.dist.ReduceOp.MIN)
        ctx.dist.all_reduce(ph_max, op=ctx.dist.ReduceOp.MAX)
        ctx.dist.all_reduce(ph_min, op=ctx.dist.ReduceOp.MIN)
    wf = max(res1.iterations, 1)
    rec = {"workload": WORKLOADS["cfg5"], "n": n, "m": m, "seed": meta["seed"], "n_gpus": world, "steps": K,
           "ms_per_plan": t_dev / K * 1e3, "e2e_ms_per_plan": wall / K * 1e3, "plans_per_s": K / t_dev,
           "steps_ms_rank0": steps_ms, "kernel_ms_last_plan_rank0": kernel_ms_last, "kernel_ms_steps_by_rank": kern_steps,
           "roadmap_build_ms": create_ms, "single_gpu_ms": single_ms, "speedup_vs_single_gpu": single_ms / (t_dev / K * 1e3),
           "exchange": "none (one GPU)" if world == 1 else
                       "in-kernel: one block stores this GPU's final keys into the peers' inboxes over NVLink and publishes a flag per peer (one release fence); a warp of the next frontier phase waits for the flag of the GPU that owns the node it expands, not for all peers",
           "nvlink_bytes_per_wavefront_per_gpu": None,
           "plan": {"status": res1.status, "wavefronts": res1.iterations, "edge_checks": res1.edge_pairs,
                    "route_len": len(single), "pairs_evaluated": res1.edge_evals, "words_swept": res1.word_checks},
           "identical_tree_on_every_rank": bool(flags[0].item()), "fused_form_identical": bool(flags[1].item()),
           "nccl_allreduce_min_form_ms": nccl_ms if world > 1 else None,
           "phase_us_per_wavefront_max_over_ranks": dict(zip(PHASES, (ph_max / wf).tolist())),
           "phase_us_per_wavefront_min_over_ranks": dict(zip(PHASES, (ph_min / wf).tolist())),
           "algorithmic_tflops": res1.edge_pairs * f_edge(m) / (t_dev / K) * 1e-12}
    # bytes a GPU sends per wavefront: 8 B per candidate node it connected, to each of world-1 peers
    reached = int(np.sum(np.isfinite(g.dev_cost))) if world > 1 else None
    if world > 1:
        rec["nvlink_bytes_per_wavefront_per_gpu"] = 8.0 * (world - 1) * (reached / float(world)) / wf
        rec["allreduce_bytes_per_wavefront_nccl_form"] = 8.0 * n
    if rank == 0 and world == 1:
        save_workload("cfg5", {"n": n, "m": m, "seed": meta["seed"], "start": int(it['start']), "goal": goal,
                               "edge_pairs": res1.edge_pairs, "wavefronts": res1.iterations, "status": res1.status})
    del g
    return rec


def run_b200(args):
    import ctypes as C
    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):   # NCCL printf()s its version banner to stdout at these
        os.environ.pop("NCCL_DEBUG")                                      # levels, in front of the JSON line

    import multi_gpu
    import synthetic_world as sw
    from gmt_planner import GMT

    ctx = Ctx()
    torch, _lib, lib = ctx.torch, ctx._lib, ctx.lib
    rank, world, local, dev = ctx.rank, ctx.world, ctx.local, ctx.dev
    W = max(args.warmup, 3)
    K = args.steps
    only = args.only
    sub = {}

    if only == "cfg4":
        args.no_sub = True
    if only and only != "cfg4":                                   # development: one workload, its record as the line
        if only == "cfg5":
            rec = time_sharded(ctx)
        else:
            rec, _ = time_single(ctx, only, K=max(K, 5), words=args.words, oriented=args.oriented)
        if rank == 0:
            print(json.dumps({"metric": METRIC, "only": only, "value": rec.get("plans_per_s_resident", rec.get("plans_per_s")),
                              "unit": UNIT, "n_gpus": world, "record": rec}), flush=True)
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return 0

    # ---- headline: cfg4 batch, this rank's share -------------------------------------------------
    init, it, meta = make_named_world("cfg4", device=local)
    n, m = meta["n"], meta["m"]
    g = GMT(init, device=local, words=args.words)
    ctx.bind(g)
    obs = np.ascontiguousarray(it['obstacles'], np.float32)
    _lib.check(lib.gmt_set_obstacles(g._handle, obs.ctypes.data if m else None, m), "gmt_set_obstacles")
    starts, goals, qcache = batch_queries(g, meta, device=local)              # same list on every rank
    s, gl = multi_gpu.shard_queries(starts, goals, rank, world)
    s, gl = np.ascontiguousarray(s, np.int32), np.ascontiguousarray(gl, np.int32)
    q = len(s)
    radius, thr0 = float(np.float32(it['radius'])), float(np.float32(it['threshold']))
    res = (_lib.GmtResult * q)()

    def plan_resident():
        _lib.check(lib.gmt_plan_queries(g._handle, radius, thr0, 1000, res, None, 0), "gmt_plan_queries")

    def plan_e2e():
        return sw.plan_batch(g, s, gl, route_stride=ROUTE_STRIDE)          # host arrays in, results + routes out

    sampler = ClockSampler(gpu_uuid(torch, dev, local))
    sampler.start()
    _lib.check(lib.gmt_set_queries(g._handle, s.ctypes.data, gl.ctypes.data, q), "gmt_set_queries")
    for _ in range(W):
        plan_resident()
    ctx.barrier()
    sampler.mark()
    acc = {"kern_ms": 0.0, "launches": 0}

    def timed_resident():
        plan_resident()
        acc["kern_ms"] += res[0].device_ms
        acc["launches"] += res[0].launches
    t_value, _ = ctx.timed(K, timed_resident)
    ctx.barrier()
    # self-extend (the same number of extra steps on every rank): a timed region under 0.5 s is too thin for the
    # clock sampler and for a latency-sensitive kernel
    (t_slowest,) = multi_gpu.max_over_ranks([t_value], device=dev)
    extra = 0
    if t_slowest < 0.5:
        extra = min(8 * K, K * int(np.ceil(0.5 / max(t_slowest, 1e-6))) - K)
    if extra > 0:
        t_more, _ = ctx.timed(extra, timed_resident)
        t_value += t_more
        ctx.barrier()
    K_value = K + extra
    reached = sum(r.status == 0 for r in res)
    pairs = sum(r.edge_pairs for r in res)
    evals = sum(r.edge_evals for r in res)
    checks = sum(r.word_checks for r in res)
    per_query_pairs = [int(r.edge_pairs) for r in res]
    # tail of the batch: when did each team run dry?
    be = (C.c_ulonglong * (2 * q))()
    teams = C.c_int()
    lib.gmt_debug_batch_times(g._handle, be, q, C.byref(teams))
    ends = np.array(be[1::2], np.float64)
    begins = np.array(be[0::2], np.float64)
    busy = float((ends - begins).sum())
    span = float(ends.max()) if q else 0.0
    tail = {"teams": teams.value, "queries_per_team": q / float(max(teams.value, 1)),
            "kernel_span_ms": span * 1e-6, "team_busy_fraction": busy / (span * teams.value) if span else None}

    for _ in range(2):
        plan_e2e()
    ctx.barrier()
    e2e_launch = {"n": 0}

    def timed_e2e():
        r_, _routes = plan_e2e()
        e2e_launch["n"] += r_[0].launches
    _, t_e2e = ctx.timed(K, timed_e2e)
    clocks = sampler.stop()
    ctx.barrier()
    t_value_pr, t_e2e_pr = t_value / K_value, t_e2e / K                   # per step, this rank
    t_value_pr, t_e2e_pr, kern_ms = multi_gpu.max_over_ranks([t_value_pr, t_e2e_pr, acc["kern_ms"] / K_value], device=dev)
    tot = torch.tensor([pairs, evals, checks, reached], dtype=torch.float64, device=dev)
    if ctx.dist is not None:
        ctx.dist.all_reduce(tot)
    pairs_all, evals_all, checks_all, reached_all = (float(v) for v in tot.cpu())
    assert kern_ms <= t_value_pr * 1e3 * 1.0005 + 1e-3, "kernel time %.3f ms exceeds the step time %.3f ms" % (kern_ms, t_value_pr * 1e3)

    # ---- sub-records ---------------------------------------------------------------------------
    fp32_peak = C.c_float()
    _lib.check(lib.gmt_measure_fp32_peak(local, C.byref(fp32_peak)), "gmt_measure_fp32_peak")
    edge = None
    if world == 1 and not args.no_sub:
        # edge kernel in isolation (the "Dubins edge checks/sec" half of the metric)
        rng = np.random.default_rng(7)
        P = 1 << 18
        nbr, num = init['neighbors'], init['num_neighbors']
        rows = np.repeat(np.arange(n, dtype=np.int32), num)
        sel = rng.integers(0, len(nbr), P)
        ey, ex = np.ascontiguousarray(rows[sel]), np.ascontiguousarray(nbr[sel].astype(np.int32))
        ms = C.c_float()
        best_ms = 1e30
        for _ in range(4):
            _lib.check(lib.gmt_edge_costs(g._handle, ey.ctypes.data, ex.ctypes.data, P, radius, None, None, None,
                                          C.byref(ms)), "gmt_edge_costs")
            best_ms = min(best_ms, ms.value)
        edge = {"pairs": P, "ms": best_ms, "edge_checks_per_s": P / (best_ms * 1e-3),
                "what": "all %d words + swept-path test of neighbour pairs, no pruning (gmt_edge_costs_kernel)" % args.words,
                "fp32_frac": P / (best_ms * 1e-3) * f_edge(m) * 1e-12 / fp32_peak.value,
                "ncu": profile_json("ncu_summary.json").get("gmt_edge_costs_kernel")}
        for name in ("cfg2", "cfg3", "cfg3u"):
            try:
                rec, keep = time_single(ctx, name, K=30, words=args.words, oriented=args.oriented)
                if name == "cfg2" and args.words == 4 and not args.oriented:
                    rec["gpu_reference"] = gpu_reference_leg(ctx, keep[0], keep[1], keep[2], "cfg2 query")
                sub[name] = rec
                del keep
            except Exception as e:          # a sub-record must never take the headline down
                sub[name] = {"failed": repr(e)}
        if args.words == 4 and not args.oriented:
            # extension mode (north star: six words, oriented rectangles; not in the reference -- parity unpinned,
            # DESIGN 7b) on the cfg2 map, so that the default line carries a number for it
            for tag, kw in (("cfg2_six_words", {"words": 6}), ("cfg2_oriented_rectangles", {"oriented": True})):
                try:
                    rec, keep = time_single(ctx, "cfg2", K=20, **kw)
                    rec["extension_mode"] = "parity unpinned: defined by oracle/gmt_oracle.c, checked against it in tests/test_gpu_extensions.py"
                    rec["ncu"] = profile_json("ncu_summary.json").get("gmt_plan_kernel_words6" if "words" in kw else "gmt_plan_kernel_oriented")
                    sub[tag] = rec
                    del keep
                except Exception as e:
                    sub[tag] = {"failed": repr(e)}
    if not args.no_sub:
        try:
            sub["cfg5"] = time_sharded(ctx)
        except Exception as e:
            sub["cfg5"] = {"failed": repr(e)}
            ctx.barrier()

    if rank != 0:
        if ctx.dist is not None:
            ctx.dist.destroy_process_group()
        return 0

    value = Q_BATCH / t_value_pr
    e2e_value = Q_BATCH / t_e2e_pr
    kern_s = kern_ms * 1e-3
    achieved = pairs_all / world * f_edge(m) / kern_s * 1e-12              # per GPU: its share of the work / its kernel time
    executed = f_executed(evals_all / world, checks_all / world, args.words) / kern_s * 1e-12
    ncu = profile_json("ncu_summary.json")
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": t_value_pr * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic", "config": headline_config(),
        "timing": {"timed_steps_value": K_value, "l2": "flushed before every timed step (256 MiB fill on the timing stream)",
                   "stream": "explicit torch stream handed to the library (gmt_set_stream)",
                   "queries_per_rank": q, "kernel_ms_per_step": kern_ms,
                   "mean_degree": round(meta["mean_degree"], 2), "words": args.words},
        "batch": {"goal_reached": int(reached_all), "edge_checks": int(pairs_all), "pairs_evaluated": int(evals_all),
                  "words_swept": int(checks_all), "tail_rank0": tail},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(8 * q),
                "d2h_bytes_per_step": int(q * (C.sizeof(_lib.GmtResult) + 4 * ROUTE_STRIDE)), "ms_per_step": t_e2e_pr * 1e3,
                "note": "gmt_plan_batch: host start/goal arrays in, host result structs + routes out, wall clock"},
        "gpu_launches": acc["launches"] + e2e_launch["n"],
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"]},
        "roofline": {"bound": "fp32", "kernel": "gmt_plan_kernel<%d,false,1024> (batch: one team of 32 warps per SM)" % args.words, "achieved": achieved,
                     "peak": fp32_peak.value, "unit": "TFLOP/s", "frac": achieved / fp32_peak.value if fp32_peak.value else None,
                     "executed": executed, "executed_frac": executed / fp32_peak.value if fp32_peak.value else None,
                     "traffic": (ncu.get("gmt_plan_kernel_batch") or {}).get("dram_bytes_per_launch"),
                     "kernel_ms": kern_ms, "ncu": ncu.get("gmt_plan_kernel_batch"),
                     "peak_source": "FP32 FMA micro-benchmark measured live on this GPU (MEASURED_PEAKS.json has no FP32 figure)",
                     "note": "achieved = reference edge checks (sum |X|*|Y| over the batch = %d per GPU) x F_edge(m=%d) = %.0f "
                             "op/check / kernel time: > 1 because exact pruning evaluates %d of those pairs and sweeps %d "
                             "words; executed = op-equivalents of the work actually done; ncu = pipe utilisation of the "
                             "committed capture of this launch" % (pairs_all / world, m, f_edge(m), evals_all / world, checks_all / world)},
        "edge_kernel": edge,
        "sub": sub,
    }
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = None
    traffic = line["roofline"]["traffic"]
    line["roofline"]["hbm"] = {"achieved": (traffic / kern_s * 1e-9) if traffic else None, "peak": hbm_peak, "unit": "GB/s",
                               "frac": (traffic / kern_s * 1e-9 / hbm_peak) if traffic and hbm_peak else None,
                               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "absent"}
    if world == 1:
        save_workload("cfg4", {"n": n, "m": m, "seed": meta["seed"], "starts": [int(v) for v in starts],
                               "goals": [int(v) for v in goals], "edge_pairs": per_query_pairs})

    # ---- cpu baseline (rank 0, N = 1 only) -------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            cores = host_cores()
            backend, kind = cpu_backend()
            init_c, it_c, _ = make_named_world("cfg4", cpu=True)
            v, desc = cpu_batch_sample(init_c, it_c, backend, starts, goals, per_query_pairs, 0, cores, args.cpu_budget, cores)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
            line["cpu_edge_checks"] = {
                "all_cores": {"edge_checks_per_s": cpu_edge_checks(init_c, it_c, backend, 100000, cores), "pairs": 100000, "cores": cores},
                "one_core": {"edge_checks_per_s": cpu_edge_checks(init_c, it_c, backend, 20000, 1), "pairs": 20000, "cores": 1},
                "what": "computeDubinsCost of the reference (kernel.py:418-431) on neighbour pairs of the cfg4 map, m = %d" % m}
        except Exception as e:  # the baseline must never take the measurement down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (e,)}
    print(json.dumps(line), flush=True)
    if ctx.dist is not None:
        ctx.dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--only", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--words", type=int, default=4, choices=[4, 6], help="6: also RLR / LRL (extension mode, parity unpinned)")
    ap.add_argument("--oriented", action="store_true", help="single-query sub-records on oriented rectangles (extension mode)")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU dubinConnection work per sampled plan")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sub", action="store_true", help="headline only")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
