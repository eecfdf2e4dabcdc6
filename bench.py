#!/usr/bin/env python
"""bench.py -- GMT* plans/sec on B200 (BASELINE.json metric), one JSON line on stdout.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C] [--impl b200|reference]

A "step" is ONE plan of the hot path (GMT.run_step: wavefront expansion + Dubins edge costs +
swept-path collision) on the BASELINE.json configuration `--config` (default 2: single query,
10k samples, 100 obstacles, the first single-GPU configuration the metric is quoted on).

* value        plans/s with the roadmap AND the obstacle set resident in HBM (gmt_plan_resident),
               CUDA events on the launching stream, L2 flushed between timed plans.
* e2e          the same metric through the reference-facing call GMT.run_step(iter_parameters)
               with HOST numpy buffers: obstacle upload (pinned staging -> H2D) and result/route
               download (D2H) inside the timed region.
* roofline     dominant kernel gmt_plan_kernel: algorithmic FP32 work (reference edge checks
               sum|X|*|Y| x F_edge(m)) / its CUDA-event duration, against the FP32 FMA
               peak measured live on this GPU (gmt_measure_fp32_peak).
* cpu_baseline the reference's own kernels compiled for the host (oracle/_ref, kind "reference") or
               the C restatement (kind "port") on this box's cores, on a bounded sample of the plan.
N > 1 (torchrun): independent queries, one per rank and step, no collective on the data path
("weak" scaling); NCCL is used only for the barrier and the max-over-ranks of the timings.
--impl reference times the reference CPU implementation instead (rank 0 only).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "robot-parallel-motion-planning_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

METRIC = "gmt_plans_per_sec"
UNIT = "plans/s"
WORKLOADS = {4: "cfg4: batch of 4096 independent start/goal queries on a 10k-sample map, sharded by rank, no collective",
             1: "cfg1: single query, 2k samples, 20 obstacles",
             2: "cfg2: single query, 10k samples, 100 obstacles",
             3: "cfg3: single query, 100k samples, road lattice, 2k obstacles",
             5: "cfg5: single query on ONE 1M-sample roadmap (2k obstacles), edge evaluation sharded by node range "
                "over the GPUs, per-wavefront exchange of the packed cost/parent keys"}


def f_edge(m):
    """Algorithmic FP32 op-equivalents per reference edge check: 4*(1142 + 604 m)."""
    return 4.0 * (1142.0 + 604.0 * m)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


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


# --------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's CPU implementation on a bounded sample
# --------------------------------------------------------------------------------------------------
def explore_oracle(init, it):
    """CPU twin of synthetic_world.explore_cuda: ids that ever enter a wavefront (oracle, no goal)."""
    from oracle import build_oracle
    from oracle.gmt_oracle import OraclePort
    build_oracle.build_port()
    r = OraclePort("libm").plan(init['states'], init['neighbors'], init['num_neighbors'], it['start'], -1,
                                it['radius'], it['threshold'], it['obstacles'], it['num_obs'], iter_limit=1000)
    members = [e["G"] for e in r["trace"] if e["gSize"] > 0 and "G" in e]
    return np.unique(np.concatenate(members)) if members else np.zeros(0, np.int32)


def bench_query(config, init, it, meta, explore):
    """The benchmark query of a configuration: start = state nearest (0.1S, 0.1S), goal = the state
    nearest (0.9S, 0.9S) that the reference's GMT variant can reach from it (see
    synthetic_world.pick_reachable_goal).  Cached in profiles/workload_cfgN.json once measured."""
    import synthetic_world as sw
    cached = workload_meta(config)
    if cached.get("n") == meta["n"] and cached.get("m") == meta["m"] and cached.get("seed") == meta["seed"] \
            and cached.get("start") == int(it['start']):
        return dict(it, goal=int(cached["goal"]))
    goal = sw.pick_reachable_goal(init, it, meta["side"], explore=explore)
    log("bench query for cfg%d: start %d goal %d (not cached in profiles/)" % (config, it['start'], goal))
    return dict(it, goal=goal)


def cpu_world(config):
    """The same world as the GPU arm, built by the NumPy restatements (bit-identical, tested)."""
    import synthetic_world as sw
    from oracle.world_oracle import build_neighbors, sample_states
    init, it, meta = sw.make_world(config, sampler=sample_states, neighbor_builder=build_neighbors)
    return init, bench_query(config, init, it, meta, explore_oracle), meta


def cpu_backend():
    from oracle import build_oracle
    from oracle.gmt_oracle import OraclePort, RefHost
    if RefHost.available():
        ref = RefHost()
        return ref, "reference", ref.num_threads()
    build_oracle.build_port()
    return OraclePort("libm"), "port", os.cpu_count() or 1


def cpu_sample(init, it, backend, budget_s, total_pairs):
    """Runs the restated GMT.run_step over the reference kernels for ~budget_s seconds of
    dubinConnection work.  Returns (plans_per_s, description)."""
    from oracle.gmt_oracle import OracleGMT
    g = OracleGMT(init, backend)
    t0 = time.perf_counter()
    g.run_step(it, iter_limit=1000, budget_s=budget_s)
    wall = time.perf_counter() - t0
    if g.status in (0, 1):
        return 1.0 / wall, "whole plan (%d wavefronts, %d edge checks) in %.1f s" % (g.iterations, g.edge_pairs, wall)
    rate = g.edge_pairs / max(wall, 1e-9)
    if not total_pairs:
        return float("nan"), "first %d wavefronts only; total edge checks unknown" % g.iterations
    return rate / total_pairs, ("first %d wavefronts = %d of %d edge checks (brute-force dubinConnection) in %.1f s; "
                                "plan time extrapolated by edge-check count" % (g.iterations, g.edge_pairs,
                                                                                total_pairs, wall))


def workload_meta(config):
    path = os.path.join(ROOT, "profiles", "workload_cfg%d.json" % config)
    if os.path.exists(path):
        return json.load(open(path))
    return {}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    init, it, meta = cpu_world(args.config)
    backend, kind, cores = cpu_backend()
    total = workload_meta(args.config).get("edge_pairs", 0)
    vals, desc = [], ""
    # every step is a bounded sample; the whole run stays within a few minutes whatever K is
    budget = min(args.cpu_budget, 150.0 / max(args.warmup + args.steps, 1))
    for s in range(args.warmup + args.steps):
        v, desc = cpu_sample(init, it, backend, budget, total)
        if s >= args.warmup:
            vals.append(v)
        log("reference step %d: %.6g plans/s (%s)" % (s, v, desc))
    value = float(np.mean(vals))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 / value if value > 0 else None,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOADS[args.config], "n": meta["n"], "m": meta["m"], "seed": meta["seed"]},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# --------------------------------------------------------------------------------------------------
# own arm
# --------------------------------------------------------------------------------------------------
def run_b200(args):
    import ctypes as C

    import torch
    import _lib
    import synthetic_world as sw
    from gmt_planner import GMT

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the planner has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)

    init, it, meta = sw.make_world(args.config)           # CUDA sampler + CUDA neighbour search
    it = bench_query(args.config, init, it, meta, None)
    n, m = meta["n"], meta["m"]
    g = GMT(init, device=local)
    stream = torch.cuda.current_stream(dev)
    _lib.check(lib.gmt_set_stream(g._handle, C.c_void_p(stream.cuda_stream)), "gmt_set_stream")
    # weak scaling: every rank plans the configuration's query on its own replica of the roadmap --
    # independent plans, identical work per GPU, nothing exchanged on the data path
    start, goal = int(it['start']), int(it['goal'])
    obs = np.ascontiguousarray(it['obstacles'], np.float32)
    radius, thr0 = float(np.float32(it['radius'])), float(np.float32(it['threshold']))
    _lib.check(lib.gmt_set_obstacles(g._handle, obs.ctypes.data if m else None, m), "gmt_set_obstacles")

    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)        # > 126 MB L2
    res = _lib.GmtResult()

    def plan_resident():
        _lib.check(lib.gmt_plan_resident(g._handle, start, goal, radius, thr0, 1000, C.byref(res)), "gmt_plan_resident")

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(gpu_uuid(torch, dev, local))
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        plan_resident()
        g.run_step(it, iter_limit=1000)
    sampler.mark()

    # ---- value: resident inputs ---------------------------------------------------------------
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    kern_ms, launches, pairs, evals, checks = 0.0, 0, 0, 0, 0
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)                                            # evict L2 between timed plans
        ev[k][0].record(stream)
        plan_resident()
        ev[k][1].record(stream)
        kern_ms += res.device_ms
        launches += res.launches
        pairs, evals, checks = res.edge_pairs, res.edge_evals, res.word_checks
    barrier()
    t_value = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    status, iterations, goal_cost = res.status, res.iterations, res.goal_cost

    # ---- e2e: host buffers through GMT.run_step -------------------------------------------------
    ev2 = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    wall = 0.0
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)
        torch.cuda.synchronize(dev)
        ev2[k][0].record(stream)
        t0 = time.perf_counter()
        route = g.run_step(it, iter_limit=1000)
        wall += time.perf_counter() - t0
        ev2[k][1].record(stream)
        launches += g.last_result.launches
    barrier()
    t_e2e = max(sum(a.elapsed_time(b) for a, b in ev2) * 1e-3, wall)
    clocks = sampler.stop()

    import multi_gpu
    t_value, t_e2e = multi_gpu.max_over_ranks([t_value, t_e2e], device=dev)     # slowest rank

    # ---- edge kernel in isolation (the "Dubins edge checks/sec" half of the metric) -------------
    edge = None
    fp32_peak = C.c_float()
    if rank == 0:
        _lib.check(lib.gmt_measure_fp32_peak(local, C.byref(fp32_peak)), "gmt_measure_fp32_peak")
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
                "what": "all 4 words + swept-path test of neighbour pairs, no pruning"}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    value = world * K / t_value
    e2e_value = world * K / t_e2e
    kern_s = kern_ms * 1e-3 / K
    achieved = pairs * f_edge(m) / kern_s * 1e-12
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": max(args.warmup, 3),
        "ms_per_step": t_value / K * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.config], "n": n, "m": m, "seed": meta["seed"],
                   "mean_degree": round(meta["mean_degree"], 2), "queries_per_step_per_gpu": 1,
                   "multi_gpu": "each rank plans this query independently on its own replica (no collective)",
                   "l2": "flushed between timed plans (256 MiB fill)", "iter_limit": 1000,
                   "plan": {"status": status, "wavefronts": iterations, "goal_cost": goal_cost,
                            "route_len": len(route), "edge_checks": pairs, "pairs_evaluated": evals,
                            "words_swept": checks}},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(16 * m),
                "d2h_bytes_per_step": int(512 + 4 * min(n + 1, 1024)), "ms_per_step": t_e2e / K * 1e3},
        "gpu_launches": launches,
        "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                   "samples": clocks["samples"]},
        "roofline": {"bound": "fp32", "kernel": "gmt_plan_kernel", "achieved": achieved, "peak": fp32_peak.value,
                     "unit": "TFLOP/s", "frac": achieved / fp32_peak.value if fp32_peak.value else None,
                     "traffic": workload_meta(args.config).get("ncu_dram_bytes_per_launch"), "kernel_ms": kern_s * 1e3,
                     "note": "achieved = reference edge checks (sum |X|*|Y| = %d) x F_edge(m=%d) = %.0f op/check "
                             "/ kernel time; peak = FP32 FMA micro-benchmark measured live on this GPU "
                             "(MEASURED_PEAKS.json has no FP32 figure). >1 is expected: exact pruning evaluates "
                             "%d of those pairs and sweeps %d words" % (pairs, m, f_edge(m), evals, checks)},
        "edge_kernel": edge,
    }
    if edge:
        line["edge_kernel"]["fp32_frac"] = edge["edge_checks_per_s"] * f_edge(m) * 1e-12 / fp32_peak.value
    # the HBM view of the same kernel, for completeness: measured DRAM bytes per launch (ncu) over its duration
    # against the measured copy bandwidth of MEASURED_PEAKS.json -- it shows how far from memory-bound the path is
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        hbm_peak = None
    traffic = line["roofline"]["traffic"]
    line["roofline"]["hbm"] = {"achieved": (traffic / kern_s * 1e-9) if traffic else None, "peak": hbm_peak, "unit": "GB/s",
                               "frac": (traffic / kern_s * 1e-9 / hbm_peak) if traffic and hbm_peak else None,
                               "peak_source": "MEASURED_PEAKS.json hbm_gbs" if hbm_peak else "absent"}
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):          # workload facts to cache under profiles/ (committed by hand)
        with open(os.path.join(out_dir, "workload_cfg%d.json" % args.config), "w") as f:
            json.dump({"n": n, "m": m, "seed": meta["seed"], "start": start, "goal": goal, "edge_pairs": pairs,
                       "wavefronts": iterations, "status": status, "goal_cost": goal_cost}, f)

    # ---- cpu baseline (rank 0, N = 1 only) -------------------------------------------------------
    if world == 1 and not args.no_cpu_baseline:
        try:
            backend, kind, cores = cpu_backend()
            init_c, it_c, _ = cpu_world(args.config)
            v, desc = cpu_sample(init_c, it_c, backend, args.cpu_budget, pairs)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc}
        except Exception as e:  # the baseline must never take the measurement down
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %r" % (e,)}
    print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_b200_batch(args):
    """BASELINE config 4: 4096 independent queries on one 10k map; each rank plans its contiguous share
    with gmt_plan_batch (one launch, teams of blocks); strong scaling, no collective on the data path."""
    import ctypes as C

    import torch
    import _lib
    import multi_gpu
    import synthetic_world as sw
    from gmt_planner import GMT

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the planner has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)
    Q = 4096
    init, it, meta = sw.make_world(4)
    n, m = meta["n"], meta["m"]
    g = GMT(init, device=local)
    stream = torch.cuda.current_stream(dev)
    _lib.check(lib.gmt_set_stream(g._handle, C.c_void_p(stream.cuda_stream)), "gmt_set_stream")
    obs = np.ascontiguousarray(it['obstacles'], np.float32)
    _lib.check(lib.gmt_set_obstacles(g._handle, obs.ctypes.data if m else None, m), "gmt_set_obstacles")
    starts, goals = sw.make_reachable_queries(g, n, Q, meta["seed"] + 7)       # same list on every rank
    s, gl = multi_gpu.shard_queries(starts, goals, rank, world)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(gpu_uuid(torch, dev, local))
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        res, _r = sw.plan_batch(g, s, gl)
    sampler.mark()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches, kern_ms = 0, 0.0
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        res, routes = sw.plan_batch(g, s, gl, route_stride=128)      # host buffers in, results + routes out
        ev[k][1].record(stream)
        launches += res[0].launches
        kern_ms += res[0].device_ms
    barrier()
    t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    clocks = sampler.stop()
    (t,) = multi_gpu.max_over_ranks([t], device=dev)
    reached = sum(r.status == 0 for r in res)
    pairs = sum(r.edge_pairs for r in res)
    if rank == 0:
        value = Q * K / t
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(args.warmup, 3), "ms_per_step": t / K * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[4], "n": n, "m": m, "seed": meta["seed"], "queries": Q,
                           "queries_per_rank": len(s), "l2": "flushed between timed batches (256 MiB fill)",
                           "goal_reached_rank0": reached, "edge_checks_rank0": pairs},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(8 * len(s)),
                        "d2h_bytes_per_step": int(len(s) * (144 + 4 * 128)),
                        "note": "the batch call takes host start/goal arrays and returns host results + routes: value is e2e"},
                "gpu_launches": launches,
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                           "samples": clocks["samples"]},
                "roofline": {"bound": "fp32", "kernel": "gmt_plan_kernel", "unit": "TFLOP/s", "traffic": None,
                             "achieved": pairs * f_edge(m) / (kern_ms / K * 1e-3) * 1e-12, "peak": None, "frac": None,
                             "kernel_ms": kern_ms / K}}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def run_b200_sharded(args):
    """BASELINE config 5: ONE plan on a 1M-sample roadmap, every GPU connects the candidate nodes of its node
    range.  N = 1: the plain persistent plan.  N > 1: the fused form (one launch per GPU, the per-wavefront
    exchange goes over NVLink peer memory inside the kernel); the NCCL all-reduce(min) form is timed beside it.
    Strong scaling: value = plans/s of the one plan all GPUs work on."""
    import ctypes as C

    import torch
    import _lib
    import multi_gpu
    import synthetic_world as sw
    from gmt_planner import GMT

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the planner has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = _lib.load()
    dev = torch.device("cuda", local)
    init, it, meta = sw.make_world(5, sampler=lambda *a: sw.cuda_sampler(*a, device=local),
                                   neighbor_builder=lambda s_, r_: sw.cuda_neighbor_builder(s_, r_, device=local))
    n, m = meta["n"], meta["m"]
    cached = workload_meta(5)
    if cached.get("n") == n and cached.get("seed") == meta["seed"] and cached.get("start") == int(it['start']):
        goal = int(cached["goal"])
    else:
        goal = sw.pick_reachable_goal(init, it, meta["side"], explore=lambda a, b: sw.explore_cuda(a, b, device=local)) \
            if rank == 0 else 0
        gt = torch.tensor([goal], device=dev)
        if dist is not None:
            dist.broadcast(gt, 0)
        goal = int(gt.item())
    it = dict(it, goal=goal)
    g = GMT(init, device=local)
    stream = torch.cuda.current_stream(dev)
    _lib.check(lib.gmt_set_stream(g._handle, C.c_void_p(stream.cuda_stream)), "gmt_set_stream")
    single = list(g.run_step(it, iter_limit=1000))                   # also the check value for the sharded forms
    res1 = g.last_result
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    if world > 1:
        multi_gpu.open_peers(g, rank, world)

        def step():
            return multi_gpu.run_step_sharded_p2p(g, it, iter_limit=1000, rank=rank, world=world)
    else:
        def step():
            return g.run_step(it, iter_limit=1000), g.last_result
    sampler = ClockSampler(gpu_uuid(torch, dev, local))
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        route, res = step()
    sampler.mark()
    K = args.steps
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    launches = 0
    barrier()
    for k in range(K):
        flush.fill_(k & 0xFF)
        ev[k][0].record(stream)
        route, res = step()
        ev[k][1].record(stream)
        launches += res.launches
    barrier()
    t = sum(a.elapsed_time(b) for a, b in ev) * 1e-3
    clocks = sampler.stop()
    same = int(list(route) == single)
    nccl_ms = None
    if world > 1:
        multi_gpu.close_peers(g)
        barrier()
        t0 = time.perf_counter()
        r2, _res2, steps2 = multi_gpu.run_step_sharded(g, it, iter_limit=1000, rank=rank, world=world)
        torch.cuda.synchronize(dev)
        nccl_ms = (time.perf_counter() - t0) * 1e3
        same = int(same and list(r2) == single)
    t, nccl_ms_max = multi_gpu.max_over_ranks([t, nccl_ms or 0.0], device=dev)
    flags = torch.tensor([same], device=dev)
    if dist is not None:
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0:
        value = K / t
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K,
                "warmup": max(args.warmup, 3), "ms_per_step": t / K * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": WORKLOADS[5], "n": n, "m": m, "seed": meta["seed"],
                           "l2": "flushed between timed plans (256 MiB fill)",
                           "exchange": "none (one GPU)" if world == 1 else
                                       "in-kernel stores into the peers' arrays over NVLink + flag barrier in peer memory",
                           "plan": {"status": res1.status, "wavefronts": res1.iterations, "edge_checks": res1.edge_pairs,
                                    "route_len": len(single), "single_gpu_ms": res1.device_ms},
                           "identical_tree_on_every_rank": bool(flags.item()),
                           "nccl_allreduce_min_path_ms": nccl_ms_max if world > 1 else None},
                "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": int(16 * m),
                        "d2h_bytes_per_step": int(512 + 4 * min(n + 1, 1024)),
                        "note": "timed through GMT.run_step / multi_gpu.run_step_sharded_p2p with host obstacle arrays"},
                "gpu_launches": launches,
                "clocks": {"sm_mhz": clocks["sm_mhz"], "sm_max_mhz": clocks["sm_max_mhz"], "reasons": clocks["reasons"],
                           "samples": clocks["samples"]},
                "roofline": {"bound": "fp32", "kernel": "gmt_plan_kernel", "unit": "TFLOP/s", "traffic": None,
                             "achieved": res1.edge_pairs * f_edge(m) / (t / K) * 1e-12, "peak": None, "frac": None,
                             "note": "algorithmic op count of the reference formulation (SURVEY 8d); exact pruning "
                                     "skips almost all of it, see profiles/ for the issue-slot utilisation"}}
        print(json.dumps(line), flush=True)
        out_dir = os.path.join(ROOT, "gpurun_out")
        if os.path.isdir(out_dir):
            with open(os.path.join(out_dir, "workload_cfg5.json"), "w") as f:
                json.dump({"n": n, "m": m, "seed": meta["seed"], "start": int(it['start']), "goal": goal,
                           "edge_pairs": res1.edge_pairs, "wavefronts": res1.iterations, "status": res1.status}, f)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--config", type=int, default=2, choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU dubinConnection work per sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        if args.config in (4, 5):
            args.config = 2          # the CPU arm times one query of the same kind of map
        return run_reference(args)
    if args.config == 4:
        if args.steps == 200:
            args.steps = 5
        return run_b200_batch(args)
    if args.config == 5:
        if args.steps == 200:
            args.steps = 10
        return run_b200_sharded(args)
    return run_b200(args)


if __name__ == "__main__":
    sys.exit(main())
