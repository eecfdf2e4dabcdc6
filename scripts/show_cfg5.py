"""Development aid: the phase split of a `bench.py --only cfg5` line."""
import json, sys
d = json.loads([l for l in open(sys.argv[1]) if l.startswith("{")][-1])
d = d.get("record", d.get("sub", {}).get("cfg5", d))
print("ms/plan %.2f  single-GPU %.2f  speed-up %.2f  identical %s  NCCL form %s ms" % (d["ms_per_plan"], d["single_gpu_ms"], d["speedup_vs_single_gpu"], d["identical_tree_on_every_rank"], d["nccl_allreduce_min_form_ms"]))
print("max over ranks:", {k: round(v, 1) for k, v in d["phase_us_per_wavefront_max_over_ranks"].items()})
print("min over ranks:", {k: round(v, 1) for k, v in d["phase_us_per_wavefront_min_over_ranks"].items()})
