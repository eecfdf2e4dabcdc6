"""Development aid: the handful of ncu metrics bench.py quotes (profiles/ncu_summary.json) from .ncu-rep captures.
usage: ncu_summary.py name=report.ncu-rep[:launch] ...   (prints JSON; run where ncu is installed, no GPU needed)"""
import csv, io, json, subprocess, sys
KEYS = {"gpu__time_duration.sum": "duration", "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_slots_busy_pct",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active": "fma_pipe_pct",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active": "alu_pipe_pct",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active": "xu_pipe_pct",
        "sm__warps_active.avg.pct_of_peak_sustained_active": "achieved_occupancy_pct",
        "launch__registers_per_thread": "registers_per_thread", "launch__block_size": "block_size", "launch__grid_size": "grid_size",
        "smsp__inst_executed.sum": "warp_instructions", "dram__bytes_read.sum": "dram_read", "dram__bytes_write.sum": "dram_write",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio": "stall_barrier_per_issue",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard_per_issue",
        "sm__inst_executed.avg.per_cycle_elapsed": "ipc_elapsed"}
UNIT = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0, "us": 1e-3, "ms": 1.0, "s": 1e3, "ns": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3, "nsecond": 1e-6}
out = {}
for arg in sys.argv[1:]:
    name, rep = arg.split("=")
    idx = 0
    if rep.rsplit(":", 1)[-1].isdigit():                     # report.ncu-rep:K = the K-th profiled launch of the report
        rep, k = rep.rsplit(":", 1)
        idx = int(k)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + idx]
    rec = {"kernel": vals[hdr.index("Kernel Name")], "report": rep.split("/")[-1]}
    for k, nm in KEYS.items():
        if k in hdr:
            i = hdr.index(k)
            v = float(vals[i].replace(",", ""))
            u = units[i]
            if nm == "duration":
                rec["duration_ms"] = v * UNIT.get(u, 1.0)
            elif nm in ("dram_read", "dram_write"):
                rec[nm + "_bytes"] = v * UNIT.get(u, 1.0)
            else:
                rec[nm] = v
    rec["dram_bytes_per_launch"] = rec.get("dram_read_bytes", 0.0) + rec.get("dram_write_bytes", 0.0)
    out[name] = rec
print(json.dumps(out, indent=1))
