"""Development aid: per-source-line instruction / stall-sample shares of a kernel from an ncu report.
usage: ncu_lines.py <report.ncu-rep> <kernel substring> [top]   (needs the in-tree libgmt_b200.so built with -lineinfo)"""
import collections, csv, os, re, subprocess, sys, tempfile
rep, kname = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "robot-parallel-motion-planning_b200", "libgmt_b200.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=tmp, capture_output=True)
dis = ""
for f in sorted(os.listdir(tmp)):
    if f.endswith(".cubin") and "-" not in f.split(".")[0]:
        d = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if kname in d:
            dis = d
fn = cur_file = line = None
ins = []
for l in dis.split("\n"):
    m = re.match(r"\s*\.text\.(\S+):", l)
    if m:
        fn = m.group(1); continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur_file, line = m.group(1).split("/")[-1], int(m.group(2)); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and fn and kname in fn:
        ins.append((int(m.group(1), 16), cur_file, line))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.split("\n")))
hdr, data = rows[1], [r for r in rows[2:] if len(r) > 5]
ia, sa = hdr.index("Instructions Executed"), hdr.index("# Samples")
base = int(data[0][0], 16)
off = {o: (f, ln) for o, f, ln in ins}
agg, smp = collections.Counter(), collections.Counter()
tot = ts = 0
for r in data:
    o = int(r[0], 16) - base
    ie, s = int(r[ia]), int(r[sa]); tot += ie; ts += s
    key = off.get(o, ("?", 0)); agg[key] += ie; smp[key] += s
srcs = {}
print("kernel %s: %d SASS instructions, %.3f G warp instructions executed, %d stall samples" % (kname, len(data), tot / 1e9, ts))
order = smp.most_common(top) if os.environ.get("BY_SAMPLES") else agg.most_common(top)
for (f, ln), _v in order:
    ie = agg[(f, ln)]
    if f not in srcs:
        pth = os.path.join(root, "robot-parallel-motion-planning_b200", "csrc", f or "")
        srcs[f] = open(pth).read().split("\n") if os.path.isfile(pth) else []
    txt = srcs[f][ln - 1].strip()[:88] if 0 < ln <= len(srcs[f]) else ""
    print("%5.1f%% inst %5.1f%% smp  %s:%s  %s" % (100 * ie / tot, 100 * smp[(f, ln)] / max(ts, 1), f, ln, txt))
