#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked text files under profiles/.
usage: python profiles/summarize_ncu.py r01"""
import csv
import collections
import glob
import os
import subprocess
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
here = os.path.dirname(os.path.abspath(__file__))
root = os.path.dirname(here)
out = []

# ---- launch list: share of device time per kernel (cold-cache, serialised: compare SHARES)
path = os.path.join(root, "gpurun_out", tag + "_launches.csv")
if os.path.isfile(path):
    rows = []
    with open(path) as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            v = float(r["Metric Value"].replace(",", ""))
            if r["Metric Unit"] in ("ns", "nsecond"):
                v /= 1e3
            elif r["Metric Unit"] in ("ms", "msecond"):
                v *= 1e3
            rows.append((r["Kernel Name"], v))
    # keep the steady-state tail: the last 40 % of launches (warm-up steps come first)
    tail = rows[int(len(rows) * 0.6):]
    agg = collections.OrderedDict()
    for k, v in tail:
        short = k.split("(")[0].replace("void ", "").replace("<unnamed>::", "")[:70]
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += v
    total = sum(a[1] for a in agg.values())
    out.append("## launch list (%s): %d launches, steady-state tail of %d, total %.1f us" % (tag, len(rows), len(tail), total))
    out.append("%-72s %6s %10s %7s" % ("kernel", "count", "us", "share"))
    for k, (c, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
        out.append("%-72s %6d %10.1f %6.1f%%" % (k, c, us, 100 * us / total))

# ---- full captures: the metrics the roofline needs
METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
           "sm__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
           "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct",
           "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", tag + "_k_*.ncu-rep"))):
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=300).stdout
    except Exception as e:
        out.append("## %s: ncu import failed: %s" % (os.path.basename(rep), e))
        continue
    rdr = list(csv.reader(txt.splitlines()))
    if len(rdr) < 3:
        continue
    hdr, units, vals = rdr[0], rdr[1], rdr[2]
    out.append("")
    out.append("## %s  (kernel: %s)" % (os.path.basename(rep), vals[hdr.index("Kernel Name")][:80] if "Kernel Name" in hdr else "?"))
    for m in METRICS:
        hits = [i for i, h in enumerate(hdr) if h == m]
        for i in hits:
            out.append("  %-75s %s %s" % (m, vals[i], units[i]))

# ---- dram traffic per launch of each captured kernel -> profiles/ncu_traffic.json (bench.py's roofline.traffic)
STAGE_OF = {"k_field_backward": "field_backward", "k_field_forward": "field_forward", "k_grid_bwd_d3c2": "grid_encode_backward",
            "k_grid_fwd_d3c2": "grid_encode_forward", "k_march_count_seg": "march_count", "k_march_count": "march_count",
            "k_march_expand": "march_write", "k_fused_adam": "adam", "k_composite_train_fwd": "composite_forward",
            "k_composite_train_bwd": "composite_backward"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
traffic = {}
for rep in sorted(glob.glob(os.path.join(root, "gpurun_out", tag + "_k_*.ncu-rep"))):
    kname = os.path.basename(rep)[len(tag) + 1:-len(".ncu-rep")]
    try:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, timeout=300).stdout
        rdr = list(csv.reader(txt.splitlines()))
        hdr, units, vals = rdr[0], rdr[1], rdr[2]
        tot = 0.0
        for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            i = hdr.index(m)
            tot += float(vals[i].replace(",", "")) * UNIT.get(units[i], 1.0)
        if kname in STAGE_OF:
            traffic[STAGE_OF[kname]] = tot
    except Exception as e:
        out.append("## traffic of %s unavailable: %s" % (kname, e))
if traffic:
    import json
    tp = os.path.join(here, "ncu_traffic.json")
    old = {}
    if os.path.isfile(tp):
        old = json.load(open(tp))
    old.update(traffic)
    old["_source"] = "dram__bytes_read.sum + dram__bytes_write.sum of one launch, ncu --set full captures tagged " + tag
    json.dump(old, open(tp, "w"), indent=1, sort_keys=True)

dst = os.path.join(here, tag + "_ncu_summary.txt")
open(dst, "w").write("\n".join(out) + "\n")
print("\n".join(out))
