#!/usr/bin/env python
"""Turns the scratch output of a gpurun profiling pass (gpurun_out/, written by scripts/profile.sh) into
the tracked summaries under profiles/: the ncu launch list (per-kernel shares), the `--set full` metrics of
the top kernels, and profiles/solve_traffic.json / spmv_traffic.json (DRAM bytes per launch of the solve
kernel / of an SpMV launch of the multi-kernel solver: bench.py's static fallback for roofline.traffic).

    python scripts/summarize_profiles.py <tag> [workload]"""
import collections
import csv
import json
import os
import re
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
workload = sys.argv[2] if len(sys.argv) > 2 else "film20m"
os.makedirs(PROF, exist_ok=True)
lines_out = ["# ncu summary %s (%s, 1 x B200, --clock-control none)\n" % (tag, workload)]

ll = os.path.join(OUT, "launches_%s.csv" % workload)
if os.path.exists(ll):
    shutil.copy(ll, os.path.join(PROF, "%s_launches_%s.csv" % (tag, workload)))
    with open(ll) as f:
        rows = [l for l in f if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(rows):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        v = float(row["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(row["Metric Unit"], 1.0)
        if v < 10:      # gated launches that exit at once
            continue
        tot[k] += v
        cnt[k] += 1
    T = sum(tot.values())
    lines_out.append("## launch list: `ncu --metrics gpu__time_duration.sum` over `bench.py --steps 2 --warmup 3`"
                     " (cold-cache, serialised: shares, not absolutes)\n")
    lines_out.append("| kernel | launches | total us | avg us | share |\n|---|---|---|---|---|")
    for k, v in sorted(tot.items(), key=lambda x: -x[1]):
        lines_out.append("| `%s` | %d | %.1f | %.1f | %.3f |" % (k, cnt[k], v, v / cnt[k], v / T))
    lines_out.append("")

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__grid_size", "launch__waves_per_multiprocessor",
        "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
for rep in sorted(f for f in os.listdir(OUT) if f.endswith(".ncu-rep")):
    try:
        raw = subprocess.check_output(["ncu", "-i", os.path.join(OUT, rep), "--page", "raw", "--csv"],
                                      text=True, stderr=subprocess.DEVNULL)
    except Exception as e:
        lines_out.append("(%s unreadable: %s)" % (rep, e))
        continue
    r = list(csv.reader(raw.splitlines()))
    hdr, units, data = r[0], r[1], r[2:]
    lines_out.append("## `%s` (ncu --set full)\n" % rep)
    for row in data:
        name = re.sub(r"\(.*", "", row[hdr.index("Kernel Name")]).replace("void ", "")
        lines_out.append("**%s**\n" % name)
        lines_out.append("| metric | value | unit |\n|---|---|---|")
        vals = {}
        for w in WANT:
            if w in hdr:
                i = hdr.index(w)
                lines_out.append("| %s | %s | %s |" % (w, row[i], units[i]))
                vals[w] = (row[i], units[i])
        lines_out.append("")
        if ("k_spmv" in name or "k_llg_solve" in name) and "dram__bytes_read.sum" in vals:
            rd = float(vals["dram__bytes_read.sum"][0]) * UNIT[vals["dram__bytes_read.sum"][1]]
            wr = float(vals["dram__bytes_write.sum"][0]) * UNIT[vals["dram__bytes_write.sum"][1]]
            json.dump(dict(workload=workload, kernel=name, dram_bytes_per_launch=rd + wr,
                           source="%s_summary.md (%s)" % (tag, rep)),
                      open(os.path.join(PROF, "solve_traffic.json" if "k_llg_solve" in name else "spmv_traffic.json"), "w"))
for f in ("bench_%s.json" % workload, "bench_tube5m.json", "pytest_gpu.log"):
    p = os.path.join(OUT, f)
    if os.path.exists(p):
        shutil.copy(p, os.path.join(PROF, "%s_%s" % (tag, f)))
open(os.path.join(PROF, "%s_summary.md" % tag), "w").write("\n".join(lines_out) + "\n")
print("\n".join(lines_out[:40]))
