#!/usr/bin/env python
"""Summarises the ncu artifacts of tools/gpu_profile.sh into profiles/ (tracked):
  profiles/<tag>_launches.csv      per-kernel launch counts, total / mean device time, share of the step
  profiles/<tag>_sweep_full.txt    key metrics of the `--set full` capture of the sweep kernel (per launch)
usage: summarise_profile.py <tag> [kernel-regex]"""
import collections
import csv
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
kre = sys.argv[2] if len(sys.argv) > 2 else "k_sweep"
out_dir = os.path.join(ROOT, "profiles")
os.makedirs(out_dir, exist_ok=True)

# ---- launch list -------------------------------------------------------------------------------
src = os.path.join(ROOT, "gpurun_out", f"launches_{tag}.csv")
rows = [r for r in csv.reader(open(src)) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pp::", "")
    if not name.startswith("k_"):
        name = "(torch) " + name[:60]
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
total = sum(v[1] for v in agg.values())
with open(os.path.join(out_dir, f"{tag}_launches.csv"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES, not absolutes)\n")
    f.write("kernel,launches,total_us,mean_us,share_pct\n")
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k},{n},{ns / 1e3:.1f},{ns / 1e3 / n:.2f},{100 * ns / total:.1f}\n")
print(open(os.path.join(out_dir, f"{tag}_launches.csv")).read())

# ---- full capture ------------------------------------------------------------------------------
rep = os.path.join(ROOT, "gpurun_out", f"sweep_{tag}.ncu-rep")
if os.path.exists(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_sectors.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "smsp__average_warp_latency_per_inst_issued.ratio",
            "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]
    with open(os.path.join(out_dir, f"{tag}_sweep_full.txt"), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on -k regex:{kre}  ({len(data)} launches captured)\n")
        for w in want:
            if w in hdr:
                i = hdr.index(w)
                vals = [d[i] for d in data]
                if w == "Kernel Name":
                    vals = [re.sub(r"\(.*", "", v) for v in vals[:1]]
                f.write(f"{w} [{units[i]}]: {', '.join(vals)}\n")
        try:
            i_r, i_w = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
            ur, uw = units[i_r], units[i_w]
            scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
            tr = [float(d[i_r]) * scale[ur] + float(d[i_w]) * scale[uw] for d in data]
            f.write(f"traffic per launch (dram read + write) [byte]: {', '.join(str(int(t)) for t in tr)}  mean {sum(tr) / len(tr):.0f}\n")
        except Exception as e:  # noqa: BLE001
            f.write(f"traffic: n/a ({e})\n")
    print(open(os.path.join(out_dir, f"{tag}_sweep_full.txt")).read())
