#!/usr/bin/env python
"""Key metrics + stall reasons of an ncu report (first kernel rows).  usage: ncu_key.py <file.ncu-rep>"""
import csv
import subprocess
import sys

raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, u = rows[0], rows[1]
want = ["launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__average_warp_latency_per_inst_issued.ratio", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "lts__t_sectors.sum", "l1tex__t_sector_hit_rate.pct", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor"]
for w in want:
    if w in h:
        i = h.index(w)
        print(w, u[i], [r[i] for r in rows[2:]])
for i, n in enumerate(h):
    if "issue_stalled" in n and "per_issue_active" in n and "not_issued" not in n:
        v = float(rows[2][i])
        if v > 0.3:
            print(n.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v)
