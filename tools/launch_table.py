#!/usr/bin/env python
"""Per-kernel table of an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` log:
launches, mean device time, share of the listed time and, when the DRAM counters were collected, mean DRAM bytes per launch
and the bandwidth they amount to.
usage: launch_table.py <launches.csv> [units-per-launch-for-B/unit]"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 0.
h = rows[0]
ki, mi, vi, ii = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
per = collections.OrderedDict()
for r in rows[1:]:
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pp::", "")
    if not n.startswith("k_"):
        n = "(torch) " + n
    per.setdefault((r[ii], n[:60]), {})[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for (_, n), m in per.items():
    a = agg.setdefault(n, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += m.get("gpu__time_duration.sum", 0.)
    a[2] += m.get("dram__bytes_read.sum", 0.)
    a[3] += m.get("dram__bytes_write.sum", 0.)
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':44s} {'n':>5s} {'mean us':>10s} {'share':>6s} {'rd MB':>9s} {'wr MB':>9s} {'GB/s':>8s}" + ("   B/unit" if units else ""))
for k, (n, ns, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    gbs = (rd + wr) / ns if ns else 0.
    line = f"{k:44s} {n:5d} {ns / 1e3 / n:10.2f} {100 * ns / tot:5.1f}% {rd / n / 1e6:9.2f} {wr / n / 1e6:9.2f} {gbs:8.1f}"
    if units:
        line += f" {(rd + wr) / n / units:8.1f}"
    print(line)
