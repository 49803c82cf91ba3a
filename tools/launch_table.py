#!/usr/bin/env python
"""Per-kernel table (launches, mean device time, share) of an `ncu --metrics gpu__time_duration.sum --csv` log.
usage: launch_table.py <launches.csv>"""
import collections
import csv
import re
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]
ki, vi = h.index("Kernel Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("pp::", "")
    a = agg.setdefault(n[:50], [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(",", ""))
tot = sum(v[1] for v in agg.values())
for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:52s} {n:4d} {ns / 1e3 / n:9.2f} us  {100 * ns / tot:5.1f}%")
