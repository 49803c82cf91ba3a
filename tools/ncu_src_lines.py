#!/usr/bin/env python
"""Per-CUDA-source-line stall samples and executed warp instructions of an ncu report captured with --import-source on
(the `--page source --print-source sass,cuda` view carries one aggregated row per source line).
usage: ncu_src_lines.py <report.ncu-rep> [top] [launch id]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"]
raw = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
fname, hdr, out, seen_fn = "?", None, {}, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No":
        hdr = r
        i_s, i_i, i_t = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
        continue
    if hdr and r[0].strip().isdigit():
        try:
            key = (fname, int(r[0]))
            v = out.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
            v[0] += int(r[i_s])
            v[1] += int(r[i_i])
            v[2] += int(r[i_t])
        except (ValueError, IndexError):
            pass
ts, ti = sum(v[0] for v in out.values()), sum(v[1] for v in out.values())
print(f"# total samples {ts} warp instructions {ti} (summed over the captured launches)")
for (f, ln), v in sorted(out.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f:24s} {ln:5d} samples {v[0]:6d} ({100 * v[0] / max(ts, 1):5.1f}%) instr {v[1]:9d} ({100 * v[1] / max(ti, 1):5.1f}%) thr/instr {v[2] / max(v[1], 1):5.1f} | {v[3]}")
