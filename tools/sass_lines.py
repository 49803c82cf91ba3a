#!/usr/bin/env python
"""Joins an `ncu --page source --print-source sass --csv` dump with `nvdisasm -g` line info and prints
executed warp-instructions / stall samples aggregated by CUDA source line.
usage: sass_lines.py <ncu_source.csv> <nvdisasm.sass> <mangled-kernel-substring> [top]"""
import csv, re, sys, collections
src_csv, sass, kern = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
# nvdisasm: sequence of (file,line) per instruction for the kernel
seq = []; cur = None; inside = False
for ln in open(sass):
    if ln.startswith('.text.'):
        inside = kern in ln
        continue
    if not inside: continue
    m = re.match(r'\s*//## File "(.*)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', ln):
        seq.append((cur, ln.split('*/')[1].strip().rstrip(';')))
rows = list(csv.reader(open(src_csv)))
for i, r in enumerate(rows):
    if 'Source' in r and '# Samples' in r:
        hdr = r; start = i; break
si = hdr.index('Source'); ns = hdr.index('# Samples'); ie = hdr.index('Instructions Executed'); ti = hdr.index('Thread Instructions Executed')
ins = []
for r in rows[start + 1:]:
    try: ins.append((r[si], int(r[ns]), int(r[ie]), int(r[ti])))
    except Exception: pass
n = min(len(ins), len(seq))
print('# sass instrs: ncu', len(ins), 'nvdisasm', len(seq))
agg = collections.defaultdict(lambda: [0, 0, 0])
for k in range(n):
    key = seq[k][0]
    agg[key][0] += ins[k][2]; agg[key][1] += ins[k][1]; agg[key][2] += ins[k][3]
tot_i = sum(v[0] for v in agg.values()); tot_s = sum(v[1] for v in agg.values())
print('# total warp-instr', tot_i, 'samples', tot_s)
for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print('%-28s line %-5s instr %10d (%5.1f%%)  samples %6d (%5.1f%%) thr/instr %.1f' % (key[0], key[1], v[0], 100 * v[0] / tot_i, v[1], 100 * v[1] / max(tot_s, 1), v[2] / max(v[0], 1)))
