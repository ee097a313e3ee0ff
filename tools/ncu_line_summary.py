#!/usr/bin/env python3
"""Per-CUDA-line totals from `ncu --page source --print-source cuda,sass --csv` (first kernel only):
executed warp-instructions and stall samples, grouped by file:line, plus coarse kernel regions.
Usage: python tools/ncu_line_summary.py src2.csv [top_n]"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
per_line = defaultdict(lambda: [0, 0, ''])
seen_funcs = 0
col = None
for r in rows:
    if not r:
        continue
    if r[0] == 'File Path':
        cur_file = r[1].split('/')[-1]
        continue
    if r[0] == 'Function Name':
        continue
    if r[0] == 'Line No':
        col = {h: i for i, h in enumerate(r)}
        continue
    if col is None or r[0] == '' or not r[0].isdigit():
        continue
    try:
        s = int(r[col['# Samples']] or 0)
        e = int(r[col['Instructions Executed']] or 0)
    except (ValueError, IndexError):
        continue
    k = (cur_file, int(r[0]))
    per_line[k][0] += e
    per_line[k][1] += s
    per_line[k][2] = r[1].strip()[:90]
tot_e = sum(v[0] for v in per_line.values())
tot_s = sum(v[1] for v in per_line.values())
print('total executed %d, samples %d' % (tot_e, tot_s))
by_file = defaultdict(lambda: [0, 0])
for (f, l), v in per_line.items():
    by_file[f][0] += v[0]
    by_file[f][1] += v[1]
for f, v in by_file.items():
    print('  %-24s executed %5.1f%%  samples %5.1f%%' % (f, 100.0 * v[0] / tot_e, 100.0 * v[1] / max(1, tot_s)))
print('\nlines by samples:')
for (f, l), v in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top_n]:
    print('  %-20s:%-4d ex %5.1f%% smp %5.1f%%  %s' % (f, l, 100.0 * v[0] / tot_e, 100.0 * v[1] / max(1, tot_s), v[2]))
