#!/usr/bin/env python3
"""Summarises `ncu --page source --csv` output: executed warp-instructions per opcode, stall samples per opcode,
and the hottest instructions.  Usage: ncu -i X.ncu-rep --page source --csv > src.csv; python tools/ncu_source_summary.py src.csv"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hdr_i]
col = {h: i for i, h in enumerate(hdr)}
ex = defaultdict(int)
smp = defaultdict(int)
tot_ex = tot_s = 0
inst = []
for r in rows[hdr_i + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        break                                  # only the first kernel of the report
    if len(r) < len(hdr):
        continue
    sass = r[col['Source']].strip()
    toks = sass.split()
    op = toks[0]
    if op.startswith('@'):
        op = toks[1]
    op = op.split('.')[0]
    e = int(r[col['Instructions Executed']] or 0)
    s = int(r[col['# Samples']] or 0)
    ex[op] += e
    smp[op] += s
    tot_ex += e
    tot_s += s
    inst.append((s, e, sass, r))
print('total executed warp-inst %d, samples %d' % (tot_ex, tot_s))
print('%-10s %12s %7s %9s %7s' % ('opcode', 'executed', '%', 'samples', '%'))
for op, e in sorted(ex.items(), key=lambda kv: -kv[1])[:28]:
    print('%-10s %12d %6.1f%% %9d %6.1f%%' % (op, e, 100.0 * e / tot_ex, smp[op], 100.0 * smp[op] / max(1, tot_s)))
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
print('\nstall sample totals:')
tot = {h: sum(int(x[3][col[h]] or 0) for x in inst) for h in stall_cols}
for h, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v:
        print('  %-24s %8d %5.1f%%' % (h, v, 100.0 * v / max(1, tot_s)))
print('\nhottest instructions by samples:')
for s, e, sass, r in sorted(inst, key=lambda x: -x[0])[:int(sys.argv[2]) if len(sys.argv) > 2 else 25]:
    top = sorted(((int(r[col[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print('  %6d smp %8d ex  %-60s %s' % (s, e, sass[:60], ', '.join('%s=%d' % (h[6:], v) for v, h in top if v)))
