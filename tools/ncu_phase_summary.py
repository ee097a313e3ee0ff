#!/usr/bin/env python3
"""Per-phase budget of extract_kernel from `ncu --page source --csv` (SASS view, first kernel of the report):
executed warp-instructions, packed-FP32 instructions, shared-memory wavefronts and stall samples, per frame pair.
Phases are cut at marker instructions in address order (first sample load, the table mbarrier wait after pass 1, the
first mirror shuffle, the team barriers), so the tool follows the kernel's structure, not line numbers.
Usage: ncu -i X.ncu-rep --page source --csv > sass.csv; python tools/ncu_phase_summary.py sass.csv n_pairs"""
import csv
import sys
from collections import defaultdict

rows = list(csv.reader(open(sys.argv[1])))
n_pairs = float(sys.argv[2])
hi = next(i for i, r in enumerate(rows) if r and r[0] == 'Address')
hdr = rows[hi]
col = {h: i for i, h in enumerate(hdr)}
inst = []
for r in rows[hi + 1:]:
    if r and r[0] in ('Kernel Name', 'Address'):
        break
    if len(r) < len(hdr):
        continue
    inst.append(r)


def num(r, name):
    v = r[col[name]] if name in col else ''
    try:
        return float(v or 0)
    except ValueError:
        return 0.0


def opcode(r):
    toks = r[col['Source']].split()
    op = toks[1] if toks[0].startswith('@') else toks[0]
    return op


hot = max(num(r, 'Instructions Executed') for r in inst)     # once-per-pair instructions execute `hot` times
phase_names = ['head', 'load', 'convert+pass1+xchg st', 'pass2', 'separate+power rows', 'bank', 'log', 'dct+store', 'tail']
cuts = []
state = 0
phase_of = []
bars = 0
for i, r in enumerate(inst):
    op = opcode(r)
    ex = num(r, 'Instructions Executed')
    if state == 0 and op.startswith('LDG') and 'S16' in op and ex >= 0.5 * hot:
        state = 1
    elif state == 1 and op.startswith('I2FP'):
        state = 2
    elif state == 2 and op.startswith('SYNCS.PHASECHK'):
        state = 3
    elif state == 3 and op.startswith('SHFL') and ex >= 0.5 * hot:
        state = 4
    elif state in (4, 5, 6) and op.startswith('BAR') and ex >= 0.5 * hot:
        phase_of.append(state)        # the barrier belongs to the phase it ends
        state += 1
        continue
    elif state == 7 and op.startswith('BRA') and ex >= 0.5 * hot and i > len(inst) - 30:
        phase_of.append(state)
        state = 8
        continue
    phase_of.append(state)

stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = defaultdict(lambda: defaultdict(float))
for r, ph in zip(inst, phase_of):
    a = agg[ph]
    ex = num(r, 'Instructions Executed')
    op = opcode(r)
    a['ex'] += ex
    a['smp'] += num(r, '# Samples')
    a['wf'] += num(r, 'L1 Wavefronts Shared')
    a['wf_ideal'] += num(r, 'L1 Wavefronts Shared Ideal')
    if op.split('.')[0] in ('FFMA2', 'FADD2', 'FMUL2'):
        a['fp2'] += ex
    if op.split('.')[0] in ('LDS', 'STS', 'LDG', 'STG', 'SHFL', 'LDSM', 'ATOMS'):
        a['lsu'] += ex
    for h in stall_cols:
        a[h] += num(r, h)
tot = defaultdict(float)
for a in agg.values():
    for k, v in a.items():
        tot[k] += v
print('%d SASS instructions; per frame pair (%d pairs): %.0f warp-instructions, %.0f packed FP32, %.0f LSU-pipe instructions, '
      '%.0f shared-memory wavefronts (ideal %.0f); %d stall samples'
      % (len(inst), n_pairs, tot['ex'] / n_pairs, tot['fp2'] / n_pairs, tot['lsu'] / n_pairs, tot['wf'] / n_pairs,
         tot['wf_ideal'] / n_pairs, tot['smp']))
print('%-24s %8s %7s %7s %7s %8s   %s' % ('phase', 'instr', 'fp32x2', 'lsu', 'smem wf', 'samples', 'top stalls (share of the phase\'s samples)'))
for ph in range(len(phase_names)):
    a = agg.get(ph)
    if not a:
        continue
    top = sorted(((a[h], h[6:]) for h in stall_cols), reverse=True)[:4]
    print('%-24s %8.1f %7.1f %7.1f %7.1f %7.1f%%   %s' % (
        phase_names[ph], a['ex'] / n_pairs, a['fp2'] / n_pairs, a['lsu'] / n_pairs, a['wf'] / n_pairs,
        100 * a['smp'] / max(1, tot['smp']), ', '.join('%s %.0f%%' % (n, 100 * v / max(1, a['smp'])) for v, n in top if v)))
print('stalls overall: ' + ', '.join('%s %.1f%%' % (h[6:], 100 * tot[h] / max(1, tot['smp']))
                                    for h in sorted(stall_cols, key=lambda h: -tot[h])[:10]))
if len(sys.argv) > 3:          # per-phase opcode mix (executed per pair)
    for ph in range(len(phase_names)):
        ops = defaultdict(float)
        for r, p_ in zip(inst, phase_of):
            if p_ == ph:
                ops[opcode(r).split('.')[0]] += num(r, 'Instructions Executed') / n_pairs
        if ops:
            print('%-24s %s' % (phase_names[ph], ' '.join('%s %.1f' % (k, v) for k, v in sorted(ops.items(), key=lambda kv: -kv[1]) if v >= 0.5)))
