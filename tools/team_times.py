#!/usr/bin/env python3
"""Start / end times of every team of one large launch (variant built with -DSCF_DEBUG_TIMES): how far apart the teams
finish under the static round-robin tile assignment.  Usage: SCFEAT_LIB=.../libscfeat_dbgtimes.so python tools/team_times.py [clips [launches back to back]]"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scfeat
from scfeat import _lib

n = int(sys.argv[1]) if len(sys.argv) > 1 else 13229
chain = int(sys.argv[2]) if len(sys.argv) > 2 else 1          # launches issued back to back; the last one is reported
plan = scfeat.get_plan()
g = torch.Generator(device='cuda')
g.manual_seed(0)
pcm = torch.randint(-32768, 32768, (n, 16000), dtype=torch.int16, device='cuda', generator=g)
out = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream()
L = ctypes.CDLL(_lib.LIB_PATH)
buf = (ctypes.c_ulonglong * (8 * 2048))()
for rep in range(4):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(chain):
        plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    assert L.scf_debug_times(buf) == 0
    t = np.frombuffer(buf, dtype=np.uint64).astype(np.int64)
    n_teams = min(2048, (n * 15 + 7) // 8, 444)
    s, e = t[:n_teams], t[2048:2048 + n_teams]
    t0 = s.min()
    print('rep %d: event time %.1f us | team starts spread %.1f us | ends: first %.1f median %.1f last %.1f us after the first start '
          '(spread %.1f us, mean idle %.1f us per team)' % (rep, e0.elapsed_time(e1) * 1e3, (s.max() - t0) / 1e3, (e.min() - t0) / 1e3,
                                                             float(np.median(e) - t0) / 1e3, (e.max() - t0) / 1e3, (e.max() - e.min()) / 1e3,
                                                             float((e.max() - e).mean()) / 1e3))
    tiles = t[2 * 2048:].reshape(6, 2048)[:, :n_teams]
    prev = s
    row = []
    for i in range(6):
        ok = tiles[i] > 0
        if not ok.any():
            break
        d = (tiles[i] - prev)[ok] / 1e3
        row.append('tile %d: %.1f us (min %.1f max %.1f)' % (i, float(np.median(d)), d.min(), d.max()))
        prev = tiles[i]
    print('      per-team tile durations, median: ' + '; '.join(row))
    pbuf = (ctypes.c_ulonglong * (2 * 8 * 2048))()
    if hasattr(L, 'scf_debug_phases') and L.scf_debug_phases(pbuf) == 0 and rep == 3:
        ph = np.frombuffer(pbuf, dtype=np.uint64).astype(np.float64).reshape(2, 8, 2048)[:, :7, :n_teams]
        names = ['FFT stage', 'wait 1st barrier', 'bank', 'wait 2nd barrier', 'log', 'wait 3rd barrier', 'DCT + stores + bookkeeping']
        for w, who in enumerate(('thread 0 (warp 0: no DCT coefficient)', 'thread 224 (warp 7: no log band)')):
            tot = ph[w].sum(axis=0)
            share = ph[w].sum(axis=1) / tot.sum()
            print('      %s: ' % who + ', '.join('%s %.1f %%' % (nm, 100 * x) for nm, x in zip(names, share)) +
                  ' | %.0f cycles per tile' % (tot.mean() / max(1.0, (n * 15 / 8) / n_teams)))
