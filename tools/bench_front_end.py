#!/usr/bin/env python3
"""Pre-emphasis + Hamming on the params.json geometry (the C++ twin's front end, inference/tflite/mfcc.h:394-410):
the loader fused into the fast kernels against the generic loader on the same samples and against the plain fast path.
Device-resident int16 input, best of 6.  SCFEAT_VARIANT=0 forces the classic 16-warp kernels.
Usage: python tools/bench_front_end.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import scfeat
from scfeat.plan import PAD_NONE

g = torch.Generator(device='cuda')
g.manual_seed(0)
st = torch.cuda.current_stream()


def rate(plan, pcm, lengths=None, pad=None, reps=6, inner=1):
    n = pcm.shape[0]
    out = torch.empty((n, 30, plan.out_cols), dtype=torch.float32, device='cuda')
    kw = {} if lengths is None else dict(d_lengths=lengths.data_ptr(), pad=pad)
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(inner):
            plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream, **kw)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / inner)
    assert torch.isfinite(out).all()
    return n / best / 1e3


for n, inner in ((512, 200), (16384, 1)):
    pcm = torch.randint(-32768, 32768, (n, 16000), dtype=torch.int16, device='cuda', generator=g)
    full = torch.full((n,), 16000, dtype=torch.int32, device='cuda')
    plain = scfeat.get_plan()
    rows = [('plain fast path (no front end)', rate(plain, pcm, inner=inner))]
    for name, kw in (('pre-emphasis 0.95 + Hamming', dict(preemph_alpha=0.95, window_fn='hamming')),
                     ('Hamming only', dict(window_fn='hamming')),
                     ('pre-emphasis only', dict(preemph_alpha=0.95))):
        plan = scfeat.get_plan(**kw)
        rows.append(('%s, fused fast loader' % name, rate(plan, pcm, inner=inner)))
        rows.append(('%s, generic loader' % name, rate(plan, pcm, full, PAD_NONE, inner=inner)))
    for name, v in rows:
        print('%6d clips  %-52s %6.2f M clips/s' % (n, name, v))
