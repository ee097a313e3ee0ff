#!/usr/bin/env python3
"""Randomised self-consistency stress: for random plans (FFT size, bank, filters, coefficients, output kind, front end)
and random job shapes (clip length, clip count, ragged lengths), one big launch (dynamic tile schedule, three-team CTAs)
must equal the same clips pushed through in small launches (round robin), bit for bit, and streams must match the batch
features of their concatenated audio (1e-4 of the largest coefficient; rows not yet produced stay zero).  Usage: python tools/stress_schedule.py [n_cases] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scfeat
from scfeat import _lib

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0
for case in range(n_cases):
    n_fft = int(rng.choice([256, 512, 1024]))
    bank = int(rng.choice([_lib.BANK_MEL_SONOPY, _lib.BANK_BARK_REF]))
    n_filt = int(rng.integers(8, 27))
    n_coeffs = int(rng.integers(4, 30))
    output = int(rng.choice([_lib.OUT_CEPSTRUM, _lib.OUT_LOG_BANK, _lib.OUT_POWER], p=[0.6, 0.3, 0.1]))
    pre = rng.random() < 0.3
    kw = dict(window=n_fft, hop=n_fft // 2, n_fft=n_fft, bank=bank, n_filt=n_filt, n_coeffs=n_coeffs, output=output)
    if pre:
        kw.update(preemph_alpha=0.95, window_fn='hamming')
    try:
        plan = scfeat.get_plan(**kw)
    except scfeat.ScfError as e:
        print('case %d: plan rejected (%s)' % (case, str(e)[:60]))
        continue
    clip_len = int(rng.integers(n_fft, 20000))
    frames = plan.frames(clip_len)
    pairs = (frames + 1) // 2
    n = int(rng.integers(1, 3)) * (12000 // max(1, pairs)) + int(rng.integers(0, 50))          # well above 3 tiles per team
    pcm = torch.randint(-32768, 32768, (n, clip_len), dtype=torch.int16, device='cuda')
    ragged = rng.random() < 0.5
    lengths = torch.from_numpy(rng.integers(0, clip_len + 1, size=n).astype(np.int32)).cuda() if ragged else None
    big = torch.zeros((n, frames, plan.out_cols), dtype=torch.float32, device='cuda')
    small = torch.zeros_like(big)
    kwl = lambda a: {} if lengths is None else dict(d_lengths=lengths[a:].data_ptr())
    plan.extract_device(pcm.data_ptr(), n, clip_len, big.data_ptr(), **kwl(0))
    step = max(1, 300 // max(1, pairs))
    for a in range(0, n, step):
        plan.extract_device(pcm[a].data_ptr(), min(step, n - a), clip_len, small[a].data_ptr(), **kwl(a))
    torch.cuda.synchronize()
    same = bool(torch.equal(big, small)) and bool(torch.isfinite(big).all())
    bad += 0 if same else 1
    print('case %2d: n_fft %4d bank %d filt %2d coeffs %2d out %d pre %d clip_len %5d n %5d ragged %d -> %s' % (
        case, n_fft, bank, n_filt, n_coeffs, output, pre, clip_len, n, ragged, 'ok' if same else 'MISMATCH %g' % float((big - small).abs().max())))
# streams of random shapes against the batch features of the same audio
for case in range(6):
    n_streams = int(rng.integers(1, 300))
    chunk = int(rng.choice([160, 512, 1024, 1600, 3000, 4100]))
    steps = int(rng.integers(3, 12))
    rows = int(rng.choice([3, 10, 30]))
    audio = rng.integers(-32768, 32768, size=(n_streams, steps * chunk), dtype=np.int16)
    fs = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk, ring_rows=rows) if 'ring_rows' in scfeat.listener.FeatureStream.__init__.__code__.co_varnames \
        else scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
    ring = None
    for t in range(steps):
        ring, new = fs.push(np.ascontiguousarray(audio[:, t * chunk:(t + 1) * chunk]))
    plan = scfeat.get_plan()
    feats = plan.extract_host(audio, pad=scfeat.plan.PAD_NONE) if audio.shape[1] >= 1024 else np.zeros((n_streams, 0, 20), np.float32)
    k = min(feats.shape[1], ring.shape[1])
    # (not bit for bit: a frame's partner in the two-for-one FFT depends on where the chunk boundaries fall)
    a_, b_ = ring[:, ring.shape[1] - k:].astype(np.float64), feats[:, feats.shape[1] - k:].astype(np.float64)
    same = k == 0 or (np.abs(a_ - b_).max() <= 1e-4 * max(1.0, np.abs(b_).max()) and not ring[:, :ring.shape[1] - k].any())
    bad += 0 if same else 1
    print('stream case %d: %3d streams chunk %4d steps %2d ring rows %2d -> %s' % (case, n_streams, chunk, steps, ring.shape[1], 'ok' if same else 'MISMATCH'))
print('STRESS_OK' if bad == 0 else 'STRESS_FAILED %d' % bad)
sys.exit(0 if bad == 0 else 1)
