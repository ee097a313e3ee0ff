#!/usr/bin/env python3
"""Throughput of the fast path for the three FFT sizes (window = n_fft, hop = n_fft / 2), device-resident int16 input:
MFCC 20/20 and the reference's Bark defaults (26 filters, 13 coefficients) -- and of the generic loader on sonopy's
default geometry (window 160, hop 80, n_fft 512), and of params.json with pre-emphasis + Hamming (fused into the fast loader).
Usage: python tools/bench_fft_sizes.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import scfeat
from scfeat import _lib

n = 16384
g = torch.Generator(device='cuda')
g.manual_seed(0)
pcm = torch.randint(-32768, 32768, (n, 16000), dtype=torch.int16, device='cuda', generator=g)
st = torch.cuda.current_stream()
for n_fft in (1024, 512, 256):
    for name, kw in (('mfcc 20/20', dict(bank=_lib.BANK_MEL_SONOPY, n_filt=20, n_coeffs=20)),
                     ('bfcc 26/13', dict(bank=_lib.BANK_BARK_REF, n_filt=26, n_coeffs=13))):
        plan = scfeat.get_plan(window=n_fft, hop=n_fft // 2, n_fft=n_fft, output=_lib.OUT_CEPSTRUM, **kw)
        frames = (16000 - n_fft) // (n_fft // 2) + 1
        out = torch.empty((n, frames, plan.out_cols), dtype=torch.float32, device='cuda')
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        assert torch.isfinite(out).all()
        print('n_fft %4d %-10s %6.2f G frames/s  %6.2f M clips/s  (%5.1f GB/s of PCM)' % (
            n_fft, name, n * frames / best / 1e6, n / best / 1e3, n * 32000 / best / 1e6))

for name, kw in (('sonopy defaults: window 160 hop 80 n_fft 512, 20/13', dict(window=160, hop=80, n_fft=512, n_filt=20, n_coeffs=13)),
                 ('params.json + pre-emphasis 0.95 + Hamming', dict(window=1024, hop=512, n_fft=1024, n_filt=20, n_coeffs=20,
                                                                     preemph_alpha=0.95, window_fn='hamming'))):
    plan = scfeat.get_plan(bank=_lib.BANK_MEL_SONOPY, output=_lib.OUT_CEPSTRUM, **kw)
    frames = (16000 - kw['window']) // kw['hop'] + 1
    out = torch.empty((n, frames, plan.out_cols), dtype=torch.float32, device='cuda')
    best = 1e9
    for _ in range(6):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    assert torch.isfinite(out).all()
    print('other geometry / front end: %-40s %6.2f G frames/s  %6.2f M clips/s' % (name, n * frames / best / 1e6, n / best / 1e3))

# ragged batch: per-clip lengths with front padding (common/data_utils.py:77-80); ~10 % of Speech Commands clips are short
plan = scfeat.get_plan()
lengths = torch.full((n,), 16000, dtype=torch.int32, device='cuda')
short = torch.rand((n,), device='cuda', generator=g) < 0.1
lengths[short] = torch.randint(4000, 16000, (int(short.sum()),), dtype=torch.int32, device='cuda', generator=g)
out = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
best = 1e9
for _ in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), d_lengths=lengths.data_ptr(), stream=st.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
assert torch.isfinite(out).all()
print('params.json, 10 %% of the clips short and front-padded (lengths given): %6.2f M clips/s' % (n / best / 1e3))
