#!/usr/bin/env python3
"""Tiny driver for ncu: a few launches of the hot path on the bench workload (512 clips, params.json MFCC).
Usage: python tools/profile_step.py [n_clips] [n_launches]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scfeat

n_clips = int(sys.argv[1]) if len(sys.argv) > 1 else 512
n_launch = int(sys.argv[2]) if len(sys.argv) > 2 else 6
plan = scfeat.get_plan()
rng = np.random.default_rng(0)
pool = [torch.from_numpy(rng.integers(-32768, 32768, size=(n_clips, 16000), dtype=np.int16)).cuda() for _ in range(3)]
out = torch.empty((n_clips, 30, 20), dtype=torch.float32, device='cuda')
st = torch.cuda.current_stream()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for i in range(n_launch):
    if i == n_launch - 1:
        e0.record()
    plan.extract_device(pool[i % 3].data_ptr(), n_clips, 16000, out.data_ptr(), stream=st.cuda_stream)
e1.record()
torch.cuda.synchronize()
print('last launch: %.2f us for %d clips -> %.2f M clips/s' % (e0.elapsed_time(e1) * 1e3, n_clips,
                                                               n_clips / e0.elapsed_time(e1) / 1e3))
assert torch.isfinite(out).all()
