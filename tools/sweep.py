#!/usr/bin/env python3
"""Throughput vs batch size for the params.json MFCC plan (device-resident int16 input).
Usage: [SCFEAT_VARIANT=0|1|3] python tools/sweep.py   (0 classic, 1 dense 3 CTAs/SM, 3 three-team CTA; default: automatic)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import scfeat

plan = scfeat.get_plan()
st = torch.cuda.current_stream()
g = torch.Generator(device='cuda')
g.manual_seed(0)
big = torch.randint(-32768, 32768, (65536, 16000), dtype=torch.int16, device='cuda', generator=g)
out = torch.empty((65536, 30, 20), dtype=torch.float32, device='cuda')
res = []
for n in (512, 1024, 2048, 4096, 8192, 16384, 65536):
    reps = max(3, 65536 // n)
    def run():
        for i in range(reps):
            plan.extract_device(big[(i * n) % (65536 - n + 1)].data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream)
    run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    res.append('%d:%.2fM' % (n, n / best / 1e3))
print('VARIANT=%s  ' % os.environ.get('SCFEAT_VARIANT', 'auto') + '  '.join(res))
