#!/usr/bin/env python3
"""A few streaming steps (config 4: 256 streams x 1600-sample chunks) for ncu / timing."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import scfeat

n_streams, chunk, T = 256, 1600, 12
g = torch.Generator(device='cuda')
g.manual_seed(2)
chunks = torch.randint(-32768, 32768, (T, n_streams, chunk), dtype=torch.int16, device='cuda', generator=g)
fs = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
ring = torch.empty((n_streams, 30, 20), dtype=torch.float32, device='cuda')
new = torch.empty((n_streams,), dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream()
for t in range(T):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    fs.push_device(chunks[t].data_ptr(), chunk, ring.data_ptr(), new.data_ptr(), stream=st.cuda_stream)
    e1.record()
    torch.cuda.synchronize()
print('last step %.1f us, new rows mean %.2f' % (e0.elapsed_time(e1) * 1e3, float(new.float().mean())))
