#!/usr/bin/env python3
"""Config 4 alone (256 streams x 1600-sample chunks, device-pointer pushes): step latency launched one at a time
(p50 / p99 over 100 steps) and per step inside a CUDA graph of 20 pushes.  Usage: python tools/bench_stream.py [n_streams]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scfeat

n_streams = int(sys.argv[1]) if len(sys.argv) > 1 else 256
chunk, T = 1600, 110
g = torch.Generator(device='cuda')
g.manual_seed(2)
chunks = torch.randint(-32768, 32768, (T, n_streams, chunk), dtype=torch.int16, device='cuda', generator=g)
ring = torch.empty((n_streams, 30, 20), dtype=torch.float32, device='cuda')
new = torch.empty((n_streams,), dtype=torch.int32, device='cuda')
st = torch.cuda.current_stream()
fs = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
lat = []
for t in range(T):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(st)
    fs.push_device(chunks[t].data_ptr(), chunk, ring.data_ptr(), new.data_ptr(), stream=st.cuda_stream)
    e1.record(st)
    torch.cuda.synchronize()
    if t >= 10:
        lat.append(e0.elapsed_time(e1) * 1e3)
lat = np.sort(np.array(lat))
check = float(ring.double().sum())
fs3 = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
cs = torch.cuda.Stream()
with torch.cuda.stream(cs):
    for t in range(4):
        fs3.push_device(chunks[t].data_ptr(), chunk, ring.data_ptr(), new.data_ptr(), stream=cs.cuda_stream)
    cs.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=cs, capture_error_mode='thread_local'):
        for t in range(20):
            fs3.push_device(chunks[10 + t].data_ptr(), chunk, ring.data_ptr(), new.data_ptr(),
                            stream=torch.cuda.current_stream().cuda_stream)
    graph.replay()
    cs.synchronize()
    best = 1e30
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(cs)
        graph.replay()
        e1.record(cs)
        cs.synchronize()
        best = min(best, e0.elapsed_time(e1))
print('%d streams: one step at a time p50 %.1f us p99 %.1f us | in a 20-step graph %.2f us per step = %.1f M stream-steps/s | ring checksum %.6f'
      % (n_streams, lat[len(lat) // 2], lat[int(len(lat) * 0.99)], best * 1e3 / 20, n_streams / (best * 1e-3 / 20) / 1e6, check))
