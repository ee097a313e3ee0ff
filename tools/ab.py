#!/usr/bin/env python3
"""A/B timing of libscfeat builds: for every library given (or the product build), times the bench workload
(512-clip batches back to back, 16-batch pool) and a large batch.  Each library runs in its own process.
Usage: python tools/ab.py [lib.so ...]"""
import os
import subprocess
import sys

CHILD = r'''
import os, sys, numpy as np, torch
sys.path.insert(0, os.getcwd())
import scfeat
plan = scfeat.get_plan()
st = torch.cuda.current_stream()
g = torch.Generator(device='cuda'); g.manual_seed(0)
pool = torch.randint(-32768, 32768, (16, 512, 16000), dtype=torch.int16, device='cuda', generator=g)
out = torch.empty((16, 512, 30, 20), dtype=torch.float32, device='cuda')      # a 16-batch feature cache
def run(k, rotate):
    for i in range(k):
        plan.extract_device(pool[i % 16].data_ptr(), 512, 16000, out[i % 16 if rotate else 0].data_ptr(), stream=st.cuda_stream)
res = []
for rotate in (True, False):
    run(50, rotate); torch.cuda.synchronize()
    b = 1e9
    for rep in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(1000, rotate); e1.record(); torch.cuda.synchronize()
        b = min(b, e0.elapsed_time(e1) / 1000)
    res.append(b)
best, same = res
big = pool.view(8192, 16000)
bout = torch.empty((8192, 30, 20), dtype=torch.float32, device='cuda')
bb = 1e9
for rep in range(8):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.extract_device(big.data_ptr(), 8192, 16000, bout.data_ptr(), stream=st.cuda_stream); e1.record(); torch.cuda.synchronize()
    bb = min(bb, e0.elapsed_time(e1))
del pool, big
huge = torch.randint(-32768, 32768, (49152, 16000), dtype=torch.int16, device='cuda', generator=g)
hout = torch.empty((49152, 30, 20), dtype=torch.float32, device='cuda')
hb = 1e9
for rep in range(6):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.extract_device(huge.data_ptr(), 49152, 16000, hout.data_ptr(), stream=st.cuda_stream); e1.record(); torch.cuda.synchronize()
    hb = min(hb, e0.elapsed_time(e1))
ref_path = os.environ.get('AB_REF', '/tmp/ab_ref.npy')
first = out[0].cpu().numpy()
if os.path.exists(ref_path):
    ref = np.load(ref_path)
    d = np.abs(first - ref).max() / np.abs(ref).max()
    note = 'bit-exact vs first lib' if (first == ref).all() else 'max|diff|/max|ref| vs first lib = %.2e' % d
else:
    np.save(ref_path, first)
    note = 'reference output saved'
print('%-28s  512-batch %.2f us/step = %.2f M clips/s (one output buffer: %.2f) | 8192 clips %.1f us = %.2f M clips/s | 49152 clips %.2f M clips/s | %s' % (
    os.path.basename(os.environ.get('SCFEAT_LIB', 'product')), best * 1e3, 512 / best / 1e3, 512 / same / 1e3, bb * 1e3, 8192 / bb / 1e3,
    49152 / hb / 1e3, note))
'''

rounds = int(os.environ.get('AB_ROUNDS', '2'))
libs = sys.argv[1:] or [None]
if os.path.exists(os.environ.get('AB_REF', '/tmp/ab_ref.npy')):
    os.remove(os.environ.get('AB_REF', '/tmp/ab_ref.npy'))
for rnd in range(rounds):
    for lib in libs:
        env = dict(os.environ)
        if lib:
            env['SCFEAT_LIB'] = os.path.abspath(lib)
        subprocess.run([sys.executable, '-c', CHILD], env=env, check=False)
