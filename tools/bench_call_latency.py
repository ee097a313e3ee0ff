#!/usr/bin/env python3
"""Latency of the drop-in calls with the reference's call shape (one clip / one chunk per call, host numpy in, host numpy
out): what a user sees after swapping the imports, next to the CPU oracle doing the same arithmetic.
Usage: python tools/bench_call_latency.py"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import scfeat
from scfeat import _lib
from scfeat.plan import PAD_FRONT_ZERO, PAD_NONE

rng = np.random.default_rng(0)
pcm = rng.integers(-32768, 32768, size=16000, dtype=np.int16)
audio = pcm.astype(np.float32) / 32768.0


def bench(name, fn, n=2000):
    for _ in range(50):
        fn()
    ts = []
    for _ in range(n):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    ts = np.sort(np.array(ts)) * 1e6
    print('%-58s p50 %7.1f us  p99 %7.1f us  (%6.1f k calls/s)' % (name, ts[len(ts) // 2], ts[int(len(ts) * 0.99)], 1e3 / ts[len(ts) // 2]))


plan = scfeat.get_plan()
out = np.empty((1, 30, 20), np.float32)
a2 = pcm[None, :].copy()
bench('Plan.extract_host, 1 clip int16 (preallocated out)', lambda: plan.extract_host(a2, out=out))
bench('Plan.extract_host, 1 clip int16', lambda: plan.extract_host(pcm))
bench('sonopy.mfcc_spec(float audio, params.json shape)', lambda: scfeat.sonopy.mfcc_spec(audio, 16000, (1024, 512), 1024, 20, 20))
bench('data_utils.audio_to_feature(float audio)', lambda: scfeat.data_utils.audio_to_feature(audio))
bench('data_utils.vectorize_raw(float audio)', lambda: scfeat.data_utils.vectorize_raw(audio))
L = _lib.lib()
import ctypes
h, ip, op = plan.handle, a2.ctypes.data, out.ctypes.data
bench('raw ctypes scf_extract_host_i16, 1 clip', lambda: L.scf_extract_host_i16(h, ip, 1, 16000, 16000, None, PAD_NONE, op))
lis = scfeat.listener.Listener()
chunk = pcm[:1600].tobytes()
bench('Listener.update_vectors(100 ms chunk)', lambda: lis.update_vectors(chunk))
try:
    from oracle import pipeline as opipe, sonopy as osonopy
    p = opipe.Params()
    bench('CPU oracle: sonopy.mfcc_spec restatement (float64 numpy)', lambda: osonopy.mfcc_spec(audio, 16000, (1024, 512), 1024, 20, 20), n=300)
    lo = opipe.ListenerOracle(p)
    bench('CPU oracle: Listener.update_vectors emulation', lambda: lo.update_vectors(chunk), n=300)
except Exception as e:
    print('oracle not available:', e)

# ---- batches from ordinary (pageable) numpy arrays: extract_features_batch, the cache builder's call ----------------
for n in (8, 64, 512, 4096):
    batch = rng.integers(-32768, 32768, size=(n, 16000), dtype=np.int16)
    fn = lambda: scfeat.data_utils.extract_features_batch(batch)
    for _ in range(5):
        fn()
    reps = max(5, min(200, 20000 // n))
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    dt = (time.perf_counter() - t0) / reps
    print('extract_features_batch(%4d pageable int16 clips)              %8.1f us per call = %7.1f k clips/s' % (n, dt * 1e6, n / dt / 1e3))
