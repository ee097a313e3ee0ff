#!/usr/bin/env python3
"""Measurements of the BASELINE configs that are NOT the bench.py line:
  config 3  corpus-scale cache build on this GPU's shard (105,829 / world clips) -- extract only
  config 4  256 concurrent streams, 1600-sample chunks: step latency p50 / p99 and stream-steps/s
  config 5  long-form Bark: 1024 x 60 s clips, fft 1024: bfcc (26 filt, 13 coeff) and bark_spec (24 filt)
and the CPU reference beside each (oracle port; for config 5 the restatement of common/bark_feature.py).
Prints one JSON object.  Usage: python tools/bench_configs.py [--clips5 1024] [--cpu-seconds 5]"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import scfeat
from oracle import bark as obark, pipeline as opipe, sonopy as osonopy


def timed(fn, reps, stream):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clips5', type=int, default=1024)
    ap.add_argument('--cpu-seconds', type=float, default=5.0)
    args = ap.parse_args()
    st = torch.cuda.current_stream()
    res = {}

    # ---- config 3: one rank's shard of the 105,829-clip corpus and the whole corpus on one GPU -------------
    plan = scfeat.get_plan()
    for name, n in (('config3_shard_13229', 13229), ('config3_full_105829', 105829)):
        g = torch.Generator(device='cuda')
        g.manual_seed(1000)
        pcm = torch.randint(-32768, 32768, (n, 16000), dtype=torch.int16, device='cuda', generator=g)
        out = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
        ms = timed(lambda: plan.extract_device(pcm.data_ptr(), n, 16000, out.data_ptr(), stream=st.cuda_stream), 5, st)
        res[name] = {'clips': n, 'ms': ms, 'clips_per_s': n / ms * 1e3, 'finite': bool(torch.isfinite(out).all())}
        del pcm, out

    # ---- config 4: streaming ---------------------------------------------------------------------------------
    n_streams, chunk, T = 256, 1600, 110
    g = torch.Generator(device='cuda')
    g.manual_seed(2)
    chunks = torch.randint(-32768, 32768, (T, n_streams, chunk), dtype=torch.int16, device='cuda', generator=g)
    fs = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
    ring = torch.empty((n_streams, 30, 20), dtype=torch.float32, device='cuda')
    new = torch.empty((n_streams,), dtype=torch.int32, device='cuda')
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
    # host-inclusive latency of the public push() (H2D chunk, 3 launches, D2H ring, sync)
    h_chunks = chunks.cpu().numpy()
    fs2 = scfeat.listener.FeatureStream(n_streams, max_chunk=chunk)
    hl = []
    for t in range(T):
        t0 = time.perf_counter()
        fs2.push(h_chunks[t])
        if t >= 10:
            hl.append((time.perf_counter() - t0) * 1e6)
    hl = np.sort(np.array(hl))
    res['config4_streaming'] = {'streams': n_streams, 'chunk': chunk, 'steps': len(lat),
                                'device_step_us_p50': float(lat[len(lat) // 2]), 'device_step_us_p99': float(lat[int(len(lat) * 0.99)]),
                                'host_push_us_p50': float(hl[len(hl) // 2]), 'host_push_us_p99': float(hl[int(len(hl) * 0.99)]),
                                'stream_steps_per_s_device': n_streams / (float(lat[len(lat) // 2]) * 1e-6),
                                'new_rows_last_step_mean': float(new.float().mean())}
    # CPU: listen.py emulation, one stream
    p = opipe.Params()
    lo = opipe.ListenerOracle(p)
    t0 = time.perf_counter()
    steps = 0
    while time.perf_counter() - t0 < min(args.cpu_seconds, 3.0):
        lo.update_vectors(h_chunks[steps % T, 0].tobytes())
        steps += 1
    res['config4_streaming']['cpu_oracle_us_per_stream_step'] = (time.perf_counter() - t0) / steps * 1e6

    # ---- config 5: long-form Bark ------------------------------------------------------------------------------
    n5 = args.clips5
    g = torch.Generator(device='cuda')
    g.manual_seed(3)
    pcm = torch.randint(-32768, 32768, (n5, 960000), dtype=torch.int16, device='cuda', generator=g)
    for name, kw, cols in (('config5_bfcc_26_13', dict(n_filt=26, n_coeffs=13, output=scfeat.plan.OUT_CEPSTRUM), 13),
                           ('config5_bark_spec_24', dict(n_filt=24, output=scfeat.plan.OUT_LOG_BANK), 24)):
        pl = scfeat.get_plan(window=1024, hop=512, n_fft=1024, bank=scfeat.plan.BANK_BARK_REF, **kw)
        out = torch.empty((n5, 1874, cols), dtype=torch.float32, device='cuda')
        ms = timed(lambda: pl.extract_device(pcm.data_ptr(), n5, 960000, out.data_ptr(), stream=st.cuda_stream), 3, st)
        frames = n5 * 1874
        res[name] = {'clips': n5, 'ms': ms, 'frames_per_s': frames / ms * 1e3, 'clip_equivalents_per_s': frames / 30 / ms * 1e3,
                     'finite': bool(torch.isfinite(out).all())}
        del out
    # CPU: restatement of common/bark_feature.py on a bounded sample of 60 s clips
    h = pcm[:2].cpu().numpy().astype(np.float32) / 32768.0
    t0 = time.perf_counter()
    k = 0
    while time.perf_counter() - t0 < args.cpu_seconds:
        obark.bfcc_spec(h[k % 2], 16000, 1024, 512, 1024, 26, 13)
        k += 1
    res['config5_bfcc_26_13']['cpu_oracle_frames_per_s_1core'] = k * 1874 / (time.perf_counter() - t0)
    print(json.dumps(res, indent=1))


if __name__ == '__main__':
    main()
