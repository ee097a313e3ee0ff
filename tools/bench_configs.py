#!/usr/bin/env python3
"""Measurements of the BASELINE configs that are NOT the bench.py line and of the section 8f rows:
  config 3  corpus-scale cache build on this GPU's shard (105,829 / world clips) -- extract only
  config 4  256 concurrent streams, 1600-sample chunks: step latency p50 / p99 and stream-steps/s
  config 5  long-form Bark: 1024 x 60 s clips, fft 1024: bfcc (26 filt, 13 coeff) and bark_spec (24 filt)
  f2 wav ingest, f3 delta columns, f4 post-processing (256 streams)
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
    # the same 256-stream step inside a CUDA graph (20 pushes per replay: an even number, so that the double-buffered
    # state is back where the capture started), no host work between the steps
    try:
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
            for _ in range(10):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(cs)
                graph.replay()
                e1.record(cs)
                cs.synchronize()
                best = min(best, e0.elapsed_time(e1))
        res['config4_streaming']['graph_20_steps_us_per_step'] = best * 1e3 / 20
        res['config4_streaming']['graph_stream_steps_per_s'] = n_streams / (best * 1e-3 / 20)
    except Exception as exc:
        res['config4_streaming']['graph_error'] = str(exc)[:200]

    # ---- f4: ThresholdDecoder + TriggerDetector for the same 256 streams, one launch per chunk -------------------
    from scfeat.postprocess import PostProcessor
    from oracle import postprocess as opost
    names = ['background', 'up', 'down', 'left', 'right']
    pp = PostProcessor(n_streams, names, chunk_size=chunk)
    probs = torch.softmax(torch.randn((T, n_streams, len(names)), device='cuda', generator=g), dim=-1).contiguous()
    d_idx = torch.empty((n_streams,), dtype=torch.int32, device='cuda')
    d_score = torch.empty((n_streams,), dtype=torch.float64, device='cuda')
    d_fired = torch.empty((n_streams,), dtype=torch.uint8, device='cuda')
    pl = []
    for t in range(T):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(st)
        pp.step_device(probs[t].data_ptr(), d_idx.data_ptr(), d_score.data_ptr(), d_fired.data_ptr(), stream=st.cuda_stream)
        e1.record(st)
        torch.cuda.synchronize()
        if t >= 10:
            pl.append(e0.elapsed_time(e1) * 1e3)
    pl = np.sort(np.array(pl))
    res['f4_postprocess'] = {'streams': n_streams, 'device_step_us_p50': float(pl[len(pl) // 2]),
                             'device_step_us_p99': float(pl[int(len(pl) * 0.99)])}
    try:                                       # CPU: the numpy mirror of listen.py:452-559, one stream
        dec = opost.ThresholdDecoder([(6, 4)], 0.2)
        trig = opost.TriggerDetector(chunk, names)
        hp = probs[:, 0].cpu().numpy()
        t0 = time.perf_counter()
        k = 0
        while time.perf_counter() - t0 < 1.0:
            pr_ = hp[k % T]
            i = int(np.argmax(pr_))
            sc = float(np.max(pr_))
            if names[i] != 'background':
                sc = dec.decode(sc)
            trig.update(i, sc)
            k += 1
        res['f4_postprocess']['cpu_oracle_us_per_stream_step'] = (time.perf_counter() - t0) / k * 1e6
    except Exception as exc:
        res['f4_postprocess']['cpu_error'] = str(exc)[:200]

    # ---- f3: delta columns on the device (first difference / central + delta-delta) over 8192 clips ------------------
    n3 = 8192
    pcm3 = torch.randint(-32768, 32768, (n3, 16000), dtype=torch.int16, device='cuda', generator=g)
    res['f3_deltas'] = {}
    for name, kw, mult in (('none', {}, 1), ('diff', dict(delta='diff'), 2), ('central2', dict(delta='central2'), 3)):
        pd = scfeat.get_plan(**kw)
        o3 = torch.empty((n3, 30, 20 * mult), dtype=torch.float32, device='cuda')
        ms = timed(lambda: pd.extract_device(pcm3.data_ptr(), n3, 16000, o3.data_ptr(), stream=st.cuda_stream), 5, st)
        res['f3_deltas'][name] = {'ms': ms, 'clips_per_s': n3 / ms * 1e3}
        del o3
    del pcm3

    # ---- f2 / f1: wav files -> features (pipelined native ingest) vs one file at a time -----------------------------
    import tempfile
    import wave
    from scfeat import cache as scache, data_utils as sdu
    n_w = 16384
    tmp = tempfile.mkdtemp(prefix='scf_wavs_')
    rngw = np.random.default_rng(7)
    paths = []
    for i in range(n_w):
        pth = os.path.join(tmp, '%05d.wav' % i)
        with wave.open(pth, 'wb') as w:
            w.setnchannels(1)
            w.setsampwidth(2)
            w.setframerate(16000)
            w.writeframes(rngw.integers(-32768, 32768, size=16000 if i % 7 else 12000, dtype=np.int16).tobytes())
        paths.append(pth)
    scache.ingest_wavs(paths[:2048])                          # (first call: staging allocation, page cache)
    t0 = time.perf_counter()
    feats, lens = scache.ingest_wavs(paths)
    t_ing = time.perf_counter() - t0
    t0 = time.perf_counter()
    one = [sdu.get_mfcc_feature(p_) for p_ in paths[:256]]         # the reference's call shape: one wav, one call
    t_one = (time.perf_counter() - t0) / 256
    res['f2_ingest'] = {'files': n_w, 'pipelined_clips_per_s': n_w / t_ing, 'one_call_per_file_clips_per_s': 1.0 / t_one,
                        'same_result': bool(np.allclose(feats[:256], np.stack(one)[..., 0], atol=1e-5)),
                        'note': 'files in the page cache (tmp dir), 16-bit mono 16 kHz, every 7th one 0.75 s long'}
    import shutil
    shutil.rmtree(tmp, ignore_errors=True)

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
