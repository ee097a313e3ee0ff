import os, sys, time, tempfile, wave, shutil, json
sys.path.insert(0, os.getcwd())
import numpy as np
import scfeat
from scfeat import cache as scache, data_utils as sdu
n_w = 16384
tmp = tempfile.mkdtemp(prefix='scf_wavs_')
rngw = np.random.default_rng(7)
paths = []
for i in range(n_w):
    pth = os.path.join(tmp, '%05d.wav' % i)
    with wave.open(pth, 'wb') as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(16000)
        w.writeframes(rngw.integers(-32768, 32768, size=16000 if i % 7 else 12000, dtype=np.int16).tobytes())
    paths.append(pth)
res = {}
scache.ingest_wavs(paths[:2048])
for batch in (256, 512, 1024, 2048):
    for thr in (4, 16):
        t0 = time.perf_counter(); feats, lens = scache.ingest_wavs(paths, batch=batch, n_threads=thr); t = time.perf_counter() - t0
        res['batch%d_thr%d' % (batch, thr)] = n_w / t
t0 = time.perf_counter(); pcm, l = scache.load_wav_batch(paths[:8192], n_threads=16); res['read_only_thr16_files_per_s'] = 8192 / (time.perf_counter() - t0)
t0 = time.perf_counter(); pcm, l = scache.load_wav_batch(paths[:8192], n_threads=4); res['read_only_thr4_files_per_s'] = 8192 / (time.perf_counter() - t0)
t0 = time.perf_counter(); one = [sdu.get_mfcc_feature(p_) for p_ in paths[:256]]; res['one_call_per_file'] = 256 / (time.perf_counter() - t0)
res['cpus'] = len(os.sched_getaffinity(0))
print(json.dumps(res, indent=1))
shutil.rmtree(tmp, ignore_errors=True)
