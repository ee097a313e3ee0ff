#!/usr/bin/env python3
"""Small case for compute-sanitizer: every kernel variant family once (fast int16, generic float with window,
n_fft 512, power output, streaming)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import scfeat

rng = np.random.default_rng(0)
pcm = rng.integers(-32768, 32768, size=(24, 16000), dtype=np.int16)
a = scfeat.data_utils.extract_features_batch(pcm)
b = scfeat.data_utils.extract_features_batch(pcm, np.array([16000, 9000, 1024, 1] * 6, dtype=np.int32))
c = scfeat.sonopy.mfcc_spec(pcm[0].astype(np.float32) / 32768, 16000, (400, 160), 512, 26, 13)
d = scfeat.sonopy.power_spec(pcm[1].astype(np.float32) / 32768, (1024, 512), 1024)
e = scfeat.get_plan(preemph_alpha=0.95, window_fn='hamming').extract_host(pcm[:3])
f = scfeat.bark_feature.bark_spec(pcm[2].astype(np.float32) / 32768, 16000, 256, 128, 256, 24)
fs = scfeat.listener.FeatureStream(6, max_chunk=1600)
for t in range(4):
    ring, new = fs.push(pcm[:6, t * 1600:(t + 1) * 1600])
for x in (a, b, c, d, e, f, ring):
    assert np.isfinite(x).all()
print('sanitize case ok', a.shape, b.shape, c.shape, d.shape, e.shape, f.shape, ring.shape, new)
