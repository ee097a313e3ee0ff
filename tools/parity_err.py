#!/usr/bin/env python3
"""Parity margins of the CUDA path against the float64 oracle (budget: 1e-4 relative on cepstra, 1e-3 absolute on log
features): the 8 example wavs, 64 uniform-noise clips, and a full-scale off-bin sine.  Usage: python tools/parity_err.py"""
import os, sys, numpy as np
sys.path.insert(0, os.getcwd())
import scfeat
from oracle import sonopy as osonopy
z = np.load('tests/golden/example_pcm.npz'); pcm = z['pcm']
got = scfeat.data_utils.extract_features_batch(pcm)[..., 0].astype(np.float64)
want = np.stack([osonopy.mfcc_spec(a.astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20, 20) for a in pcm])
err = np.abs(got - want).reshape(8, -1).max(axis=1); scale = np.abs(want).reshape(8, -1).max(axis=1)
rng = np.random.default_rng(0); x = rng.integers(-32768, 32768, size=(64, 16000), dtype=np.int16)
g2 = scfeat.data_utils.extract_features_batch(x)[..., 0].astype(np.float64)
w2 = np.stack([osonopy.mfcc_spec(a.astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20, 20) for a in x])
t = np.arange(16000); tone = (np.round(20000 * np.sin(2 * np.pi * 1000.5 * t / 16000))).astype(np.int16)[None]
g3 = scfeat.sonopy.mel_spec(tone[0].astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20).astype(np.float64)
w3 = osonopy.mel_spec(tone[0].astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20)
print(os.path.basename(os.environ.get('SCFEAT_LIB', 'product')), 'wavs rel %.3g | noise rel %.3g | tone logmel abs %.3g' % ((err / scale).max(), (np.abs(g2 - w2).max(axis=(1, 2)) / np.abs(w2).max(axis=(1, 2))).max(), np.abs(g3 - w3).max()))
