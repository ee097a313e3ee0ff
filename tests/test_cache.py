"""Feature-cache format and PCM ingest: legacy layout round trip, parameter
fingerprint invalidation, wav batch loader; the GPU test runs the whole get_dataset flow against the oracle."""
import json
import os
import wave

import numpy as np
import pytest

import scfeat
from oracle import pipeline as opipe
from scfeat import cache


def write_wav(path, pcm, rate=16000, channels=1):
    with wave.open(str(path), 'wb') as w:
        w.setnchannels(channels)
        w.setsampwidth(2)
        w.setframerate(rate)
        w.writeframes(np.asarray(pcm, dtype='<i2').tobytes())


def make_tree(tmp_path, example):
    classes = ['up', 'down']
    lens = {'up': [16000, 9000, 20000], 'down': [16000, 1500]}
    k = 0
    for c in classes:
        d = tmp_path / 'sounds' / c
        d.mkdir(parents=True)
        for i, n in enumerate(lens[c]):
            x = np.concatenate([example[k % 8], example[(k + 1) % 8]])[:n]
            write_wav(d / ('%d.wav' % i), x)
            k += 1
    return classes


def test_load_wav_batch(tmp_path, example_pcm):
    _, pcm = example_pcm
    write_wav(tmp_path / 'a.wav', pcm[0])
    write_wav(tmp_path / 'b.wav', pcm[1][:700])
    write_wav(tmp_path / 'c.wav', np.concatenate([pcm[2], pcm[3]]))
    stereo = np.stack([pcm[4], pcm[5]], axis=1).reshape(-1)
    write_wav(tmp_path / 'd.wav', stereo, channels=2)
    got, lengths = cache.load_wav_batch([str(tmp_path / n) for n in ('a.wav', 'b.wav', 'c.wav', 'd.wav')])
    assert got.shape == (4, 16000) and got.dtype == np.int16
    assert list(lengths) == [16000, 700, 16000, 16000]
    assert np.array_equal(got[0], pcm[0]) and np.array_equal(got[1][:700], pcm[1][:700]) and not got[1][700:].any()
    assert np.array_equal(got[2], pcm[2])                                   # keeps the head
    assert np.abs(got[3].astype(np.int32) - ((pcm[4].astype(np.int32) + pcm[5]) / 2)).max() <= 1
    write_wav(tmp_path / 'e.wav', pcm[0], rate=8000)
    with pytest.raises(ValueError):
        cache.load_wav_batch([str(tmp_path / 'e.wav')])


def test_legacy_layout_round_trip_and_fingerprint(tmp_path):
    rng = np.random.default_rng(0)
    feats = [{'data': rng.normal(size=(30, 20, 1)), 'label': l} for l in ('up', 'up', 'down')]
    fp = tmp_path / 'features'
    cache.save_features(feats, str(fp))
    files = sorted(p.name for p in (fp / 'up').iterdir())
    assert len(files) == 2 and all(f.endswith('.npy') and len(f) == 36 for f in files)
    assert np.load(fp / 'up' / files[0]).dtype == np.float32
    x, y = cache.load_features(str(fp), ['up', 'down'])
    assert len(x) == 3 and sorted(y) == [0, 0, 1] and x[0].shape == (30, 20, 1)
    assert cache.cache_is_current(str(fp))
    meta = json.loads((fp / 'params.json').read_text())
    meta['n_mfcc'] = 13
    (fp / 'params.json').write_text(json.dumps(meta))
    assert not cache.cache_is_current(str(fp))
    os.remove(fp / 'params.json')
    assert cache.cache_is_current(str(fp))                                 # legacy caches carry no fingerprint
    xt, yt, xv, yv = cache.split_data(x, y, 0.34, seed=1)
    assert len(xt) == 1 and len(xv) == 2 and xt.shape[1:] == (30, 20, 1)


@pytest.mark.gpu
def test_get_dataset_end_to_end(tmp_path, example_pcm):
    _, pcm = example_pcm
    classes = make_tree(tmp_path, pcm)
    x, y, xv, yv = cache.get_dataset(str(tmp_path), classes)
    assert x.shape == (5, 30, 20, 1) and x.dtype == np.float32 and xv is None
    p = opipe.Params()
    by_label = {}
    for c in classes:
        for f in sorted((tmp_path / 'sounds' / c).iterdir()):
            want = opipe.get_mfcc_feature(str(f), p)
            by_label.setdefault(classes.index(c), []).append(want)
    for feat, label in zip(x, y):
        errs = [np.abs(feat - w).max() / np.abs(w).max() for w in by_label[int(label)]]
        assert min(errs) <= 1e-4
    # second call reuses the cache; changing pr invalidates it
    x2, y2, _, _ = cache.get_dataset(str(tmp_path), classes)
    assert x2.shape == x.shape
    pr = scfeat.params.pr
    try:
        pr.__dict__.update(n_mfcc=13)
        x3, _, _, _ = cache.get_dataset(str(tmp_path), classes)
        assert x3.shape == (5, 30, 13, 1)
    finally:
        pr.__dict__.update(n_mfcc=20)
