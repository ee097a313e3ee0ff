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
    with pytest.raises(ValueError, match='e.wav'):
        cache.load_wav_batch([str(tmp_path / 'a.wav'), str(tmp_path / 'e.wav')])
    with pytest.raises(ValueError, match='cannot open'):
        cache.load_wav_batch([str(tmp_path / 'missing.wav')])


def test_native_wav_reader_header_variants_and_threads(tmp_path, example_pcm):
    """scf_wav_read_batch walks the RIFF chunks like Python's `wave`: chunks in front of "data" (odd sizes are padded),
    WAVE_FORMAT_EXTENSIBLE headers, data chunks shorter than their header says, empty files; any thread count gives the
    same rows as reading the files with `wave`."""
    import struct
    _, pcm = example_pcm
    rng = np.random.default_rng(3)

    def raw_wav(path, samples, fmt_ext=False, junk=b'', declared=None):
        data = np.asarray(samples, dtype='<i2').tobytes()
        if fmt_ext:
            fmt = struct.pack('<HHIIHHHHIH14s', 0xFFFE, 1, 16000, 32000, 2, 16, 22, 16, 4, 1, b'\x00\x00\x00\x00\x10\x00\x80\x00\x00\xaa\x00\x38\x9b\x71')
        else:
            fmt = struct.pack('<HHIIHH', 1, 1, 16000, 32000, 2, 16)
        body = b'WAVE' + b'fmt ' + struct.pack('<I', len(fmt)) + fmt + junk
        body += b'data' + struct.pack('<I', len(data) if declared is None else declared) + data
        with open(path, 'wb') as f:
            f.write(b'RIFF' + struct.pack('<I', len(body)) + body)

    paths, want, lens = [], [], []
    for i in range(37):
        n = int(rng.choice([0, 1, 700, 15999, 16000, 16001, 30000]))
        x = rng.integers(-32768, 32768, size=n, dtype=np.int16)
        p = str(tmp_path / ('f%02d.wav' % i))
        if i % 4 == 0:
            write_wav(p, x)
        elif i % 4 == 1:
            raw_wav(p, x, junk=b'LIST' + struct.pack('<I', 5) + b'abcde\x00')          # odd-sized chunk + pad byte
        elif i % 4 == 2:
            raw_wav(p, x, fmt_ext=True)
        else:
            raw_wav(p, x, declared=2 * n + 4000)                                      # header promises more than there is
        paths.append(p)
        row = np.zeros(16000, dtype=np.int16)
        row[:min(n, 16000)] = x[:16000]
        want.append(row)
        lens.append(min(n, 16000))
    for threads in (1, 3, 16):
        got, lengths = cache.load_wav_batch(paths, n_threads=threads)
        assert np.array_equal(got, np.stack(want)) and list(lengths) == lens
    # the files `wave` itself can read give the same samples through `wave`
    for p, row, n in zip(paths[::4], want[::4], lens[::4]):
        with wave.open(p, 'rb') as w:
            x = np.frombuffer(w.readframes(16000), dtype='<i2')
        assert np.array_equal(x, row[:n])
    not_wav = tmp_path / 'x.wav'
    not_wav.write_bytes(b'hello world, not a wav file')
    with pytest.raises(ValueError, match='not a RIFF/WAVE'):
        cache.load_wav_batch([str(not_wav)])
    write_wav(tmp_path / 'w8.wav', pcm[0])
    with wave.open(str(tmp_path / 'w8b.wav'), 'wb') as w:
        w.setnchannels(1); w.setsampwidth(1); w.setframerate(16000); w.writeframes(bytes(100))
    with pytest.raises(ValueError, match='16-bit'):
        cache.load_wav_batch([str(tmp_path / 'w8b.wav')])


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


@pytest.mark.gpu
def test_pipelined_ingest_and_device_dataset(tmp_path, example_pcm):
    """scf_ingest_wavs (reader threads -> pinned staging slots -> async extraction) equals reading the files and
    extracting them in one batch, for any slot size; get_dataset_device hands the same rows over as a DLPack tensor
    [N, 30, 20, 1] on the GPU with the label vector, without writing a feature file."""
    import torch
    _, pcm = example_pcm
    rng = np.random.default_rng(8)
    classes = ['up', 'down', 'left']
    files = []
    for c in classes:
        d = tmp_path / 'sounds' / c
        d.mkdir(parents=True)
        for i in range(23):
            n = int(rng.choice([16000, 16000, 16000, 9000, 700, 0, 25000]))
            x = np.concatenate([pcm[(i + len(c)) % 8], pcm[i % 8]])[:n]
            write_wav(d / ('%02d.wav' % i), x)
    sample_list = cache.get_sample_list(str(tmp_path / 'sounds'), classes)
    paths = [s['file'] for s in sample_list]
    raw, lengths = cache.load_wav_batch(paths)
    want = scfeat.data_utils.extract_features_batch(raw, lengths)[..., 0]
    for batch in (7, 64, 4096):
        got, lens = cache.ingest_wavs(paths, batch=batch, n_threads=4)
        assert np.array_equal(lens, lengths) and np.array_equal(got, want)
    x, y = cache.get_dataset_device(str(tmp_path), classes, batch=16)
    t = torch.from_dlpack(x)
    assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (69, 30, 20, 1)
    assert np.array_equal(t.cpu().numpy()[..., 0], want)
    assert y.dtype == np.int64 and np.array_equal(y, [classes.index(s['word']) for s in sample_list])
    assert not (tmp_path / 'features').exists()
    del t, x
    # a broken file aborts the pipeline with its name
    (tmp_path / 'sounds' / 'up' / 'zz.wav').write_bytes(b'junk')
    with pytest.raises(ValueError, match='zz.wav'):
        cache.ingest_wavs(paths + [str(tmp_path / 'sounds' / 'up' / 'zz.wav')], batch=16)
