"""Feature cache and dataset assembly: the callers either side of the hot path.

Mirrors classifier/data.py: get_sample_list :15-27, extract_features :30-46, save_features :49-68,
split_data :71-77, get_dataset :80-120 -- with the per-wav Python loop replaced by ONE batched GPU extraction,
and with the legacy on-disk layout kept readable and writable:

    <dataset>/sounds/<class>/*.wav                       input
    <dataset>/features/<class>/<uuid4>.npy               float32 (n_features, feature_size, 1), one file per clip

The reference never invalidates the cache when params.json changes (data.py:89-90); here a small
``features/params.json`` records the parameters the cache was built with and a mismatch triggers a rebuild
(a cache without that file is accepted as-is, so existing caches stay valid).
"""
import glob
import json
import os
import uuid
import wave
from shutil import rmtree

import numpy as np

from .data_utils import extract_features_batch
from .params import pr

_PARAM_KEYS = ('buffer_t', 'window_t', 'hop_t', 'sample_rate', 'sample_depth', 'n_fft', 'n_filt', 'n_mfcc', 'use_delta')


def params_fingerprint():
    """The fields of the current ``pr`` that determine the feature values."""
    return {k: getattr(pr, k) for k in _PARAM_KEYS}


def get_sample_list(audio_path, class_names):
    sample_list = []
    for class_name in class_names:
        class_path = os.path.join(audio_path, class_name)
        if not os.path.isdir(class_path):
            raise Exception('audio path for \'' + class_name + '\' not found at ' + class_path + '!')
        for audio_file in glob.glob(os.path.join(class_path, '*.wav')):
            sample_list.append({'file': audio_file, 'word': class_name})
    return sample_list


def load_wav_batch(paths):
    """PCM ingest for a batch of 16-bit wav files at pr.sample_rate: returns (pcm int16 [n, max_samples], lengths
    int32 [n]).  Clips keep their FIRST max_samples (common/data_utils.py:77); shorter ones are left-aligned and
    their length recorded -- the front padding happens inside the extraction kernel.  Multi-channel files are mixed
    down to mono like librosa.load(mono=True)."""
    m = pr.max_samples
    pcm = np.zeros((len(paths), m), dtype=np.int16)
    lengths = np.zeros((len(paths),), dtype=np.int32)
    for i, path in enumerate(paths):
        with wave.open(path, 'rb') as w:
            if w.getsampwidth() != 2:
                raise ValueError('only 16-bit PCM wav is supported: ' + path)
            if w.getframerate() != pr.sample_rate:
                raise ValueError('sample rate %d != pr.sample_rate %d (no resampler): %s'
                                 % (w.getframerate(), pr.sample_rate, path))
            ch = w.getnchannels()
            x = np.frombuffer(w.readframes(min(w.getnframes(), m)), dtype='<i2')
        if ch > 1:
            x = np.round(x.reshape(-1, ch).astype(np.float32).mean(axis=1)).astype(np.int16)
        pcm[i, :len(x)] = x
        lengths[i] = len(x)
    return pcm, lengths


def extract_features(audio_path, class_names, batch=8192):
    """wav tree -> list of {'data': (n_features, feature_size, 1) float32, 'label': class} like data.py:30-46,
    extracted `batch` clips per GPU call."""
    sample_list = get_sample_list(audio_path, class_names)
    features = []
    for s in range(0, len(sample_list), batch):
        chunk = sample_list[s:s + batch]
        pcm, lengths = load_wav_batch([c['file'] for c in chunk])
        if (lengths == 0).any():
            raise ValueError('Cannot vectorize empty audio: ' + chunk[int(np.argmin(lengths))]['file'])
        feats = extract_features_batch(pcm, lengths)
        features += [{'data': f, 'label': c['word']} for f, c in zip(feats, chunk)]
    return features


def save_features(features, feature_path):
    """one float32 .npy per clip under features/<label>/ (data.py:49-68) + the parameter fingerprint"""
    if os.path.isdir(feature_path):
        rmtree(feature_path)
    os.makedirs(feature_path, exist_ok=True)
    for feature in features:
        class_path = os.path.join(feature_path, feature['label'])
        os.makedirs(class_path, exist_ok=True)
        np.save(os.path.join(class_path, uuid.uuid4().hex + '.npy'), np.asarray(feature['data']).astype(np.float32))
    with open(os.path.join(feature_path, 'params.json'), 'w') as f:
        json.dump(params_fingerprint(), f, indent=2)


def cache_is_current(feature_path):
    """False only when the cache records parameters that differ from the current ``pr``."""
    meta = os.path.join(feature_path, 'params.json')
    if not os.path.isfile(meta):
        return True
    try:
        with open(meta) as f:
            return json.load(f) == json.loads(json.dumps(params_fingerprint()))
    except (OSError, ValueError):
        return False


def load_features(feature_path, class_names):
    """features/<class>/*.npy -> (x float32 [N, n_features, feature_size, 1], y int [N])  (data.py:97-114)"""
    x, y = [], []
    for feature_file in glob.glob(os.path.join(feature_path, '*', '*.npy')):
        _, class_name = os.path.split(os.path.dirname(feature_file))
        y.append(class_names.index(class_name.lower()))
        x.append(np.load(feature_file).astype(np.float32))
    return x, y


def split_data(x, y, val_split, seed=None):
    """shuffled train/val split (data.py:71-77 uses sklearn's train_test_split; same contract)"""
    n = len(x)
    idx = np.random.default_rng(seed).permutation(n)
    n_val = int(np.ceil(n * val_split))
    val, train = idx[:n_val], idx[n_val:]
    x, y = np.asarray(x), np.asarray(y)
    return x[train], y[train], x[val], y[val]


def get_dataset(dataset_path, class_names, val_split=None):
    """Same contract as classifier/data.py:80-120."""
    audio_path = os.path.join(dataset_path, 'sounds')
    feature_path = os.path.join(dataset_path, 'features')
    if os.path.exists(feature_path) and cache_is_current(feature_path):
        print('feature files path {} already exists, ignore feature extraction'.format(feature_path))
    else:
        save_features(extract_features(audio_path, class_names), feature_path)
    x, y = load_features(feature_path, class_names)
    if val_split:
        return split_data(x, y, val_split)
    return np.asarray(x), np.asarray(y), None, None
