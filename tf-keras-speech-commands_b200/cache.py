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
import ctypes
import glob
import json
import os
import uuid
from shutil import rmtree

import numpy as np

from . import _lib
from .data_utils import _mfcc_plan
from .params import pr

READER_THREADS = min(16, len(os.sched_getaffinity(0)))

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


def load_wav_batch(paths, n_threads=None):
    """PCM ingest for a batch of 16-bit wav files at pr.sample_rate: returns (pcm int16 [n, max_samples], lengths
    int32 [n]).  Clips keep their FIRST max_samples (common/data_utils.py:77); shorter ones are left-aligned and
    their length recorded -- the front padding happens inside the extraction kernel.  Multi-channel files are mixed
    down to mono like librosa.load(mono=True).  Header parse and reads run in libscfeat's reader threads
    (scf_wav_read_batch); a file that is not 16-bit PCM at pr.sample_rate raises ValueError (no resampler)."""
    m = int(pr.max_samples)
    pcm = np.empty((len(paths), m), dtype=np.int16)
    lengths = np.zeros((len(paths),), dtype=np.int32)
    if len(paths):
        arr, keep = _lib.path_array(paths)
        rc = _lib.lib().scf_wav_read_batch(arr, len(paths), int(pr.sample_rate), m, pcm.ctypes.data, m, lengths.ctypes.data,
                                           int(n_threads or READER_THREADS))
        if rc:
            raise ValueError(_lib.lib().scf_last_error().decode('utf-8', 'replace'))
    return pcm, lengths


def ingest_wavs(paths, batch=512, n_threads=None, out=None):
    """wav files -> float32 [n, n_features, feature_size] through the pipelined native path (scf_ingest_wavs): reader
    threads fill pinned staging slots while earlier slots are uploaded, transformed and downloaded.  Returns
    (features, lengths)."""
    plan = _mfcc_plan('diff' if pr.use_delta else None)
    m = int(pr.max_samples)
    shape = (len(paths), plan.frames(m), plan.out_cols)
    if out is None:
        out = np.empty(shape, dtype=np.float32)
    elif out.shape != shape or out.dtype != np.float32 or not out.flags['C_CONTIGUOUS']:
        raise ValueError('out must be a C-contiguous float32 array of shape %r' % (shape,))
    lengths = np.zeros((len(paths),), dtype=np.int32)
    if len(paths):
        arr, keep = _lib.path_array(paths)
        rc = _lib.lib().scf_ingest_wavs(plan.handle, arr, len(paths), m, int(batch), int(n_threads or READER_THREADS),
                                        out.ctypes.data, lengths.ctypes.data)
        if rc == -1:
            raise ValueError(_lib.lib().scf_last_error().decode('utf-8', 'replace'))
        _lib.check(rc)
    return out, lengths


def extract_features(audio_path, class_names, batch=512):
    """wav tree -> list of {'data': (n_features, feature_size, 1) float32, 'label': class} like data.py:30-46, through
    the pipelined ingest (one file of zero length gives all-silence rows, as the reference's front padding does)."""
    sample_list = get_sample_list(audio_path, class_names)
    feats, _ = ingest_wavs([c['file'] for c in sample_list], batch)
    return [{'data': f[..., None], 'label': c['word']} for f, c in zip(feats, sample_list)]


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


def labels_of(sample_list, class_names):
    """int64 label vector in sample order (data.py:107-110)"""
    return np.asarray([class_names.index(s['word'].lower()) for s in sample_list], dtype=np.int64)


def get_dataset_device(dataset_path, class_names, batch=512, device=-1):
    """The training set straight from the wav tree to the framework, without the 105k one-clip .npy files of
    data.py:49-68 and :97-114: reader threads + pinned staging feed the extraction kernels (scf_ingest_wavs_device), the
    features stay on the GPU.  Returns (x, y): x = DLPack capsule "dltensor", float32 [N, n_features, feature_size, 1]
    on the device (tf.experimental.dlpack.from_dlpack(x) / torch.from_dlpack(x)); y = int64 labels [N] (host).
    Sample order = get_sample_list order; shuffle / split on the consumer's side (model.fit(shuffle=True))."""
    from .plan import dlpack_alloc
    sample_list = get_sample_list(os.path.join(dataset_path, 'sounds'), class_names)
    plan = _mfcc_plan('diff' if pr.use_delta else None)
    m = int(pr.max_samples)
    d_ptr, finish = dlpack_alloc((len(sample_list), plan.frames(m), plan.out_cols, 1), device)
    if sample_list:
        arr, keep = _lib.path_array([c['file'] for c in sample_list])
        rc = _lib.lib().scf_ingest_wavs_device(plan.handle, arr, len(sample_list), m, int(batch), READER_THREADS, d_ptr, None)
        if rc == -1:
            finish()            # (an unconsumed capsule frees the buffer)
            raise ValueError(_lib.lib().scf_last_error().decode('utf-8', 'replace'))
        _lib.check(rc)
    return finish(), labels_of(sample_list, class_names)


def get_dataset(dataset_path, class_names, val_split=None):
    """Same contract as classifier/data.py:80-120 (legacy on-disk cache; see get_dataset_device for the path without it)."""
    audio_path = os.path.join(dataset_path, 'sounds')
    feature_path = os.path.join(dataset_path, 'features')
    if os.path.exists(feature_path) and cache_is_current(feature_path):
        print('feature files path {} already exists, ignore feature extraction'.format(feature_path))
    else:
        if os.path.exists(feature_path):
            print('WARNING: feature files path {} was built with other parameters than the current ones '
                  '(features/params.json): rebuilding it'.format(feature_path))
        save_features(extract_features(audio_path, class_names), feature_path)
    x, y = load_features(feature_path, class_names)
    if val_split:
        return split_data(x, y, val_split)
    return np.asarray(x), np.asarray(y), None, None
