"""Drop-in for the feature entry points of common/data_utils.py, plus the batched form the
feature-cache builder (classifier/data.py:30-46) needs.

buffer_to_audio :13-21, audio_to_buffer :25-33, add_deltas :50-58, vectorize_raw :61-70,
audio_to_feature :73-86, get_mfcc_feature :89-97.  The MFCC arithmetic (sonopy.mfcc_spec in the
reference) runs in libscfeat's CUDA kernels and is configured from the CURRENT value of the global
``pr`` at call time.  Results are float32.
"""
import wave

import numpy as np

from . import _lib
from .params import pr
from .plan import BANK_MEL_SONOPY, OUT_CEPSTRUM, PAD_FRONT_ZERO, PAD_NONE, get_plan


def _mfcc_plan(delta=None):
    """The plan for the CURRENT ``pr``; delta='diff' appends the Python path's delta columns on the device."""
    return get_plan(sample_rate=int(pr.sample_rate), window=int(pr.window_samples), hop=int(pr.hop_samples),
                    n_fft=int(pr.n_fft), n_filt=int(pr.n_filt), n_coeffs=int(pr.n_mfcc), bank=BANK_MEL_SONOPY,
                    output=OUT_CEPSTRUM, delta=delta)


def buffer_to_audio(buffer):
    """raw mono 16-bit LE byte string -> float32 in [-1, 1)  (np.fromstring in the reference; its binary
    mode is gone in numpy >= 2.3, frombuffer is the same conversion)"""
    assert pr.sample_depth == 2, 'only support 16-bit sample depth.'
    return np.frombuffer(buffer, dtype='<i2').astype(np.float32, order='C') / (np.iinfo(np.int16).max + 1)


def audio_to_buffer(audio):
    assert pr.sample_depth == 2, 'only support 16-bit sample depth.'
    return (np.asarray(audio) * (np.iinfo(np.int16).max + 1)).astype('<i2').tobytes()


def add_deltas(features, kind='diff'):
    """appends delta features on the last axis (host helper for arrays that are already on the host, as in the
    reference; the feature entry points below compute their delta columns on the GPU, scf_config.delta).
    kind='diff'    : the Python path (common/data_utils.py:50-58): first difference, row 0 -> zeros;
    kind='central' : the C++ twin (inference/tflite/mfcc.h:432-441): (f[i+1] - f[i-1]) / 2 with the edges clamped."""
    features = np.asarray(features)
    deltas = np.zeros_like(features)
    if kind == 'diff':
        deltas[1:] = features[1:] - features[:-1]
    elif kind == 'central':
        n = len(features)
        if n:
            nxt = np.minimum(np.arange(n) + 1, n - 1)
            prv = np.maximum(np.arange(n) - 1, 0)
            deltas = (features[nxt] - features[prv]) / 2
    else:
        raise ValueError("kind must be 'diff' or 'central'")
    return np.concatenate([features, deltas], -1)


def vectorize_raw(audio):
    """turns audio into feature vectors, without clipping for length
    (int16 input = PCM, scaled by 1/32768 in the loader like buffer_to_audio; float input is used as is)"""
    if len(audio) == 0:
        raise ValueError('Cannot vectorize empty audio!')      # the reference raises an undefined name here
    a = np.asarray(audio)
    if a.dtype != np.int16:
        a = a.astype(np.float32, copy=False)
    plan = _mfcc_plan()
    if plan.frames(len(a)) == 0:
        return np.empty((0, plan.out_cols), dtype=np.float32)
    return plan.extract_host(a, pad=PAD_NONE)


def audio_to_feature(audio_data):
    """audio data -> mfcc feature: keep the FIRST max_samples, pad with zeros in FRONT"""
    a = np.asarray(audio_data)[:pr.max_samples]
    if a.dtype != np.int16:
        a = a.astype(np.float32, copy=False)
    # (empty audio is front-padded to max_samples like any short clip, common/data_utils.py:79-80: all-silence rows)
    buf = np.zeros((1, pr.max_samples), dtype=a.dtype)
    buf[0, :len(a)] = a                                    # the kernel applies the front padding itself
    # pr.use_delta: the delta columns (common/data_utils.py:50-58, :85) come out of the same call
    return _mfcc_plan('diff' if pr.use_delta else None).extract_host(buf, lengths=[len(a)], pad=PAD_FRONT_ZERO)[0]


def load_wav(audio_path):
    """16-bit PCM wav -> int16 mono samples.  Stands in for librosa.load(path, sr=pr.sample_rate, mono=True)
    for files already at pr.sample_rate (resampling is out of scope)."""
    with wave.open(audio_path, 'rb') as w:
        if w.getsampwidth() != 2:
            raise ValueError('only 16-bit PCM wav is supported: ' + audio_path)
        if w.getframerate() != pr.sample_rate:
            raise ValueError('sample rate %d != pr.sample_rate %d (no resampler): %s'
                             % (w.getframerate(), pr.sample_rate, audio_path))
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype='<i2')
        ch = w.getnchannels()
    if ch > 1:
        return pcm.reshape(-1, ch).astype(np.float32).mean(axis=1) / 32768.0   # float mono mix-down
    return pcm


def get_mfcc_feature(audio_path):
    """audio sample file -> mfcc feature vectors (n_features, n_mfcc, 1)"""
    return np.expand_dims(audio_to_feature(load_wav(audio_path)), axis=-1)


def extract_features_batch(clips, lengths=None):
    """Batched audio_to_feature: clips [n, L] int16 (or float32), optional per-clip valid lengths.
    Clips are cropped to the first pr.max_samples and front-padded on the GPU.
    Returns float32 [n, n_features, feature_size, 1] -- the array classifier/data.py:101-114 assembles."""
    a = np.asarray(clips)
    if a.ndim != 2:
        raise ValueError('clips must be [n, L]')
    n, L = a.shape
    m = pr.max_samples
    if lengths is None:
        lengths = np.full((n,), min(L, m), dtype=np.int32)
    lengths = np.minimum(np.asarray(lengths, dtype=np.int32), min(L, m))
    if L != m:
        buf = np.zeros((n, m), dtype=a.dtype)
        buf[:, :min(L, m)] = a[:, :m]
        a = buf
    full = bool((lengths == m).all())
    feats = _mfcc_plan('diff' if pr.use_delta else None).extract_host(a, lengths=None if full else lengths,
                                                                      pad=PAD_FRONT_ZERO)
    return feats[..., None]
