"""float64 numpy restatement of the reference's callers around the MFCC arithmetic.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Follows
  * classifier/params.py:59-91        derived sizes (window/hop/max samples, n_features)
  * common/data_utils.py:13-21        buffer_to_audio  (LE int16 bytes -> float32 / 32768)
  * common/data_utils.py:50-58        add_deltas
  * common/data_utils.py:61-70        vectorize_raw
  * common/data_utils.py:73-86        audio_to_feature (keep HEAD, FRONT-pad with float64 zeros)
  * common/data_utils.py:89-97        get_mfcc_feature (expand_dims(-1))
  * listen.py:88-114                  Listener.update_vectors streaming state machine
"""
import wave
from math import floor

import numpy as np

from . import sonopy as _sonopy


class Params:
    """Derived sizes of classifier/params.py:16-91 (defaults = configs/params.json)."""

    def __init__(self, buffer_t=1.0, window_t=0.064, hop_t=0.032, sample_rate=16000,
                 sample_depth=2, n_fft=1024, n_filt=20, n_mfcc=20, use_delta=False):
        self.buffer_t, self.window_t, self.hop_t = buffer_t, window_t, hop_t
        self.sample_rate, self.sample_depth = sample_rate, sample_depth
        self.n_fft, self.n_filt, self.n_mfcc, self.use_delta = n_fft, n_filt, n_mfcc, use_delta

    @property
    def window_samples(self):
        return int(self.sample_rate * self.window_t + 0.5)

    @property
    def hop_samples(self):
        return int(self.sample_rate * self.hop_t + 0.5)

    @property
    def buffer_samples(self):
        samples = int(self.sample_rate * self.buffer_t + 0.5)
        return self.hop_samples * (samples // self.hop_samples)

    @property
    def max_samples(self):
        return int(self.buffer_t * self.sample_rate)

    @property
    def n_features(self):
        return 1 + int(floor((self.buffer_samples - self.window_samples) / self.hop_samples))

    @property
    def feature_size(self):
        return self.n_mfcc * (2 if self.use_delta else 1)


def buffer_to_audio(buffer):
    """np.fromstring(buffer, '<i2') / 32768 as float32 (np.frombuffer: fromstring's binary
    mode no longer exists in numpy >= 2.3)."""
    return np.frombuffer(buffer, dtype='<i2').astype(np.float32, order='C') / 32768.0


def read_wav_int16(path):
    """Mono 16-bit PCM wav -> (int16 array, rate).  librosa.load on such a file returns
    exactly int16/32768 as float32 (no resampling when the rate already matches)."""
    with wave.open(path, 'rb') as w:
        assert w.getsampwidth() == 2 and w.getnchannels() == 1, 'mono int16 only'
        pcm = np.frombuffer(w.readframes(w.getnframes()), dtype='<i2').copy()
        return pcm, w.getframerate()


def add_deltas(features):
    deltas = np.zeros_like(features)
    deltas[1:] = features[1:] - features[:-1]
    return np.concatenate([features, deltas], -1)


def vectorize_raw(audio, p):
    if len(audio) == 0:
        raise ValueError('Cannot vectorize empty audio!')
    return _sonopy.mfcc_spec(audio, p.sample_rate, (p.window_samples, p.hop_samples),
                             num_filt=p.n_filt, fft_size=p.n_fft, num_coeffs=p.n_mfcc)


def audio_to_feature(audio, p):
    audio = np.asarray(audio)[:p.max_samples]
    if len(audio) < p.max_samples:
        audio = np.concatenate([np.zeros((p.max_samples - len(audio),)), audio])
    feat = vectorize_raw(audio, p)
    if p.use_delta:
        feat = add_deltas(feat)
    return feat


def get_mfcc_feature(path, p):
    pcm, rate = read_wav_int16(path)
    assert rate == p.sample_rate, 'oracle does not resample'
    return np.expand_dims(audio_to_feature(pcm.astype(np.float32) / 32768.0, p), -1)


class ListenerOracle:
    """State machine of listen.py:88-114 (window_audio carry + mfccs ring)."""

    def __init__(self, p):
        self.p = p
        self.window_audio = np.array([])
        self.mfccs = np.zeros((p.n_features, p.n_mfcc))

    def update_vectors(self, chunk):
        p = self.p
        self.window_audio = np.concatenate((self.window_audio, buffer_to_audio(chunk)))
        if len(self.window_audio) >= p.window_samples:
            new = vectorize_raw(self.window_audio, p)
            self.window_audio = self.window_audio[len(new) * p.hop_samples:]
            if len(new) > len(self.mfccs):
                new = new[-len(self.mfccs):]
            self.mfccs = np.concatenate((self.mfccs[len(new):], new))
        return np.expand_dims(self.mfccs, -1)
