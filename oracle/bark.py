"""float64 numpy restatement of /root/reference/common/bark_feature.py (Bark bank, BFCC).

TEST INFRASTRUCTURE (see oracle/__init__.py).  Pinned against the real reference file by
tests/golden/make_golden.py (which imports it with a stub ``librosa``) -- the fixtures it
writes are checked in tests/test_oracle.py.

Reference quirks restated on purpose (SURVEY.md section 8 a-notes 3):
  * bark2fft / fft2bark are called WITHOUT nfft / sample_rate (bark_feature.py:112,134), so
    the bin mapping always uses nfft=512, sample_rate=16000 and the factor (nfft+1)
    (bark_feature.py:47-56) whatever the caller passed;
  * filter i covers bins [bins[i], bins[i+4]) with centre bark_points[i+2] (:130-135).
"""
from functools import lru_cache

import numpy as np
from scipy.fftpack import dct

from .sonopy import power_spec, safe_log

_MAP_NFFT = 512        # bark_feature.py:47,52 defaults, never overridden by bark_filterbanks
_MAP_RATE = 16000


def hz2bark(f):
    return 6.0 * np.arcsinh(np.asarray(f, dtype=float) / 600.0)     # bark_feature.py:27-29


def bark2hz(b):
    return 600.0 * np.sinh(np.asarray(b, dtype=float) / 6.0)        # bark_feature.py:32-34


def fft2bark(k):
    return hz2bark((np.asarray(k, dtype=float) * _MAP_RATE) / (_MAP_NFFT + 1))   # :47-49


def bark2fft(b):
    return (_MAP_NFFT + 1) * bark2hz(b) / _MAP_RATE                 # :52-56


def skirt(fb, fc):
    """Bark critical-band shape Fm (bark_feature.py:59-72), vectorised over fb."""
    fb = np.asarray(fb, dtype=float)
    d = fb - fc
    out = np.zeros_like(fb)
    # comparisons are written against fc +/- const exactly as the reference does, so that
    # boundary bins round the same way
    rise = (fc - 2.5 <= fb) & (fb <= fc - 0.5)
    flat = (fc - 0.5 < fb) & (fb < fc + 0.5)
    fall = (fc + 0.5 <= fb) & (fb <= fc + 1.3)
    out[rise] = 10.0 ** (2.5 * (d[rise] + 0.5))
    out[flat] = 1.0
    out[fall] = 10.0 ** (-2.5 * (d[fall] - 0.5))
    return out


def scale_factors(nfilts, scale):
    """Per-filter amplitude c (bark_feature.py:118-129)."""
    c = 1.0 if scale in ("descendant", "constant") else 0.0
    out = []
    for _ in range(nfilts):
        if scale == "descendant":
            c -= 1 / nfilts
            c = c * (c > 0) + 0 * (c < 0)
        elif scale == "ascendant":
            c += 1 / nfilts
            c = c * (c < 1) + 1 * (c > 1)
        out.append(c)
    return out


@lru_cache(maxsize=None)
def bark_filterbanks(nfilts=20, nfft=512, sample_rate=16000, low_freq=0, high_freq=None,
                     scale="constant"):
    high_freq = high_freq or sample_rate / 2
    low_freq = low_freq or 0
    points = np.linspace(hz2bark(low_freq), hz2bark(high_freq), nfilts + 4)
    bins = np.floor(bark2fft(points)).astype(int)
    bank = np.zeros((nfilts, nfft // 2 + 1))
    cs = scale_factors(nfilts, scale)
    for i in range(nfilts):
        lo, hi = int(bins[i]), int(bins[i + 4])
        if hi > lo:
            cols = np.arange(lo, hi)
            bank[i, lo:hi] = cs[i] * skirt(fft2bark(cols), points[i + 2])
    return np.abs(bank)


def bark_spec(audio, sample_rate, window_size, hop_size, fft_size=512, num_filt=24):
    powers = power_spec(audio, (window_size, hop_size), fft_size)
    bank = bark_filterbanks(nfilts=num_filt, nfft=fft_size, sample_rate=sample_rate)
    return safe_log(np.dot(powers, bank.T))


def bfcc_spec(audio, sample_rate, window_size, hop_size, fft_size=512, num_filt=26, num_coeffs=13):
    powers = power_spec(audio, (window_size, hop_size), fft_size)
    if powers.size == 0:
        return np.empty((0, min(num_filt, num_coeffs)))
    bank = bark_filterbanks(nfilts=num_filt, nfft=fft_size, sample_rate=sample_rate)
    barks = safe_log(np.dot(powers, bank.T))
    out = dct(barks, norm='ortho')[:, :num_coeffs]
    out[:, 0] = safe_log(np.sum(powers, 1))
    return out
