"""float64 numpy restatement of the MFCC arithmetic the reference delegates to ``sonopy``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  ``sonopy`` (MycroftAI, PyPI, unpinned in
/root/reference/requirements.txt:7) is not vendored under /root/reference and is not
installable offline, so this file restates its published algorithm.  Anchors inside the
reference:

  * call site                     common/data_utils.py:69 (vectorize_raw)
  * chop_array/power_spec/safe_log  verbatim copies at common/bark_feature.py:75-89
  * mfcc_spec structure           common/bark_feature.py:156-175 (bfcc_spec = same with Bark bank)
  * mel scale + triangular bank   inference/tflite/mfcc.h:134-145, 230-264
  * low = 0 Hz, high = sample_rate  inference/tflite/speech_commands.h:304-307
  * DCT-II ortho                  inference/tflite/mfcc.h:42-71 (== scipy.fftpack.dct(norm='ortho'))
  * c0 := log frame energy        inference/tflite/mfcc.h:359, bark_feature.py:173
"""
from functools import lru_cache

import numpy as np
from scipy.fftpack import dct

EPS = np.finfo(float).eps  # 2.220446049250313e-16, bark_feature.py:77 / mfcc.h:18


def safe_log(x):
    """log with the argument floored at eps (bark_feature.py:75-77)."""
    return np.log(np.clip(x, EPS, None))


def n_frames(n_samples, window, hop):
    """Number of frames chop_array yields: range(window, n+1, hop) (bark_feature.py:80-82)."""
    if n_samples < window:
        return 0
    return (n_samples - window) // hop + 1


def frames_of(audio, window, hop):
    """[k, window] matrix of the rectangular, un-padded, tail-dropping frames."""
    audio = np.asarray(audio)
    k = n_frames(len(audio), window, hop)
    if k == 0:
        return np.empty((0, window))
    idx = np.arange(window)[None, :] + hop * np.arange(k)[:, None]
    return audio[idx]


def power_spec(audio, window_stride=(160, 80), fft_size=512):
    """|rfft(frame, n=fft_size)|^2 / fft_size  (bark_feature.py:85-89).

    rfft(n=fft_size) crops frames longer than fft_size and zero-pads shorter ones.
    """
    frames = frames_of(audio, *window_stride)
    spec = np.fft.rfft(frames, n=fft_size)
    return (spec.real ** 2 + spec.imag ** 2) / fft_size


def hertz_to_mels(f):
    return 1127.0 * np.log(1.0 + f / 700.0)      # mfcc.h:134-138


def mels_to_hertz(m):
    return 700.0 * (np.exp(m / 1127.0) - 1.0)    # mfcc.h:141-145


def mel_grid(sample_rate, num_filt, fft_len):
    """Filter edge indices: mel-uniform between 0 Hz and *sample_rate* (not Nyquist), mapped
    with int(hz * fft_len / sample_rate) where fft_len = n_fft/2+1 (mfcc.h:230-249 as called
    from speech_commands.h:304-307).

    Repeated indices are kept (the filters between them are empty and give log(eps)).  sonopy has a
    ``correct_grid`` helper meant to push duplicates forward, but as published it is a no-op: it is
    handed an ndarray, so ``[x[0] - 1] + x`` broadcasts to ``x - 1`` and its offset never leaves 0.
    The reference's own C++ twin has no such step either; tests/golden/ref_mfcc_cpp.npz pins a
    duplicate-grid case (n_fft 512, 40 filters) against it."""
    mels = np.linspace(hertz_to_mels(0.0), hertz_to_mels(float(sample_rate)), num_filt + 2, True)
    hz = mels_to_hertz(mels)
    idx = (hz * fft_len / sample_rate).astype(int)
    return [int(i) for i in idx]


@lru_cache(maxsize=None)
def filterbanks(sample_rate, num_filt, fft_len):
    """[num_filt, fft_len] triangular bank (mfcc.h:251-261): rising edge j-left / (mid-left)
    on [left, mid), falling edge (right-j)/(right-mid) on [mid, right)."""
    grid = mel_grid(sample_rate, num_filt, fft_len)
    banks = np.zeros((num_filt, fft_len))
    for i in range(num_filt):
        left, mid, right = grid[i], grid[i + 1], grid[i + 2]
        banks[i, left:mid] = np.linspace(0.0, 1.0, mid - left, False)
        banks[i, mid:right] = np.linspace(1.0, 0.0, right - mid, False)
    return banks


def mel_spec(audio, sample_rate, window_stride=(160, 80), fft_size=512, num_filt=20):
    spec = power_spec(audio, window_stride, fft_size)
    return safe_log(np.dot(spec, filterbanks(sample_rate, num_filt, spec.shape[1]).T))


def mfcc_spec(audio, sample_rate, window_stride=(160, 80), fft_size=512, num_filt=20,
              num_coeffs=13, return_parts=False):
    powers = power_spec(audio, window_stride, fft_size)
    if powers.size == 0:
        return np.empty((0, min(num_filt, num_coeffs)))
    filters = filterbanks(sample_rate, num_filt, powers.shape[1])
    mels = safe_log(np.dot(powers, filters.T))
    mfccs = dct(mels, norm='ortho')[:, :num_coeffs]
    mfccs[:, 0] = safe_log(np.sum(powers, 1))
    if return_parts:
        return powers, filters, mels, mfccs
    return mfccs


def dct2_ortho_matrix(n, n_out):
    """Explicit DCT-II ortho matrix [n, n_out]; y = x @ M equals dct(x, norm='ortho')[:, :n_out]
    (mfcc.h:56-66)."""
    k = np.arange(min(n, n_out))[None, :]
    i = np.arange(n)[:, None]
    m = np.sqrt(2.0 / n) * np.cos(np.pi * (i + 0.5) * k / n)
    m[:, 0] *= np.sqrt(0.5)
    return m
