"""sonopy-shaped entry points (mfcc_spec / mel_spec / power_spec / filterbanks) on the B200.

Call shapes are fixed by the reference's call sites: common/data_utils.py:69,
tools/misc/plot_spectrogram.py:25,28, tools/audio_process/mfcc_feature.py:44.  Every numeric result
comes from libscfeat's CUDA kernels; results are float32 (the reference computes float64 and stores
float32, classifier/data.py:67).  There is no CPU fallback.

Input convention (one deliberate difference from sonopy): an **int16** array is taken as PCM and scaled by 1/32768 inside
the loader -- the result equals ``mfcc_spec(buffer_to_audio(pcm.tobytes()), ...)``, the only way the reference ever
feeds these functions (common/data_utils.py:21,69).  sonopy itself would compute on the raw integer values (features
offset by log(32768^2) in every log band); pass ``pcm.astype(np.float32)`` to get exactly that.  Every other dtype is
converted to float32 and used as is.
"""
from functools import lru_cache

import numpy as np

from . import _lib
from .plan import BANK_MEL_SONOPY, OUT_CEPSTRUM, OUT_LOG_BANK, OUT_POWER, PAD_NONE, get_plan


def safe_log(x):
    """Prevents error on log(0) or log(-1) (host helper, common/bark_feature.py:75-77)"""
    return np.log(np.clip(x, np.finfo(float).eps, None))


def chop_array(arr, window_size, hop_size):
    """chop_array([1,2,3], 2, 1) -> [[1,2], [2,3]]  (host helper, common/bark_feature.py:80-82)"""
    return [arr[i - window_size:i] for i in range(window_size, len(arr) + 1, hop_size)]


def _as_input(audio):
    a = np.asarray(audio)
    if a.ndim != 1:
        raise ValueError('audio must be 1-D')
    if a.dtype != np.int16:
        a = a.astype(np.float32, copy=False)
    return a


def _run(audio, out_kind, window_stride, fft_size, **kw):
    a = _as_input(audio)
    window, hop = int(window_stride[0]), int(window_stride[1])
    plan = get_plan(window=window, hop=hop, n_fft=int(fft_size), output=out_kind, **kw)
    if plan.frames(len(a)) == 0:
        return np.empty((0, plan.out_cols), dtype=np.float32)
    return plan.extract_host(a, pad=PAD_NONE)


def power_spec(audio, window_stride=(160, 80), fft_size=512):
    """Calculates power spectrogram: |rfft(frame, n=fft_size)|^2 / fft_size"""
    return _run(audio, OUT_POWER, window_stride, fft_size)


@lru_cache()
def filterbanks(sample_rate, num_filt, fft_len):
    """Triangular mel bank [num_filt, fft_len]; fft_len = n_fft/2+1 (sonopy's own convention)."""
    return _lib.build_bank(sample_rate=sample_rate, n_fft=2 * (fft_len - 1), n_filt=num_filt, bank=BANK_MEL_SONOPY)


def mel_spec(audio, sample_rate, window_stride=(160, 80), fft_size=512, num_filt=20):
    """Calculates mel spectrogram (condensed spectrogram)"""
    return _run(audio, OUT_LOG_BANK, window_stride, fft_size, sample_rate=int(sample_rate), n_filt=int(num_filt),
                bank=BANK_MEL_SONOPY)


def mfcc_spec(audio, sample_rate, window_stride=(160, 80), fft_size=512, num_filt=20, num_coeffs=13,
              return_parts=False):
    """Calculates mel frequency cepstrum coefficient spectrogram"""
    mfccs = _run(audio, OUT_CEPSTRUM, window_stride, fft_size, sample_rate=int(sample_rate), n_filt=int(num_filt),
                 n_coeffs=int(num_coeffs), bank=BANK_MEL_SONOPY)
    if return_parts:
        powers = power_spec(audio, window_stride, fft_size)
        filters = filterbanks(sample_rate, num_filt, int(fft_size) // 2 + 1)
        mels = mel_spec(audio, sample_rate, window_stride, fft_size, num_filt)
        return powers, filters, mels, mfccs
    return mfccs
