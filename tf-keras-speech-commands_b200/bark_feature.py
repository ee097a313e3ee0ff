"""Drop-in for common/bark_feature.py: same public names and signatures, computed on the B200.

power_spec :85-89, bark_filterbanks :92-136, bark_spec :139-153, bfcc_spec :156-175 and the scale
helpers :16-72.  The scale helpers and Fm are host-side scalar functions (as in the reference); every
per-frame quantity comes from libscfeat's kernels.  Results are float32.

Input convention: an int16 array is taken as PCM and scaled by 1/32768 inside the loader (the reference's callers
always convert first, common/data_utils.py:21); the reference functions themselves would compute on the raw integer
values -- pass ``pcm.astype(np.float32)`` for that.  Other dtypes are converted to float32 and used as is.
"""
from functools import lru_cache

import numpy as np

from . import _lib
from .plan import BANK_BARK_REF, OUT_CEPSTRUM, OUT_LOG_BANK, PAD_NONE, get_plan
from .sonopy import chop_array, power_spec, safe_log  # noqa: F401  (same names as the reference module)


def hz2bark_1961(Hz):
    return 13.0 * np.arctan(0.00076 * Hz) + 3.5 * np.arctan((Hz / 7500.0) ** 2)


def hz2bark_1990(Hz):
    return (26.81 * Hz) / (1960 + Hz) - 0.5


def hz2bark_1992(Hz):
    return 6 * np.arcsinh(Hz / 600)


def hz2bark(f):
    """Hz -> Bark (Wang, Sekey & Gersho, 1992)"""
    return 6. * np.arcsinh(f / 600.)


def bark2hz(fb):
    return 600. * np.sinh(fb / 6.)


def fft2hz(fft, sample_rate=16000, nfft=512):
    return (fft * sample_rate) / (nfft + 1)


def hz2fft(fb, sample_rate=16000, nfft=512):
    return (nfft + 1) * fb / sample_rate


def fft2bark(fft, sample_rate=16000, nfft=512):
    return hz2bark((fft * sample_rate) / (nfft + 1))


def bark2fft(fb, sample_rate=16000, nfft=512):
    return (nfft + 1) * bark2hz(fb) / sample_rate


def Fm(fb, fc):
    """Bark critical-band filter amplitude at fb for centre fc (both in Bark)"""
    if fc - 2.5 <= fb <= fc - 0.5:
        return 10 ** (2.5 * (fb - fc + 0.5))
    elif fc - 0.5 < fb < fc + 0.5:
        return 1
    elif fc + 0.5 <= fb <= fc + 1.3:
        return 10 ** (-2.5 * (fb - fc - 0.5))
    else:
        return 0


@lru_cache()
def bark_filterbanks(nfilts=20, nfft=512, sample_rate=16000, low_freq=0, high_freq=None, scale="constant"):
    """Bark filterbank [nfilts, nfft/2+1] as libscfeat builds it (float64, host); same arguments as
    common/bark_feature.py:93 -- `low_freq or 0`, `high_freq or sample_rate / 2` (:104-105).  Bins past the last
    column (high_freq above the 8 kHz the reference's fixed bin mapping covers) are dropped where the reference
    raises IndexError."""
    return _lib.build_bank(sample_rate=sample_rate, n_fft=nfft, n_filt=nfilts, bank=BANK_BARK_REF, bank_scale=scale,
                           bank_low_hz=float(low_freq or 0), bank_high_hz=float(high_freq or 0))


def _run(audio, out_kind, sample_rate, window_size, hop_size, fft_size, **kw):
    a = np.asarray(audio)
    if a.dtype != np.int16:
        a = a.astype(np.float32, copy=False)
    plan = get_plan(sample_rate=int(sample_rate), window=int(window_size), hop=int(hop_size), n_fft=int(fft_size),
                    bank=BANK_BARK_REF, output=out_kind, **kw)
    if plan.frames(len(a)) == 0:
        return np.empty((0, plan.out_cols), dtype=np.float32)
    return plan.extract_host(a, pad=PAD_NONE)


def bark_spec(audio, sample_rate, window_size, hop_size, fft_size=512, num_filt=24):
    return _run(audio, OUT_LOG_BANK, sample_rate, window_size, hop_size, fft_size, n_filt=int(num_filt))


def bfcc_spec(audio, sample_rate, window_size, hop_size, fft_size=512, num_filt=26, num_coeffs=13):
    return _run(audio, OUT_CEPSTRUM, sample_rate, window_size, hop_size, fft_size, n_filt=int(num_filt),
                n_coeffs=int(num_coeffs))
