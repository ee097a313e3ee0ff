"""Plan objects: one immutable libscfeat plan per distinct feature configuration.

The reference reads its configuration from the mutable module-global ``pr`` at call time
(classifier/params.py:99-115, common/data_utils.py:69), so plans are cached BY VALUE of the
configuration, never captured at import.
"""
import ctypes
import threading

import numpy as np

from . import _lib
from ._lib import (BANK_BARK_REF, BANK_CUSTOM, BANK_MEL_SONOPY, OUT_CEPSTRUM, OUT_LOG_BANK, OUT_POWER,  # noqa: F401
                   PAD_FRONT_ZERO, PAD_NONE, ScfError, check)

_cache = {}
_cache_lock = threading.Lock()


class Plan:
    """Owns one ``scf_plan*``.  Thread-safe to share (the C plan is immutable)."""

    def __init__(self, **kw):
        self.cfg, self._keep = _lib.make_config(**kw)
        self._h = ctypes.c_void_p()
        check(_lib.lib().scf_plan_create(ctypes.byref(self.cfg), ctypes.byref(self._h)))
        self.out_cols = int(_lib.lib().scf_out_cols(ctypes.byref(self.cfg)))
        self.window, self.hop, self.n_fft = self.cfg.window, self.cfg.hop, self.cfg.n_fft

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                _lib.lib().scf_plan_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def frames(self, n_samples):
        """scf_num_frames (chop_array's frame count) without the ctypes round trip."""
        n = int(n_samples)
        return 0 if n < self.window else (n - self.window) // self.hop + 1

    # ---- host numpy in -> host numpy out (what the drop-in functions use) ----------------------
    def extract_host(self, clips, lengths=None, pad=PAD_FRONT_ZERO, out=None):
        """clips: [n, L] (or [L]) int16 PCM or float32 audio.  Returns float32 [n, frames(L), cols].
        `out` may be a preallocated C-contiguous float32 array of that shape (e.g. pinned memory) to write into."""
        a = np.asarray(clips)
        single = a.ndim == 1
        if single:
            a = a[None, :]
        if a.ndim != 2:
            raise ValueError('clips must be 1-D or 2-D')
        if a.dtype == np.int16:
            fn = _lib.lib().scf_extract_host_i16
        else:
            a = a.astype(np.float32, copy=False)
            fn = _lib.lib().scf_extract_host_f32
        a = np.ascontiguousarray(a)
        n, L = a.shape
        shape = (n, self.frames(L), self.out_cols)
        if out is None:
            out = np.empty(shape, dtype=np.float32)
        elif out.shape != shape or out.dtype != np.float32 or not out.flags['C_CONTIGUOUS']:
            raise ValueError('out must be a C-contiguous float32 array of shape %r' % (shape,))
        lp = None
        if lengths is not None:
            lengths = np.ascontiguousarray(lengths, dtype=np.int32)
            if lengths.shape != (n,):
                raise ValueError('lengths must have one entry per clip')
            lp = lengths.ctypes.data
        if out.size:
            check(fn(self._h, a.ctypes.data, n, L, L, lp, pad, out.ctypes.data))
        return out[0] if single else out

    def extract_host_async(self, clips, out, lengths=None, pad=PAD_FRONT_ZERO):
        """Enqueue one batch (int16 [n, L], C-contiguous, ideally pinned) and return immediately; `out` (float32
        [n, frames(L), cols], C-contiguous, ideally pinned) is valid after host_sync().  Two staging slots alternate
        inside the library, so the upload of the next batch overlaps the kernel and download of this one.  The
        caller keeps `clips`, `lengths` and `out` alive until host_sync()."""
        a = np.asarray(clips)
        if a.ndim != 2 or a.dtype != np.int16 or not a.flags['C_CONTIGUOUS']:
            raise ValueError('clips must be a C-contiguous int16 [n, L] array')
        n, L = a.shape
        shape = (n, self.frames(L), self.out_cols)
        if out.shape != shape or out.dtype != np.float32 or not out.flags['C_CONTIGUOUS']:
            raise ValueError('out must be a C-contiguous float32 array of shape %r' % (shape,))
        lp = None
        if lengths is not None:
            if lengths.dtype != np.int32 or lengths.shape != (n,) or not lengths.flags['C_CONTIGUOUS']:
                raise ValueError('lengths must be a C-contiguous int32 [n] array (kept alive by the caller)')
            lp = lengths.ctypes.data
        check(_lib.lib().scf_extract_host_i16_async(self._h, a.ctypes.data, n, L, L, lp, pad, out.ctypes.data))

    def host_sync(self):
        """Wait for every extract_host_async() issued on this plan."""
        check(_lib.lib().scf_host_sync(self._h))

    # ---- raw device pointers (torch / cupy / DLPack producers hand in .data_ptr()) -------------
    def extract_device(self, d_in, n_clips, clip_len, d_out, clip_stride=None, d_lengths=None, pad=PAD_FRONT_ZERO,
                       stream=0, is_f32=False):
        fn = _lib.lib().scf_extract_f32 if is_f32 else _lib.lib().scf_extract_i16
        check(fn(self._h, d_in, n_clips, clip_len if clip_stride is None else clip_stride, clip_len,
                 d_lengths, pad, d_out, stream))

    def extract_dlpack(self, d_in, n_clips, clip_len, clip_stride=None, d_lengths=None, pad=PAD_FRONT_ZERO, stream=0):
        """Runs the extraction into a library-owned device buffer and returns a PyCapsule named
        "dltensor" (consume with tf.experimental.dlpack.from_dlpack / torch.from_dlpack)."""
        dl = ctypes.c_void_p()
        check(_lib.lib().scf_extract_i16_dlpack(self._h, d_in, n_clips, clip_len if clip_stride is None else clip_stride,
                                                clip_len, d_lengths, pad, ctypes.byref(dl), stream))
        return _make_capsule(dl.value)


# -- DLPack capsule plumbing: the capsule (and its destructor) is made in C, see scf_dlpack_make_capsule ----
_pydll = None


def _make_capsule(ptr):
    global _pydll
    if _pydll is None:
        _pydll = ctypes.PyDLL(_lib.LIB_PATH)          # PyDLL: keeps the GIL while PyCapsule_New runs
        _pydll.scf_dlpack_make_capsule.restype = ctypes.py_object
        _pydll.scf_dlpack_make_capsule.argtypes = [ctypes.c_void_p]
    return _pydll.scf_dlpack_make_capsule(ptr)


def dlpack_alloc(shape, device=-1):
    """Library-owned float32 device buffer of `shape` (1..4 dims): returns (device pointer, finish) where finish() turns
    it into the PyCapsule "dltensor" whose deleter frees it.  Fill the buffer through the device entry points first."""
    shp = (ctypes.c_int64 * len(shape))(*[int(v) for v in shape])
    dl, ptr = ctypes.c_void_p(), ctypes.c_void_p()
    check(_lib.lib().scf_dlpack_alloc(int(device), shp, len(shape), ctypes.byref(dl), ctypes.byref(ptr)))
    return ptr.value, (lambda: _make_capsule(dl.value))


_RELEASE_T = ctypes.CFUNCTYPE(None, ctypes.c_void_p)
_wrapped = {}                      # token -> (owner object, callback): keeps both alive until the consumer lets go
_wrapped_lock = threading.Lock()
_wrapped_next = [1]
_released = []                     # entries whose consumer let go: dropped at the next call, never inside their own callback


def dlpack_wrap(owner, d_ptr, shape, device=-1):
    """PyCapsule "dltensor" over device memory that `owner` (any Python object) keeps alive: the object is referenced
    until the consumer's deleter runs."""
    with _wrapped_lock:
        token = _wrapped_next[0]
        _wrapped_next[0] += 1
        del _released[:]

    def _release(_ctx, token=token):
        with _wrapped_lock:
            _released.append(_wrapped.pop(token, None))

    cb = _RELEASE_T(_release)
    with _wrapped_lock:
        _wrapped[token] = (owner, cb)
    shp = (ctypes.c_int64 * len(shape))(*[int(v) for v in shape])
    dl = ctypes.c_void_p()
    try:
        check(_lib.lib().scf_dlpack_wrap(d_ptr, int(device), shp, len(shape), ctypes.cast(cb, ctypes.c_void_p), None,
                                         ctypes.byref(dl)))
    except Exception:
        _release(None)
        raise
    return _make_capsule(dl.value)


def get_plan(**kw):
    """Plan cached by the VALUE of its configuration (custom banks are never cached)."""
    if kw.get('custom_bank') is not None:
        return Plan(**kw)
    key = tuple(sorted(kw.items()))
    with _cache_lock:
        p = _cache.get(key)
        if p is None:
            p = _cache[key] = Plan(**kw)
        return p


def clear_plan_cache():
    with _cache_lock:
        _cache.clear()


def launch_count():
    return int(_lib.lib().scf_launch_count())


def measure_fp32_flops(device=-1):
    v = ctypes.c_double()
    check(_lib.lib().scf_measure_fp32_flops(device, ctypes.byref(v)))
    return v.value
