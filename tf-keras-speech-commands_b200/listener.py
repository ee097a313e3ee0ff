"""Streaming feature update: the part of listen.py's Listener that is on the hot path
(update_vectors, listen.py:96-114; buffers listen.py:88-92), for one stream or a batch of streams.

State per stream lives on the GPU (carry < window samples + ring of n_features rows, double buffered); each push is
ONE launch of the fused MFCC kernel, which reads concat(carry, chunk), writes the new ring rows and carries the state
over (include/scfeat.h, scf_stream_*).
use_delta: the delta columns of the returned ring are computed on the device from the ring's base columns (the
reference re-applies add_deltas to the already widened ring every chunk, listen.py:111-112, which is a bug this does
not copy).  Every call returns a fresh array, as the reference does.
"""
import ctypes

import numpy as np

from . import _lib
from .data_utils import _mfcc_plan
from .params import pr


class FeatureStream:
    """n_streams concurrent listeners sharing one configuration (the current ``pr``)."""

    def __init__(self, n_streams=1, max_chunk=4096):
        self.use_delta = bool(pr.use_delta)
        self.plan = _mfcc_plan('diff' if self.use_delta else None)      # delta columns of the returned ring: on the device
        self.n_streams = int(n_streams)
        self.rows = int(pr.n_features)
        self.cols = self.plan.out_cols
        self.max_chunk = int(max_chunk)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().scf_stream_create(self.plan.handle, self.n_streams, self.rows, self.max_chunk,
                                                ctypes.byref(self._h)))
        self._ring = np.zeros((self.n_streams, self.rows, self.cols), dtype=np.float32)
        self._new = np.zeros((self.n_streams,), dtype=np.int32)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                _lib.lib().scf_stream_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def reset(self):
        _lib.check(_lib.lib().scf_stream_reset(self._h, None))

    def push(self, chunks, copy=True):
        """chunks: int16 [n_streams, chunk_len].  Returns (ring [n_streams, rows, cols] float32 oldest->newest,
        new_rows [n_streams] int32) -- fresh arrays, like the reference, unless copy=False (then views of buffers the
        next push overwrites)."""
        c = np.ascontiguousarray(chunks, dtype=np.int16)
        if c.ndim != 2 or c.shape[0] != self.n_streams:
            raise ValueError('chunks must be [n_streams, chunk_len]')
        _lib.check(_lib.lib().scf_stream_push_host_i16(self._h, c.ctypes.data, c.shape[1], self._ring.ctypes.data,
                                                       self._new.ctypes.data))
        if copy:
            return self._ring.copy(), self._new.copy()
        return self._ring, self._new

    def push_device(self, d_chunks, chunk_len, d_ring_out=None, d_new_rows=None, stream=0):
        """Device-pointer form (stream-ordered, no host sync)."""
        _lib.check(_lib.lib().scf_stream_push_i16(self._h, d_chunks, chunk_len, d_ring_out, d_new_rows, stream))


class Listener:
    """Single-stream object with the reference's method name and return shape."""

    def __init__(self, max_chunk=4096):
        self._fs = FeatureStream(1, max_chunk)

    def update_vectors(self, chunk):
        """chunk: raw 16-bit LE mono bytes -> (n_features, feature_size, 1) float32"""
        pcm = np.frombuffer(chunk, dtype='<i2')
        if len(pcm) == 0:
            ring = self._fs._ring.copy()
        else:
            ring, _ = self._fs.push(pcm[None, :])
        return np.expand_dims(ring[0], axis=-1)
