"""Streaming post-processing on the device: listen.py's ThresholdDecoder (:452-521) and
TriggerDetector (:525-559; C++ twins inference/tflite/threshold_decoder.h:19-113, speech_commands.h:263-289) for a
batch of concurrent streams, behind the C ABI's scf_post_* calls (csrc/scfeat_post.cu).

One ``PostProcessor.step`` does what one iteration of the listen.py loop does after the model (listen.py:411-425) for
every stream in ONE kernel launch: arg-max / max of the class scores, decode of a non-background score, trigger update.
State (activation counters, recorded class) lives on the GPU.  The classes with the reference's names keep its
constructor and method signatures and run on the same kernels with one stream.  No CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


def build_cd(mu_stds, resolution=200, min_z=-4, max_z=4):
    """ThresholdDecoder.__init__'s table (host arithmetic inside libscfeat, float64): (min_out, max_out, cd)."""
    ms = np.ascontiguousarray(np.asarray(mu_stds, dtype=np.float64).reshape(-1, 2))
    lo, hi, n = ctypes.c_int32(), ctypes.c_int32(), ctypes.c_int64(0)
    L = _lib.lib()
    check(L.scf_post_build_cd(ms.ctypes.data, len(ms), resolution, float(min_z), float(max_z), ctypes.byref(lo),
                              ctypes.byref(hi), None, ctypes.byref(n)))
    cd = np.zeros(max(n.value, 0), dtype=np.float64)
    check(L.scf_post_build_cd(ms.ctypes.data, len(ms), resolution, float(min_z), float(max_z), ctypes.byref(lo),
                              ctypes.byref(hi), cd.ctypes.data if len(cd) else None, ctypes.byref(n)))
    return lo.value, hi.value, cd


class PostProcessor:
    """ThresholdDecoder + n_streams TriggerDetectors on the GPU."""

    def __init__(self, n_streams, class_names, threshold_config=((6, 4),), threshold_center=0.2, chunk_size=1024,
                 sensitivity=0.5, trigger_level=3, resolution=200, min_z=-4, max_z=4, device=-1):
        self.n_streams, self.class_names = int(n_streams), list(class_names)
        self.n_classes = len(self.class_names)
        ms = np.ascontiguousarray(np.asarray(threshold_config, dtype=np.float64).reshape(-1, 2))
        bg = np.ascontiguousarray([1 if c == 'background' else 0 for c in self.class_names], dtype=np.uint8)
        self._h = ctypes.c_void_p()
        check(_lib.lib().scf_post_create(ms.ctypes.data, len(ms), float(threshold_center), int(resolution), float(min_z),
                                         float(max_z), bg.ctypes.data, self.n_classes, self.n_streams, int(chunk_size),
                                         float(sensitivity), int(trigger_level), int(device), ctypes.byref(self._h)))
        self._dev = {}                  # name -> (device pointer, bytes)

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h:
            try:
                for p, _ in self._dev.values():
                    _lib.lib().scf_device_free(-1, p)
                _lib.lib().scf_post_destroy(h)
            except Exception:
                pass

    @property
    def handle(self):
        return self._h

    def _buf(self, name, nbytes):
        have = self._dev.get(name)
        if have is None or have[1] < nbytes:
            if have is not None:
                check(_lib.lib().scf_device_free(-1, have[0]))
            p = ctypes.c_void_p()
            check(_lib.lib().scf_device_malloc(-1, int(nbytes), ctypes.byref(p)))
            self._dev[name] = (p.value, int(nbytes))
        return self._dev[name][0]

    def reset(self, stream=None):
        check(_lib.lib().scf_post_reset(self._h, stream))

    # ---- device-pointer forms (stream-ordered, no host synchronisation) -----------------------------------------
    def step_device(self, d_probs, d_index=None, d_score=None, d_fired=None, stream=None):
        """d_probs: float32 [n_streams, n_classes] on the device; outputs int32 / float64 / uint8 [n_streams] or None."""
        check(_lib.lib().scf_post_step(self._h, d_probs, d_index, d_score, d_fired, stream))

    def decode_device(self, d_raw, n, d_out, stream=None):
        check(_lib.lib().scf_post_decode(self._h, d_raw, int(n), d_out, stream))

    # ---- host-array conveniences (copies through scf_memcpy) ----------------------------------------------------
    def step(self, probs):
        """probs: [n_streams, n_classes] model outputs -> (index int32 [n], score float64 [n], fired bool [n])."""
        pr = np.ascontiguousarray(probs, dtype=np.float32)
        if pr.shape != (self.n_streams, self.n_classes):
            raise ValueError('probs must be [n_streams, n_classes]')
        L, n = _lib.lib(), self.n_streams
        d_p, d_i = self._buf('probs', pr.nbytes), self._buf('index', 4 * n)
        d_s, d_f = self._buf('score', 8 * n), self._buf('fired', n)
        check(L.scf_memcpy(-1, d_p, pr.ctypes.data, pr.nbytes, 0, None))
        self.step_device(d_p, d_i, d_s, d_f)
        idx, score, fired = np.empty(n, np.int32), np.empty(n, np.float64), np.empty(n, np.uint8)
        check(L.scf_memcpy(-1, idx.ctypes.data, d_i, idx.nbytes, 1, None))
        check(L.scf_memcpy(-1, score.ctypes.data, d_s, score.nbytes, 1, None))
        check(L.scf_memcpy(-1, fired.ctypes.data, d_f, fired.nbytes, 1, None))
        return idx, score, fired.astype(bool)

    def decode(self, raw):
        """Element-wise ThresholdDecoder.decode of float64 raw outputs (any shape)."""
        r = np.ascontiguousarray(raw, dtype=np.float64)
        if r.size == 0:
            return r.copy()
        L = _lib.lib()
        d_in, d_out = self._buf('raw', r.nbytes), self._buf('dec', r.nbytes)
        check(L.scf_memcpy(-1, d_in, r.ctypes.data, r.nbytes, 0, None))
        self.decode_device(d_in, r.size, d_out)
        out = np.empty_like(r)
        check(L.scf_memcpy(-1, out.ctypes.data, d_out, r.nbytes, 1, None))
        return out

    def state(self):
        """(activation int32 [n], record_index int32 [n], -1 = None)"""
        a, r = np.empty(self.n_streams, np.int32), np.empty(self.n_streams, np.int32)
        check(_lib.lib().scf_post_state(self._h, a.ctypes.data, r.ctypes.data, None))
        return a, r


class ThresholdDecoder:
    """Same constructor and decode() as the reference class (listen.py:452-509); the arithmetic runs on the GPU.
    encode() (listen.py:511-517, used once at start-up to turn a threshold into a raw value) needs only the table and
    stays a host-side scalar function, as in the reference."""

    def __init__(self, mu_stds, center=0.5, resolution=200, min_z=-4, max_z=4):
        mu_stds = [tuple(ms) for ms in mu_stds]
        self._post = PostProcessor(1, ['background', 'x'], mu_stds, center, resolution=resolution, min_z=min_z, max_z=max_z)
        self.min_out, self.max_out, self.cd = build_cd(mu_stds, resolution, min_z, max_z)
        self.out_range = self.max_out - self.min_out
        self.center = center

    def decode(self, raw_output):
        return float(self._post.decode(np.asarray([raw_output], dtype=np.float64))[0])

    def decode_batch(self, raw):
        return self._post.decode(raw)

    def encode(self, threshold):
        threshold = 0.5 * threshold / self.center
        if threshold < 0.5:
            cp = threshold * self.center * 2
        else:
            cp = (threshold - 0.5) * 2 * (1 - self.center) + self.center
        ratio = np.searchsorted(self.cd, cp) / len(self.cd)
        return 1 / (1 + np.exp(-(self.min_out + self.out_range * ratio)))


class TriggerDetector:
    """Same constructor and update(index, score) as the reference class (listen.py:525-559), one stream on the GPU."""

    def __init__(self, chunk_size, class_names, sensitivity=0.5, trigger_level=3):
        self._b = BatchTriggerDetector(1, chunk_size, class_names, sensitivity, trigger_level)

    @property
    def activation(self):
        return int(self._b.activation[0])

    def update(self, index, score):
        return bool(self._b.update(np.asarray([index]), np.asarray([score]))[0])


class BatchTriggerDetector:
    """n_streams trigger state machines on the GPU; update(index [n], score [n]) -> fired bool [n] (scores are taken
    as given, i.e. already decoded: scf_post_trigger_update)."""

    def __init__(self, n_streams, chunk_size, class_names, sensitivity=0.5, trigger_level=3):
        self._post = PostProcessor(n_streams, class_names, chunk_size=chunk_size, sensitivity=sensitivity,
                                   trigger_level=trigger_level)
        self.n = int(n_streams)

    @property
    def activation(self):
        return self._post.state()[0]

    def update(self, index, score):
        idx = np.ascontiguousarray(index, dtype=np.int32)
        sc = np.ascontiguousarray(score, dtype=np.float64)
        if idx.shape != (self.n,) or sc.shape != (self.n,):
            raise ValueError('index and score must be [n_streams]')
        L, post = _lib.lib(), self._post
        d_i, d_s, d_f = post._buf('index', 4 * self.n), post._buf('score', 8 * self.n), post._buf('fired', self.n)
        check(L.scf_memcpy(-1, d_i, idx.ctypes.data, idx.nbytes, 0, None))
        check(L.scf_memcpy(-1, d_s, sc.ctypes.data, sc.nbytes, 0, None))
        check(L.scf_post_trigger_update(post.handle, d_i, d_s, d_f, None))
        fired = np.empty(self.n, np.uint8)
        check(L.scf_memcpy(-1, fired.ctypes.data, d_f, fired.nbytes, 1, None))
        return fired.astype(bool)
