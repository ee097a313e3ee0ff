"""float64 numpy restatement of listen.py's ThresholdDecoder (:452-521) and TriggerDetector (:525-559; C++ twins
inference/tflite/threshold_decoder.h:19-113, speech_commands.h:263-289), single-stream and batched.

TEST INFRASTRUCTURE (see oracle/__init__.py): the checker for the device implementation behind scf_post_* (csrc/
scfeat_post.cu, scfeat.postprocess).  Pinned against outputs of the reference's own classes, extracted verbatim from
listen.py by tests/golden/make_golden.py -> tests/golden/ref_postprocess.npz (tests/test_postprocess.py).
"""
import math

import numpy as np


class ThresholdDecoder:
    """Maps raw network outputs to a roughly linear confidence through the cumulative distribution of a mixture of
    logit-normal components (mu, std).  Same constructor, decode() and encode() as the reference."""

    def __init__(self, mu_stds, center=0.5, resolution=200, min_z=-4, max_z=4):
        mu_stds = [tuple(ms) for ms in mu_stds]
        self.min_out = int(min(mu + min_z * std for mu, std in mu_stds))
        self.max_out = int(max(mu + max_z * std for mu, std in mu_stds))
        self.out_range = self.max_out - self.min_out
        self.cd = np.cumsum(self._calc_pd(mu_stds, resolution))
        self.center = center

    @staticmethod
    def sigmoid(x):
        return 1 / (1 + math.exp(-x))

    @staticmethod
    def asigmoid(x):
        return -math.log(1 / x - 1) if (0 < x < 1) else -10

    @staticmethod
    def pdf(x, mu, std):
        if std == 0:
            return 0
        return (1.0 / (std * math.sqrt(2 * math.pi))) * np.exp(-(x - mu) ** 2 / (2 * std ** 2))

    def _calc_pd(self, mu_stds, resolution):
        points = np.linspace(self.min_out, self.max_out, resolution * self.out_range)
        return np.sum([self.pdf(points, mu, std) for mu, std in mu_stds], axis=0) / (resolution * len(mu_stds))

    def decode(self, raw_output):
        """scalar form, as in the reference"""
        return float(self.decode_batch(np.asarray([raw_output], dtype=np.float64))[0])

    def decode_batch(self, raw):
        """raw: array of network outputs in [0, 1] -> decoded confidences, element-wise identical to decode()."""
        raw = np.asarray(raw, dtype=np.float64)
        out = np.empty_like(raw)
        fixed = (raw == 1.0) | (raw == 0.0)
        inside = (raw > 0) & (raw < 1)
        logit = np.full(raw.shape, -10.0)
        with np.errstate(divide='ignore', over='ignore'):
            logit[inside] = -np.log(1 / raw[inside] - 1)
        if self.out_range == 0:
            cp = (raw > self.min_out).astype(np.float64)
        else:
            ratio = np.clip((logit - self.min_out) / self.out_range, 0.0, 1.0)
            cp = self.cd[(ratio * (len(self.cd) - 1) + 0.5).astype(np.int64)]
        low = cp < self.center
        out[low] = 0.5 * cp[low] / self.center
        out[~low] = 0.5 + 0.5 * (cp[~low] - self.center) / (1 - self.center)
        out[fixed] = raw[fixed]
        return out

    def encode(self, threshold):
        threshold = 0.5 * threshold / self.center
        if threshold < 0.5:
            cp = threshold * self.center * 2
        else:
            cp = (threshold - 0.5) * 2 * (1 - self.center) + self.center
        ratio = np.searchsorted(self.cd, cp) / len(self.cd)
        return self.sigmoid(self.min_out + self.out_range * ratio)


class TriggerDetector:
    """Single-stream detector with the reference's interface: update(index, score) -> bool."""

    def __init__(self, chunk_size, class_names, sensitivity=0.5, trigger_level=3):
        self._b = BatchTriggerDetector(1, chunk_size, class_names, sensitivity, trigger_level)

    @property
    def activation(self):
        return int(self._b.activation[0])

    def update(self, index, score):
        return bool(self._b.update(np.asarray([index]), np.asarray([score]))[0])


class BatchTriggerDetector:
    """N independent trigger state machines updated with one vector operation per step."""

    def __init__(self, n_streams, chunk_size, class_names, sensitivity=0.5, trigger_level=3):
        self.chunk_size = chunk_size
        self.is_background = np.asarray([c == 'background' for c in class_names], dtype=bool)
        self.sensitivity = sensitivity
        self.trigger_level = trigger_level
        self.activation = np.zeros(n_streams, dtype=np.int64)
        self.record_index = np.full(n_streams, -1, dtype=np.int64)          # -1 == None

    def update(self, index, score):
        """index: int [N] (arg-max class per stream), score: float [N] -> bool [N] "stream activated now"."""
        index = np.asarray(index, dtype=np.int64)
        score = np.asarray(score, dtype=np.float64)
        hot = (~self.is_background[index]) & (index == self.record_index) & (score > self.sensitivity)
        act = self.activation.copy()
        act[hot] += 1
        fired = hot & (act > self.trigger_level)
        act[fired] = -(8 * 2048) // self.chunk_size
        cool = ~hot
        act[cool & (self.activation < 0)] += 1
        act[cool & (self.activation > 0)] -= 1
        self.activation = act
        # the reference returns before recording the index when it fires (listen.py:545-547)
        self.record_index = np.where(fired, self.record_index, index)
        return fired
