"""Audio-pipeline parameters: mirror of classifier/params.py (ListenerParams :16-91, pr :99-103,
inject_params :107-115, save_params :118-121) so that callers written against the reference keep working.

``pr`` is a process-wide singleton that callers mutate through ``inject_params`` AFTER import; every
feature call therefore reads it at call time and looks its plan up by value (plan.get_plan).
"""
import json
import os
from math import floor


class ListenerParams:
    """Same fields and derived properties as the reference's frozen attrs class.  Assignment raises
    (frozen) but ``__dict__.update`` works, which is how inject_params overrides values."""

    _FIELDS = ('buffer_t', 'window_t', 'hop_t', 'sample_rate', 'sample_depth', 'n_fft', 'n_filt', 'n_mfcc',
               'use_delta', 'threshold_config', 'threshold_center')

    def __init__(self, buffer_t, window_t, hop_t, sample_rate, sample_depth, n_fft, n_filt, n_mfcc, use_delta,
                 threshold_config, threshold_center):
        self.__dict__.update(buffer_t=buffer_t, window_t=window_t, hop_t=hop_t, sample_rate=sample_rate,
                             sample_depth=sample_depth, n_fft=n_fft, n_filt=n_filt, n_mfcc=n_mfcc,
                             use_delta=use_delta, threshold_config=threshold_config,
                             threshold_center=threshold_center)

    def __setattr__(self, name, value):
        raise AttributeError('ListenerParams is frozen; use inject_params()')

    def __repr__(self):
        return 'ListenerParams(%s)' % ', '.join('%s=%r' % (k, self.__dict__[k]) for k in self._FIELDS)

    @property
    def buffer_samples(self):
        """buffer_t converted to samples, truncating partial frames (params.py:59-63)"""
        samples = int(self.sample_rate * self.buffer_t + 0.5)
        return self.hop_samples * (samples // self.hop_samples)

    @property
    def n_features(self):
        """Number of timesteps in one input to the network (params.py:65-68)"""
        return 1 + int(floor((self.buffer_samples - self.window_samples) / self.hop_samples))

    @property
    def window_samples(self):
        return int(self.sample_rate * self.window_t + 0.5)

    @property
    def hop_samples(self):
        return int(self.sample_rate * self.hop_t + 0.5)

    @property
    def max_samples(self):
        return int(self.buffer_t * self.sample_rate)

    @property
    def feature_size(self):
        return self.n_mfcc * (2 if self.use_delta else 1)


# configs/params.json == classifier/params.py:99-103
pr = ListenerParams(
    buffer_t=1.0, window_t=0.064, hop_t=0.032, sample_rate=16000,
    sample_depth=2, n_fft=1024, n_filt=20, n_mfcc=20, use_delta=False,
    threshold_config=((6, 4),), threshold_center=0.2
)


def inject_params(params_file):
    """Set the global listener params from a saved JSON (params.py:107-115)"""
    try:
        with open(params_file) as f:
            pr.__dict__.update(**json.load(f))
    except (OSError, ValueError, TypeError):
        if os.path.isfile(params_file):
            print('Warning: Failed to load parameters from ' + params_file)
    return pr


def save_params(params_file):
    with open(params_file, 'w') as f:
        json.dump(pr.__dict__, f, indent=2)
