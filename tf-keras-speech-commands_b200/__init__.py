"""B200-native batched speech-feature extraction: drop-in for the feature hot path of
david8862/tf-keras-speech-commands (common/bark_feature.py, common/data_utils.py feature entry points,
sonopy.mfcc_spec call shape, listen.py update_vectors).  Python host code over the C ABI in
include/scfeat.h; all arithmetic runs in hand-written sm_100a CUDA kernels (csrc/).  No CPU fallback.

The directory name contains '-', so import it with importlib (or via the top-level alias ``scfeat``):
    import scfeat
    feats = scfeat.data_utils.extract_features_batch(pcm_int16)
"""
from . import _lib, bark_feature, cache, data_utils, dist, listener, params, plan, postprocess, sonopy  # noqa: F401
from ._lib import ScfError, build  # noqa: F401
from .plan import Plan, get_plan, launch_count, measure_fp32_flops  # noqa: F401

__all__ = ['bark_feature', 'data_utils', 'listener', 'params', 'plan', 'sonopy', 'Plan', 'get_plan', 'ScfError',
           'build', 'launch_count', 'measure_fp32_flops']
