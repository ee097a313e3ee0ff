"""ctypes binding of libscfeat.so (C ABI: include/scfeat.h).  No torch, no pybind.

The library is built in-tree (``csrc/Makefile`` -> ``libscfeat.so`` next to this file) so that it
travels with the repository snapshot.  There is no CPU fallback: if the shared library is missing or
no sm_100 GPU is present, the calls raise.
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SCFEAT_LIB lets the kernel-tuning scripts load an experimental build; the product default is the in-tree library
LIB_PATH = os.environ.get('SCFEAT_LIB') or os.path.join(_HERE, 'libscfeat.so')
CSRC = os.path.join(_HERE, 'csrc')

# enums of include/scfeat.h
BANK_MEL_SONOPY, BANK_BARK_REF, BANK_CUSTOM = 0, 1, 2
SCALE = {'constant': 0, 'ascendant': 1, 'descendant': 2}
OUT_POWER, OUT_LOG_BANK, OUT_CEPSTRUM = 0, 1, 2
WIN = {'rect': 0, 'hamming': 1, 'hann': 2}
PAD_FRONT_ZERO, PAD_NONE = 0, 1
DELTA = {None: 0, 'none': 0, 'diff': 1, 'central': 2, 'central2': 3}

EXPORTS = [
    'scf_config_default', 'scf_num_frames', 'scf_out_cols', 'scf_build_bank', 'scf_build_dct',
    'scf_bank_apply_tasks',
    'scf_plan_create', 'scf_plan_destroy', 'scf_plan_config',
    'scf_extract_i16', 'scf_extract_f32', 'scf_extract_host_i16', 'scf_extract_host_f32',
    'scf_extract_host_i16_async', 'scf_host_sync',
    'scf_extract_i16_dlpack', 'scf_dlpack_make_capsule', 'scf_extract_i16_gather', 'scf_extract_i16_gather_multicast', 'scf_allgather_nccl',
    'scf_stream_create', 'scf_stream_destroy', 'scf_stream_reset', 'scf_stream_push_i16',
    'scf_stream_push_host_i16', 'scf_last_error', 'scf_version', 'scf_launch_count',
    'scf_measure_fp32_flops', 'scf_parallel_memcpy', 'scf_device_malloc', 'scf_device_free', 'scf_memcpy', 'scf_ipc_export',
    'scf_ipc_import', 'scf_ipc_close',
    'scf_post_build_cd', 'scf_post_create', 'scf_post_destroy', 'scf_post_reset', 'scf_post_decode', 'scf_post_step',
    'scf_post_trigger_update', 'scf_post_state', 'scf_post_info',
    'scf_wav_read_batch', 'scf_ingest_wavs', 'scf_ingest_wavs_device', 'scf_dlpack_alloc', 'scf_dlpack_wrap',
]


class ScfError(RuntimeError):
    """A libscfeat call failed (carries the status code and scf_last_error())."""

    def __init__(self, code, msg):
        super().__init__('libscfeat error %d: %s' % (code, msg))
        self.code = code


class Config(ctypes.Structure):
    _fields_ = [
        ('sample_rate', ctypes.c_int32), ('window', ctypes.c_int32), ('hop', ctypes.c_int32),
        ('n_fft', ctypes.c_int32), ('n_filt', ctypes.c_int32), ('n_coeffs', ctypes.c_int32),
        ('bank', ctypes.c_int32), ('bank_scale', ctypes.c_int32), ('output', ctypes.c_int32),
        ('window_fn', ctypes.c_int32), ('preemph_alpha', ctypes.c_float), ('pcm_scale', ctypes.c_float),
        ('device', ctypes.c_int32), ('delta', ctypes.c_int32), ('custom_bank', ctypes.c_void_p),
        ('bank_low_hz', ctypes.c_float), ('bank_high_hz', ctypes.c_float),
    ]


_lib = None
_lock = threading.Lock()


def build(force=False, verbose=False):
    """Compiles libscfeat.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    if force:
        subprocess.run(['make', '-C', CSRC, 'clean'], check=True, capture_output=not verbose)
    r = subprocess.run(['make', '-C', CSRC], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError('building libscfeat.so failed:\n' + r.stdout + r.stderr)
    if verbose:
        print(r.stdout)
    return LIB_PATH


def lib():
    """The loaded library (ctypes.CDLL) with argument types declared."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise ScfError(-3, 'libscfeat.so not built (run `make -C %s` or __graft_entry__.build()); '
                               'there is no CPU fallback' % CSRC)
        L = ctypes.CDLL(LIB_PATH)
        vp, i32, i64, f32p = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_void_p
        cfgp = ctypes.POINTER(Config)
        L.scf_config_default.argtypes = [cfgp]
        L.scf_num_frames.argtypes = [i64, i32, i32]
        L.scf_num_frames.restype = i64
        L.scf_out_cols.argtypes = [cfgp]
        L.scf_out_cols.restype = i32
        L.scf_build_bank.argtypes = [cfgp, vp]
        L.scf_build_dct.argtypes = [i32, i32, vp]
        L.scf_bank_apply_tasks.argtypes = [cfgp, vp, vp, vp, vp, vp]
        L.scf_plan_create.argtypes = [cfgp, ctypes.POINTER(vp)]
        L.scf_plan_destroy.argtypes = [vp]
        L.scf_plan_destroy.restype = None
        L.scf_plan_config.argtypes = [vp, cfgp]
        dev_args = [vp, vp, i64, i64, i32, vp, i32, f32p, vp]
        L.scf_extract_i16.argtypes = dev_args
        L.scf_extract_f32.argtypes = dev_args
        host_args = [vp, vp, i64, i64, i32, vp, i32, f32p]
        L.scf_extract_host_i16.argtypes = host_args
        L.scf_extract_host_f32.argtypes = host_args
        L.scf_extract_host_i16_async.argtypes = host_args
        L.scf_host_sync.argtypes = [vp]
        L.scf_extract_i16_dlpack.argtypes = [vp, vp, i64, i64, i32, vp, i32, ctypes.POINTER(vp), vp]
        L.scf_extract_i16_gather.argtypes = [vp, vp, i64, i64, i32, ctypes.POINTER(vp), i32, i32, i64, vp]
        L.scf_extract_i16_gather_multicast.argtypes = [vp, vp, i64, i64, i32, vp, i32, i32, i64, vp]
        L.scf_allgather_nccl.argtypes = [vp, vp, i64, vp, vp]
        L.scf_stream_create.argtypes = [vp, i32, i32, i32, ctypes.POINTER(vp)]
        L.scf_stream_destroy.argtypes = [vp]
        L.scf_stream_destroy.restype = None
        L.scf_stream_reset.argtypes = [vp, vp]
        L.scf_stream_push_i16.argtypes = [vp, vp, i32, vp, vp, vp]
        L.scf_stream_push_host_i16.argtypes = [vp, vp, i32, vp, vp]
        L.scf_last_error.restype = ctypes.c_char_p
        L.scf_version.restype = i32
        L.scf_launch_count.restype = i64
        L.scf_measure_fp32_flops.argtypes = [i32, ctypes.POINTER(ctypes.c_double)]
        L.scf_parallel_memcpy.argtypes = [vp, vp, i64]
        L.scf_device_malloc.argtypes = [i32, i64, ctypes.POINTER(vp)]
        L.scf_device_free.argtypes = [i32, vp]
        L.scf_memcpy.argtypes = [i32, vp, vp, i64, i32, vp]
        L.scf_ipc_export.argtypes = [i32, vp, vp]
        L.scf_ipc_import.argtypes = [i32, vp, ctypes.POINTER(vp)]
        L.scf_ipc_close.argtypes = [i32, vp]
        f64 = ctypes.c_double
        L.scf_post_build_cd.argtypes = [vp, i32, i32, f64, f64, ctypes.POINTER(i32), ctypes.POINTER(i32), vp,
                                        ctypes.POINTER(i64)]
        L.scf_post_create.argtypes = [vp, i32, f64, i32, f64, f64, vp, i32, i32, i32, f64, i32, i32, ctypes.POINTER(vp)]
        L.scf_post_destroy.argtypes = [vp]
        L.scf_post_destroy.restype = None
        L.scf_post_reset.argtypes = [vp, vp]
        L.scf_post_decode.argtypes = [vp, vp, i64, vp, vp]
        L.scf_post_step.argtypes = [vp, vp, vp, vp, vp, vp]
        L.scf_post_trigger_update.argtypes = [vp, vp, vp, vp, vp]
        L.scf_post_state.argtypes = [vp, vp, vp, vp]
        L.scf_post_info.argtypes = [vp, ctypes.POINTER(i32), ctypes.POINTER(i32), ctypes.POINTER(i64)]
        L.scf_wav_read_batch.argtypes = [vp, i64, i32, i32, vp, i64, vp, i32]
        L.scf_ingest_wavs.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp]
        L.scf_ingest_wavs_device.argtypes = [vp, vp, i64, i32, i32, i32, vp, vp]
        L.scf_dlpack_alloc.argtypes = [i32, vp, i32, ctypes.POINTER(vp), ctypes.POINTER(vp)]
        L.scf_dlpack_wrap.argtypes = [vp, i32, vp, i32, vp, vp, ctypes.POINTER(vp)]
        _lib = L
        return _lib


def check(rc):
    if rc != 0:
        raise ScfError(rc, lib().scf_last_error().decode('utf-8', 'replace'))


def make_config(sample_rate=16000, window=1024, hop=512, n_fft=1024, n_filt=20, n_coeffs=20,
                bank=BANK_MEL_SONOPY, bank_scale='constant', output=OUT_CEPSTRUM, window_fn='rect',
                preemph_alpha=0.0, pcm_scale=1.0 / 32768.0, device=-1, custom_bank=None, delta=None,
                bank_low_hz=0.0, bank_high_hz=0.0):
    c = Config()
    check(lib().scf_config_default(ctypes.byref(c)))
    c.sample_rate, c.window, c.hop, c.n_fft = int(sample_rate), int(window), int(hop), int(n_fft)
    c.n_filt, c.n_coeffs = int(n_filt), int(n_coeffs)
    c.bank, c.bank_scale, c.output = int(bank), SCALE[bank_scale], int(output)
    c.window_fn, c.preemph_alpha, c.pcm_scale, c.device = WIN[window_fn], float(preemph_alpha), float(pcm_scale), int(device)
    c.delta = DELTA[delta]
    c.bank_low_hz, c.bank_high_hz = float(bank_low_hz or 0.0), float(bank_high_hz or 0.0)
    keep = None
    if custom_bank is not None:
        keep = np.ascontiguousarray(custom_bank, dtype=np.float64)
        c.custom_bank = keep.ctypes.data
    return c, keep


def build_bank(**kw):
    """Dense float64 bank [n_filt, n_fft/2+1] as the library builds it (host only, no GPU needed)."""
    c, keep = make_config(**kw)
    out = np.zeros((c.n_filt, c.n_fft // 2 + 1), dtype=np.float64)
    check(lib().scf_build_bank(ctypes.byref(c), out.ctypes.data))
    return out


def build_dct(n_filt, n_coeffs):
    out = np.zeros((n_filt, min(n_filt, n_coeffs)), dtype=np.float64)
    check(lib().scf_build_dct(n_filt, n_coeffs, out.ctypes.data))
    return out


def bank_apply_tasks(power_a, power_b, **kw):
    """Filterbank sums of one frame pair through the kernel's task list, on the host (test hook).
    Returns (sums_a, sums_b, stats) with stats = dict(tasks, partial_rows, groups, longest_group)."""
    c, keep = make_config(**kw)
    pa = np.ascontiguousarray(power_a, dtype=np.float64)
    pb = np.ascontiguousarray(power_b, dtype=np.float64)
    assert pa.shape == pb.shape == (c.n_fft // 2 + 1,)
    sa, sb = np.zeros(c.n_filt), np.zeros(c.n_filt)
    st = np.zeros(4, dtype=np.int32)
    check(lib().scf_bank_apply_tasks(ctypes.byref(c), pa.ctypes.data, pb.ctypes.data, sa.ctypes.data, sb.ctypes.data,
                                     st.ctypes.data))
    return sa, sb, dict(zip(('tasks', 'partial_rows', 'groups', 'longest_group'), (int(v) for v in st)))


def path_array(paths):
    """list of str -> (char*[] for the C ABI, keep-alive object)"""
    enc = [os.fsencode(p) for p in paths]
    arr = (ctypes.c_char_p * len(enc))(*enc)
    return arr, enc


def num_frames(n_samples, window, hop):
    return int(lib().scf_num_frames(int(n_samples), int(window), int(hop)))
