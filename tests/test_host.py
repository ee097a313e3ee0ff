"""CPU-only tests of the host side: the C-ABI library loads and exports every symbol include/scfeat.h
declares, its float64 table builders equal the oracle's (and therefore the reference's), the params mirror
derives the reference's sizes, and the product path fails loudly without a GPU (no CPU fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

import scfeat
from oracle import bark as obark, pipeline as opipe, sonopy as osonopy
from scfeat import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, 'include', 'scfeat.h')).read()
    declared = set(re.findall(r'^(?:int|void|void\*|int32_t|int64_t|const char\*)\s+(scf_[a-z0-9_]+)\(', header, re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    L = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(L, name), name
    assert _lib.lib().scf_version() == 100


def test_config_default_is_params_json():
    c = _lib.Config()
    _lib.check(_lib.lib().scf_config_default(ctypes.byref(c)))
    assert (c.sample_rate, c.window, c.hop, c.n_fft, c.n_filt, c.n_coeffs) == (16000, 1024, 512, 1024, 20, 20)
    assert (c.bank, c.output, c.window_fn, c.preemph_alpha) == (_lib.BANK_MEL_SONOPY, _lib.OUT_CEPSTRUM, 0, 0.0)
    assert c.pcm_scale == np.float32(1 / 32768)
    assert _lib.lib().scf_out_cols(ctypes.byref(c)) == 20


@pytest.mark.parametrize('n,w,h', [(16000, 1024, 512), (1023, 1024, 512), (1024, 1024, 512), (960000, 1024, 512),
                                   (16000, 160, 80), (0, 1024, 512), (2047, 1024, 512), (2048, 1024, 512)])
def test_num_frames_equals_chop_array(n, w, h):
    assert _lib.num_frames(n, w, h) == len(scfeat.sonopy.chop_array(np.zeros(n), w, h)) == osonopy.n_frames(n, w, h)


@pytest.mark.parametrize('sr,nf,nfft', [(16000, 20, 1024), (16000, 20, 512), (16000, 40, 512), (16000, 26, 1024),
                                        (8000, 13, 256), (44100, 32, 1024), (16000, 64, 1024)])
def test_mel_bank_equals_oracle(sr, nf, nfft):
    got = _lib.build_bank(sample_rate=sr, n_fft=nfft, n_filt=nf, bank=_lib.BANK_MEL_SONOPY)
    want = osonopy.filterbanks(sr, nf, nfft // 2 + 1)
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=1e-15, atol=0)
    assert np.array_equal(scfeat.sonopy.filterbanks(sr, nf, nfft // 2 + 1), got)


def test_mel_bank_equals_cpp_twin(ref_cpp):
    got = _lib.build_bank(sample_rate=16000, n_fft=1024, n_filt=20, bank=_lib.BANK_MEL_SONOPY)
    np.testing.assert_allclose(got, ref_cpp['bank_16000_20_1024'], rtol=1e-14, atol=0)
    # repeated grid points are kept, as the twin does (empty / one-sided filters)
    for n_fft, key in ((512, 'bank_16000_40_512'), (256, 'bank_16000_40_256')):
        got = _lib.build_bank(sample_rate=16000, n_fft=n_fft, n_filt=40, bank=_lib.BANK_MEL_SONOPY)
        want = ref_cpp[key]
        assert np.array_equal(got != 0, want != 0)
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)
    assert (want.sum(axis=1) == 0).any()           # n_fft 256: filter 0 is empty altogether


@pytest.mark.parametrize('nf,nfft,scale', [(20, 512, 'constant'), (20, 1024, 'constant'), (24, 512, 'constant'),
                                           (24, 1024, 'constant'), (26, 512, 'constant'), (26, 1024, 'constant'),
                                           (22, 512, 'ascendant'), (22, 512, 'descendant')])
def test_bark_bank_equals_reference_file(ref_bark, nf, nfft, scale):
    got = scfeat.bark_feature.bark_filterbanks(nfilts=nf, nfft=nfft, sample_rate=16000, scale=scale)
    want = ref_bark['bank_%d_%d_%s' % (nf, nfft, scale)]
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=0)
    np.testing.assert_allclose(got, obark.bark_filterbanks(nfilts=nf, nfft=nfft, scale=scale), rtol=1e-12, atol=0)


@pytest.mark.parametrize('nf,nc', [(20, 20), (26, 13), (20, 30), (24, 24)])
def test_dct_equals_oracle(nf, nc):
    np.testing.assert_allclose(_lib.build_dct(nf, nc), osonopy.dct2_ortho_matrix(nf, nc), rtol=0, atol=1e-15)


# The kernel's bank phase does not multiply by the dense bank: the host cuts it into two-filter tasks of 8 bins, runs
# and partial sums (scfeat_host.cu build_tasks).  scf_bank_apply_tasks walks that work list on the host exactly as the
# kernel does, so the decomposition is checked against bank @ power without a GPU.
@pytest.mark.parametrize('kw', [
    dict(n_fft=1024, n_filt=20, bank=_lib.BANK_MEL_SONOPY),                    # params.json
    dict(n_fft=1024, n_filt=26, bank=_lib.BANK_BARK_REF),                      # config 5 (bfcc)
    dict(n_fft=1024, n_filt=24, bank=_lib.BANK_BARK_REF),                      # config 5 (bark_spec)
    dict(n_fft=512, n_filt=26, bank=_lib.BANK_BARK_REF),                       # bfcc_spec defaults
    dict(n_fft=512, n_filt=24, bank=_lib.BANK_BARK_REF, bank_scale='ascendant'),
    dict(n_fft=512, n_filt=20, bank=_lib.BANK_MEL_SONOPY),                     # sonopy defaults
    dict(n_fft=512, n_filt=40, bank=_lib.BANK_MEL_SONOPY),                     # repeated grid points: empty filters
    dict(n_fft=256, n_filt=40, bank=_lib.BANK_MEL_SONOPY),                     # ... with an empty filter
    dict(n_fft=256, n_filt=13, bank=_lib.BANK_MEL_SONOPY, sample_rate=8000),
    dict(n_fft=1024, n_filt=64, bank=_lib.BANK_MEL_SONOPY),
    dict(n_fft=1024, n_filt=64, bank=_lib.BANK_BARK_REF),
])
def test_bank_task_list_equals_dense_bank(kw):
    bank = _lib.build_bank(**kw)
    rng = np.random.default_rng(11)
    n_bins = bank.shape[1]
    pa = rng.random(n_bins) * 10.0 ** rng.uniform(-8, 4, n_bins)               # wide dynamic range
    pb = rng.random(n_bins)
    sa, sb, st = _lib.bank_apply_tasks(pa, pb, **kw)
    np.testing.assert_allclose(sa, bank @ pa, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(sb, bank @ pb, rtol=1e-12, atol=1e-300)
    assert st['groups'] == 256 // (8 * (1024 // kw['n_fft']))
    # every thread group gets about the same number of tasks (the bank phase is as long as the longest list)
    assert st['longest_group'] <= -(-st['tasks'] // st['groups']) + 2


def test_bank_task_list_custom_bank_with_holes_and_empty_filter():
    rng = np.random.default_rng(5)
    bank = rng.random((7, 513)) * (rng.random((7, 513)) < 0.3)                 # non-contiguous supports
    bank[3] = 0.0                                                               # a filter without any weight
    pa, pb = rng.random(513), rng.random(513)
    sa, sb, _ = _lib.bank_apply_tasks(pa, pb, n_fft=1024, n_filt=7, bank=_lib.BANK_CUSTOM, custom_bank=bank)
    np.testing.assert_allclose(sa, bank @ pa, rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(sb, bank @ pb, rtol=1e-12, atol=1e-14)
    assert sa[3] == 0.0 and sb[3] == 0.0


def test_bank_too_dense_for_shared_memory_is_rejected():
    # 64 Bark filters on 129 bins need more partial-sum rows than the n_fft = 256 kernel keeps in shared memory
    with pytest.raises(_lib.ScfError, match='partial sums'):
        _lib.bank_apply_tasks(np.ones(129), np.ones(129), n_fft=256, n_filt=64, bank=_lib.BANK_BARK_REF)


@pytest.mark.parametrize('nf,nfft,lo,hi,scale', [(20, 512, 300, 6000, 'constant'), (24, 1024, 100, None, 'ascendant'),
                                                 (26, 512, 0, 4000, 'descendant'), (13, 1024, 50.5, 7600, 'constant')])
def test_bark_bank_band_edges_equal_reference_file(nf, nfft, lo, hi, scale):
    """bark_filterbanks(low_freq, high_freq) (common/bark_feature.py:93,104-105) against the unmodified reference file
    (tests/golden/make_golden.py bark_edges) and against the oracle."""
    ref = np.load(os.path.join(os.path.dirname(__file__), 'golden', 'ref_bark_edges.npz'))
    want = ref['bank_%d_%d_%s_%s_%s' % (nf, nfft, lo, hi, scale)]
    got = scfeat.bark_feature.bark_filterbanks(nfilts=nf, nfft=nfft, sample_rate=16000, low_freq=lo, high_freq=hi, scale=scale)
    np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-13)
    np.testing.assert_allclose(obark.bark_filterbanks(nf, nfft, 16000, lo, hi, scale), want, rtol=1e-10, atol=1e-13)
    assert (got != 0).sum() > 0


def test_bark_bank_past_the_last_column_is_dropped_not_an_error():
    # the reference maps bins with nfft = 512 whatever the bank width is: with nfft = 256 its loop runs past column 128
    # and raises IndexError (even with the default band edges); the library keeps the columns that exist
    bank = scfeat.bark_feature.bark_filterbanks(nfilts=13, nfft=256, sample_rate=16000)
    assert bank.shape == (13, 129) and np.isfinite(bank).all() and (bank != 0).any()
    with pytest.raises(_lib.ScfError):
        _lib.build_bank(n_fft=512, n_filt=20, bank=_lib.BANK_BARK_REF, bank_low_hz=4000.0, bank_high_hz=1000.0)


def test_bark_scale_helpers_match_oracle():
    f = np.array([0.0, 100.0, 1000.0, 7999.0])
    np.testing.assert_array_equal(scfeat.bark_feature.hz2bark(f), obark.hz2bark(f))
    np.testing.assert_array_equal(scfeat.bark_feature.bark2hz(f / 400), obark.bark2hz(f / 400))
    np.testing.assert_array_equal(scfeat.bark_feature.fft2bark(f / 40), obark.fft2bark(f / 40))
    np.testing.assert_array_equal(scfeat.bark_feature.bark2fft(f / 400), obark.bark2fft(f / 400))
    assert scfeat.bark_feature.Fm(3.0, 3.0) == 1 and scfeat.bark_feature.Fm(9.0, 3.0) == 0


def test_params_mirror(tmp_path):
    pr = scfeat.params.pr
    o = opipe.Params()
    for k in ('window_samples', 'hop_samples', 'max_samples', 'buffer_samples', 'n_features', 'feature_size'):
        assert getattr(pr, k) == getattr(o, k)
    with pytest.raises(AttributeError):
        pr.n_fft = 512
    f = tmp_path / 'p.json'
    f.write_text('{"n_mfcc": 13, "use_delta": true}')
    try:
        scfeat.params.inject_params(str(f))
        assert pr.n_mfcc == 13 and pr.feature_size == 26
    finally:
        pr.__dict__.update(n_mfcc=20, use_delta=False)
    assert scfeat.params.inject_params(str(tmp_path / 'missing.json')) is pr


def test_host_helpers():
    pcm = np.array([0, 1, -1, 32767, -32768], dtype='<i2')
    a = scfeat.data_utils.buffer_to_audio(pcm.tobytes())
    assert a.dtype == np.float32
    np.testing.assert_array_equal(a, opipe.buffer_to_audio(pcm.tobytes()))
    assert scfeat.data_utils.audio_to_buffer(a) == pcm.tobytes()
    x = np.arange(12, dtype=np.float32).reshape(4, 3) ** 2
    np.testing.assert_array_equal(scfeat.data_utils.add_deltas(x), opipe.add_deltas(x))


def test_bad_config_is_rejected():
    c, _ = _lib.make_config(n_fft=2048)
    h = ctypes.c_void_p()
    assert _lib.lib().scf_plan_create(ctypes.byref(c), ctypes.byref(h)) == -1
    assert b'n_fft' in _lib.lib().scf_last_error()
    with pytest.raises(ValueError):
        scfeat.data_utils.vectorize_raw(np.zeros(0, np.float32))
    with pytest.raises(scfeat.ScfError):          # band edges in the wrong order
        scfeat.bark_feature.bark_filterbanks(nfilts=20, nfft=512, sample_rate=16000, low_freq=3000, high_freq=300)


def test_short_audio_frame_count():
    assert _lib.num_frames(1000, 1024, 512) == 0


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(scfeat.ScfError) as ei:
        scfeat.sonopy.mfcc_spec(np.zeros(16000, np.float32), 16000, (1024, 512), 1024, 20, 20)
    assert ei.value.code == -3


def test_dlpack_capsule_plumbing_on_cpu():
    """The capsule made by scf_dlpack_make_capsule is a valid "dltensor": torch consumes it (deleter runs when torch
    drops the tensor) and an unconsumed capsule runs the deleter when it is garbage collected."""
    import gc
    import torch
    from scfeat import plan

    class DLManagedTensor(ctypes.Structure):
        pass

    DELETER = ctypes.CFUNCTYPE(None, ctypes.POINTER(DLManagedTensor))
    DLManagedTensor._fields_ = [
        ('data', ctypes.c_void_p), ('device_type', ctypes.c_int32), ('device_id', ctypes.c_int32),
        ('ndim', ctypes.c_int32), ('code', ctypes.c_uint8), ('bits', ctypes.c_uint8), ('lanes', ctypes.c_uint16),
        ('shape', ctypes.POINTER(ctypes.c_int64)), ('strides', ctypes.POINTER(ctypes.c_int64)),
        ('byte_offset', ctypes.c_uint64), ('manager_ctx', ctypes.c_void_p), ('deleter', DELETER)]
    deleted = []

    @DELETER
    def deleter(p):
        deleted.append(1)

    def make():
        data = np.arange(6, dtype=np.float32)
        shape = (ctypes.c_int64 * 2)(2, 3)
        mt = DLManagedTensor(data.ctypes.data, 1, 0, 2, 2, 32, 1, shape, None, 0, None, deleter)
        return plan._make_capsule(ctypes.addressof(mt)), (data, shape, mt)

    cap, keep = make()
    t = torch.from_dlpack(cap)
    assert t.shape == (2, 3) and t.dtype == torch.float32 and t.flatten().tolist() == [0, 1, 2, 3, 4, 5]
    del t, cap
    gc.collect()
    assert len(deleted) == 1            # torch owned it and released it
    cap2, keep2 = make()
    del cap2
    gc.collect()
    assert len(deleted) == 2            # nobody consumed it: the capsule destructor released it


def test_add_deltas_both_definitions(ref_cpp):
    """kind='diff' is the Python path's first difference; kind='central' reproduces the C++ twin's deltas
    (mfcc.h:432-441) on the twin's own MFCC output."""
    both = ref_cpp['mfcc_central_delta']
    base = both[:, :20]
    got = scfeat.data_utils.add_deltas(base, kind='central')
    np.testing.assert_allclose(got, both, rtol=0, atol=1e-6)
    d = scfeat.data_utils.add_deltas(base)
    assert d.shape == (30, 40) and not d[0, 20:].any()
    np.testing.assert_array_equal(d[1:, 20:], base[1:] - base[:-1])
    with pytest.raises(ValueError):
        scfeat.data_utils.add_deltas(base, kind='other')


def test_parallel_memcpy_sizes_and_concurrent_callers():
    """The copy threads of the pageable staging path (scf_parallel_memcpy): every size class (one part, uneven parts,
    tails that are not multiples of the 64-byte split) and four callers at once."""
    import threading
    L = _lib.lib()
    rng = np.random.default_rng(9)
    for n in (0, 1, 63, 4096, 256 * 1024 - 1, 256 * 1024, 1_000_003, 5 * 1024 * 1024 + 17):
        src = rng.integers(0, 256, size=n, dtype=np.uint8)
        dst = np.zeros(n + 64, dtype=np.uint8)
        _lib.check(L.scf_parallel_memcpy(dst.ctypes.data, src.ctypes.data, n))
        assert np.array_equal(dst[:n], src) and not dst[n:].any()
    assert L.scf_parallel_memcpy(None, None, 8) != 0           # NULL with a size is an error, not a crash
    srcs = [rng.integers(0, 256, size=3_000_000 + 1000 * i, dtype=np.uint8) for i in range(4)]
    dsts = [np.zeros_like(s) for s in srcs]
    errs = []

    def work(i):
        for _ in range(5):
            dsts[i][:] = 0
            if L.scf_parallel_memcpy(dsts[i].ctypes.data, srcs[i].ctypes.data, srcs[i].size) != 0 or \
                    not np.array_equal(dsts[i], srcs[i]):
                errs.append(i)
    th = [threading.Thread(target=work, args=(i,)) for i in range(4)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs


def test_parallel_memcpy_in_a_forked_child():
    """A forked child inherits the copy pool's bookkeeping but not its threads: it must copy on its own and exit cleanly."""
    import multiprocessing as mp
    L = _lib.lib()
    src = np.arange(2_000_000, dtype=np.uint8)
    dst = np.zeros_like(src)
    _lib.check(L.scf_parallel_memcpy(dst.ctypes.data, src.ctypes.data, src.size))      # the pool exists in the parent

    def child(q):
        d = np.zeros_like(src)
        rc = L.scf_parallel_memcpy(d.ctypes.data, src.ctypes.data, src.size)
        q.put((rc, bool(np.array_equal(d, src))))
    ctx = mp.get_context('fork')
    q = ctx.Queue()
    p = ctx.Process(target=child, args=(q,))
    p.start()
    rc, same = q.get(timeout=60)
    p.join(timeout=60)
    assert rc == 0 and same and p.exitcode == 0
