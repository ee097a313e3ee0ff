"""Parity of the CUDA path (through the C ABI, via the ctypes host package) against the CPU oracle and the
committed reference fixtures.  Run on the B200 box:  python -m pytest tests -m gpu

Tolerances (BASELINE.json north_star, fp32 against the float64 reference):
  * log features (mel_spec / bark_spec):  max |diff| <= 1e-3
  * cepstra (mfcc_spec / bfcc_spec):      max |diff| <= 1e-4 * max |ref| per clip,
                                          and allclose(rtol=1e-4, atol=1e-4)
    (element-wise relative error is meaningless where |ref| ~ 1e-5)
  * power spectrum:                       |diff| <= 1e-5 * max(ref) per frame
"""
import numpy as np
import pytest

import scfeat
from oracle import bark as obark, pipeline as opipe, sonopy as osonopy

pytestmark = pytest.mark.gpu

LOG_TOL = 1e-3
CEP_REL = 1e-4


def audio_of(pcm):
    return pcm.astype(np.float32) / 32768.0


def assert_cepstrum_close(got, want, elementwise=True):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    scale = np.abs(want).reshape(want.shape[0], -1).max(axis=1) if want.ndim == 3 else np.abs(want).max()
    err = np.abs(got - want)
    if want.ndim == 3:
        err = err.reshape(want.shape[0], -1).max(axis=1)
        assert (err <= CEP_REL * scale).all(), (err / scale).max()
    else:
        assert err.max() <= CEP_REL * scale, err.max() / scale
    if elementwise:
        np.testing.assert_allclose(got, want, rtol=1e-4, atol=1e-4)


def assert_log_close(got, want):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= LOG_TOL, np.abs(got - want).max()


# ------------------------------------------------------------------ config 1: the 8 example wavs
def test_example_wavs_mfcc_vs_oracle(example_pcm):
    _, pcm = example_pcm
    want = np.stack([osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 20) for a in audio_of(pcm)])
    # int16 PCM in, one batched call (classifier/data.py path)
    got = scfeat.data_utils.extract_features_batch(pcm)
    assert got.shape == (8, 30, 20, 1) and got.dtype == np.float32
    assert_cepstrum_close(got[..., 0], want)
    # float audio in, one call per clip (sonopy.mfcc_spec call shape, data_utils.py:69)
    for a, w in zip(audio_of(pcm), want):
        g = scfeat.sonopy.mfcc_spec(a, 16000, (1024, 512), num_filt=20, fft_size=1024, num_coeffs=20)
        assert_cepstrum_close(g, w)


def test_example_wavs_vs_compiled_reference_cpp(example_pcm, ref_cpp):
    _, pcm = example_pcm
    got = scfeat.data_utils.extract_features_batch(pcm)[..., 0]
    assert_cepstrum_close(got, ref_cpp['mfcc_params_json'])        # inference/tflite/mfcc.h output
    names, _ = example_pcm
    r = names.index('right_1')
    np.testing.assert_allclose(got[r, 0, :5], [-1.71329153, 1.68849206, 2.06822109, 3.96351266, 2.23058176], atol=2e-5)


def test_example_wavs_logmel_and_power(example_pcm, ref_bark):
    _, pcm = example_pcm
    a = audio_of(pcm)
    for x in a:
        assert_log_close(scfeat.sonopy.mel_spec(x, 16000, (1024, 512), 1024, 20),
                         osonopy.mel_spec(x, 16000, (1024, 512), 1024, 20))
    for i in range(2):
        got = scfeat.sonopy.power_spec(a[i], (1024, 512), 1024).astype(np.float64)
        want = ref_bark['power_1024_512_1024'][i]
        assert got.shape == want.shape == (30, 513)
        assert (np.abs(got - want).max(axis=1) <= 1e-5 * want.max(axis=1)).all()


def test_example_wavs_bark_and_bfcc_vs_reference_file(example_pcm, ref_bark):
    _, pcm = example_pcm
    a = audio_of(pcm)
    for key, args in [('bfcc_1024_512_1024_20_20', (1024, 512, 1024, 20, 20)),
                      ('bfcc_1024_512_1024_26_13', (1024, 512, 1024, 26, 13)),
                      ('bfcc_512_256_512_26_13', (512, 256, 512, 26, 13))]:
        got = np.stack([scfeat.bark_feature.bfcc_spec(x, 16000, *args) for x in a])
        assert_cepstrum_close(got, ref_bark[key])
    for key, args in [('bark_1024_512_1024_24', (1024, 512, 1024, 24)), ('bark_1024_512_1024_20', (1024, 512, 1024, 20)),
                      ('bark_512_256_512_24', (512, 256, 512, 24))]:
        got = np.stack([scfeat.bark_feature.bark_spec(x, 16000, *args) for x in a])
        assert_log_close(got, ref_bark[key])


def test_long_clip_bark_config5_shape(ref_bark):
    lf = audio_of(ref_bark['long_pcm'])
    assert_cepstrum_close(scfeat.bark_feature.bfcc_spec(lf, 16000, 1024, 512, 1024, 26, 13),
                          ref_bark['long_bfcc_1024_512_1024_26_13'])
    assert_log_close(scfeat.bark_feature.bark_spec(lf, 16000, 1024, 512, 1024, 24), ref_bark['long_bark_1024_512_1024_24'])
    # int16 input gives the same answer as the float path
    g16 = scfeat.bark_feature.bfcc_spec(ref_bark['long_pcm'], 16000, 1024, 512, 1024, 26, 13)
    assert_cepstrum_close(g16, ref_bark['long_bfcc_1024_512_1024_26_13'])


# ------------------------------------------------------------------ config 2: batch of 512 synthetic clips
def synth_batch(kind, example):
    rng = np.random.default_rng(0)
    if kind == 'uniform':
        return rng.integers(-32768, 32768, size=(512, 16000), dtype=np.int16)
    if kind == 'gaussian':
        return np.clip(np.round(rng.normal(0, 3000, size=(512, 16000))), -32768, 32767).astype(np.int16)
    return np.tile(example, (64, 1))


@pytest.mark.parametrize('kind', ['uniform', 'gaussian', 'tiled'])
def test_batch512_vs_oracle(example_pcm, kind):
    pcm = synth_batch(kind, example_pcm[1])
    got = scfeat.data_utils.extract_features_batch(pcm)[..., 0]
    assert got.shape == (512, 30, 20)
    idx = np.arange(512) if kind != 'tiled' else np.arange(8)
    want = np.stack([osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 20) for a in audio_of(pcm[idx])])
    assert_cepstrum_close(got[idx], want)
    if kind == 'tiled':      # identical inputs -> bit-identical rows wherever they sit in the batch
        for r in range(1, 64):
            assert np.array_equal(got[8 * r:8 * r + 8], got[:8])


# ------------------------------------------------------------------ edge cases the reference defines
def test_all_zero_clip_and_short_inputs():
    z = scfeat.sonopy.mfcc_spec(np.zeros(16000, np.float32), 16000, (1024, 512), 1024, 20, 20)
    assert z.shape == (30, 20)
    np.testing.assert_allclose(z[:, 0], -36.04365339, atol=1e-4)
    assert np.abs(z[:, 1:]).max() < 1e-4
    assert scfeat.sonopy.mfcc_spec(np.zeros(1023, np.float32), 16000, (1024, 512), 1024, 20, 20).shape == (0, 20)
    assert scfeat.sonopy.mfcc_spec(np.zeros(1023, np.float32), 16000, (1024, 512), 1024, 20, 30).shape == (0, 20)
    assert scfeat.bark_feature.bfcc_spec(np.zeros(10, np.float32), 16000, 1024, 512).shape == (0, 13)
    assert scfeat.sonopy.power_spec(np.zeros(100, np.float32)).shape == (0, 257)
    one = scfeat.sonopy.mfcc_spec(np.ones(1024, np.float32), 16000, (1024, 512), 1024, 20, 20)
    assert one.shape == (1, 20)
    assert_cepstrum_close(one, osonopy.mfcc_spec(np.ones(1024, np.float32), 16000, (1024, 512), 1024, 20, 20))


def test_odd_frame_count_and_ncoeffs_gt_nfilt(example_pcm):
    _, pcm = example_pcm
    a = audio_of(pcm[3])[:1024 + 512 * 6]          # 7 frames: the last pair is half empty
    want = osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 30)
    got = scfeat.sonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 30)
    assert got.shape == want.shape == (7, 20)
    assert_cepstrum_close(got, want)


@pytest.mark.parametrize('window,hop,nfft,nf,nc', [(160, 80, 512, 20, 13), (512, 256, 512, 20, 13), (400, 160, 512, 26, 13),
                                                   (256, 128, 256, 20, 13), (1200, 400, 1024, 20, 20),
                                                   (1024, 300, 1024, 40, 20), (640, 320, 1024, 24, 12)])
def test_other_geometries_vs_oracle(example_pcm, window, hop, nfft, nf, nc):
    _, pcm = example_pcm
    for a in audio_of(pcm[:2]):
        want = osonopy.mfcc_spec(a, 16000, (window, hop), nfft, nf, nc)
        got = scfeat.sonopy.mfcc_spec(a, 16000, (window, hop), nfft, nf, nc)
        assert_cepstrum_close(got, want)
        assert_log_close(scfeat.sonopy.mel_spec(a, 16000, (window, hop), nfft, nf),
                         osonopy.mel_spec(a, 16000, (window, hop), nfft, nf))


def test_sonopy_defaults(example_pcm, oracle_pin):
    _, pcm = example_pcm
    got = np.stack([scfeat.sonopy.mfcc_spec(a, 16000) for a in audio_of(pcm[:2])])
    assert_cepstrum_close(got, oracle_pin['mfcc_sonopy_defaults'])
    p, f, m, c = scfeat.sonopy.mfcc_spec(audio_of(pcm[0]), 16000, return_parts=True)
    assert p.shape[1] == 257 and f.shape == (20, 257) and m.shape[1] == 20 and c.shape[1] == 13


def test_audio_to_feature_crop_pad_and_ragged_batch(example_pcm):
    _, pcm = example_pcm
    p = opipe.Params()
    a = audio_of(pcm[0])
    # head crop
    longer = np.concatenate([a, a[:5000]])
    assert_cepstrum_close(scfeat.data_utils.audio_to_feature(longer), opipe.audio_to_feature(longer, p))
    # front pad
    for n in (9000, 1024, 700, 1):
        assert_cepstrum_close(scfeat.data_utils.audio_to_feature(a[:n]), opipe.audio_to_feature(a[:n], p))
    # ragged batch of int16 clips: per-clip lengths, pad applied by the kernel
    lengths = np.array([16000, 9000, 12345, 1024, 1023, 1, 15999, 8000], dtype=np.int32)
    got = scfeat.data_utils.extract_features_batch(pcm, lengths)[..., 0]
    want = np.stack([opipe.audio_to_feature(audio_of(pcm[i][:lengths[i]]), p) for i in range(8)])
    assert_cepstrum_close(got, want)
    assert scfeat.data_utils.vectorize_raw(a[:1000]).shape == (0, 20)
    assert_cepstrum_close(scfeat.data_utils.vectorize_raw(a[:5000]), opipe.vectorize_raw(a[:5000], p))


def test_silent_frame_next_to_loud_frame(example_pcm):
    """An exactly-zero frame (front padding) that shares an FFT with a loud frame must still come out as the
    reference's floor value ln(eps) in every band: the two-for-one FFT may not leak its partner into it."""
    _, pcm = example_pcm
    p = opipe.Params()
    rng = np.random.default_rng(11)
    loud = rng.integers(-32768, 32768, size=16000, dtype=np.int16)
    for n in (1500, 2012, 2524, 9000, 9512):             # both parities of the first non-silent frame
        for src in (pcm[4], loud):
            x = src[:n]
            want = opipe.audio_to_feature(audio_of(x), p)
            got = scfeat.data_utils.audio_to_feature(x)                       # int16, front pad in the kernel
            assert_cepstrum_close(got, want)
            gotf = scfeat.data_utils.audio_to_feature(audio_of(x))           # float path
            assert_cepstrum_close(gotf, want)
            silent = np.flatnonzero(np.abs(want[:, 1:]).max(axis=1) < 1e-9)
            assert len(silent) > 0
            np.testing.assert_allclose(got[silent, 0], -36.04365339, atol=2e-5)
    # same through the fast path: zeros written into a full-length clip, log-mel bands too
    clip = np.zeros(16000, dtype=np.int16)
    clip[14336:] = loud[:1664]                                               # frames 0..26 silent, 27.. loud
    got = scfeat.sonopy.mel_spec(clip, 16000, (1024, 512), 1024, 20)
    want = osonopy.mel_spec(audio_of(clip), 16000, (1024, 512), 1024, 20)
    assert_log_close(got, want)
    np.testing.assert_allclose(got[:27], -36.04365339, atol=2e-5)


def test_use_delta_and_changed_pr(example_pcm):
    _, pcm = example_pcm
    pr = scfeat.params.pr
    a = audio_of(pcm[1])
    try:
        pr.__dict__.update(n_mfcc=13, n_filt=26, use_delta=True, window_t=0.032, hop_t=0.016, n_fft=512)
        p = opipe.Params(n_mfcc=13, n_filt=26, use_delta=True, window_t=0.032, hop_t=0.016, n_fft=512)
        got = scfeat.data_utils.audio_to_feature(a)
        want = opipe.audio_to_feature(a, p)
        assert got.shape == want.shape == (p.n_features, 26)
        assert_cepstrum_close(got, want)
    finally:
        pr.__dict__.update(n_mfcc=20, n_filt=20, use_delta=False, window_t=0.064, hop_t=0.032, n_fft=1024)


def test_preemphasis_and_hamming_vs_oracle(example_pcm):
    """Optional front end of the C++ twin (mfcc.h:394-410): x[n] - 0.95 x[n-1] (x[-1] := 0), Hamming with
    denominator window-1.  Oracle = the same two steps in float64 followed by the sonopy restatement."""
    _, pcm = example_pcm
    a = audio_of(pcm[2]).astype(np.float64)
    pre = a.copy()
    pre[1:] -= np.float32(0.95).astype(np.float64) * a[:-1]
    W, H = 1024, 512
    ham = 0.54 - 0.46 * np.cos(2 * np.pi * np.arange(W) / (W - 1))
    frames = osonopy.frames_of(pre, W, H) * ham
    spec = np.fft.rfft(frames, n=1024)
    powers = (spec.real ** 2 + spec.imag ** 2) / 1024
    from scipy.fftpack import dct
    mels = osonopy.safe_log(powers @ osonopy.filterbanks(16000, 20, 513).T)
    want = dct(mels, norm='ortho')[:, :20]
    want[:, 0] = osonopy.safe_log(powers.sum(1))
    plan = scfeat.get_plan(window=W, hop=H, n_fft=1024, n_filt=20, n_coeffs=20, preemph_alpha=0.95, window_fn='hamming')
    assert_cepstrum_close(plan.extract_host(pcm[2]), want)
    assert_cepstrum_close(plan.extract_host(audio_of(pcm[2])), want)


def _front_end_oracle(audio, alpha, window_fn, W, H, n_fft, n_filt, n_coeffs):
    """float64: x[n] - alpha x[n-1] (x[-1] := 0) over the whole clip, frames times the window (denominator W - 1,
    mfcc.h:394-410), then the sonopy restatement's power -> mel -> log -> DCT with c0 := log energy."""
    from scipy.fftpack import dct
    a = np.asarray(audio, dtype=np.float64)
    pre = a.copy()
    pre[1:] -= np.float64(np.float32(alpha)) * a[:-1]
    n = np.arange(W)
    win = {'rect': np.ones(W), 'hamming': 0.54 - 0.46 * np.cos(2 * np.pi * n / (W - 1)),
           'hann': 0.5 - 0.5 * np.cos(2 * np.pi * n / (W - 1))}[window_fn]
    frames = osonopy.frames_of(pre, W, H) * win
    spec = np.fft.rfft(frames, n=n_fft)
    powers = (spec.real ** 2 + spec.imag ** 2) / n_fft
    mels = osonopy.safe_log(powers @ osonopy.filterbanks(16000, n_filt, n_fft // 2 + 1).T)
    out = dct(mels, norm='ortho')[:, :n_coeffs]
    out[:, 0] = osonopy.safe_log(powers.sum(1))
    return out


@pytest.mark.parametrize('n_fft, alpha, window_fn, nf', [(1024, 0.95, 'hamming', 20), (1024, 0.97, 'rect', 20),
                                                         (1024, 0.0, 'hann', 20), (512, 0.95, 'hamming', 20),
                                                         (256, 0.95, 'hamming', 10)])
def test_fused_front_end_on_the_fast_kernels(example_pcm, n_fft, alpha, window_fn, nf):
    """Pre-emphasis and / or a window on the fast geometry (window = n_fft, hop = n_fft / 2) run inside the fast loader
    (mfcc.h:394-410): small and large jobs (all kernel variants), int16 and float input, short clips padded in front,
    an odd number of frames -- against the float64 oracle, and against the generic loader on the same input."""
    _, pcm = example_pcm
    W, H = n_fft, n_fft // 2
    # (n_fft = 256 with 10 filters: with 20, filter 0 of sonopy's grid is bin 0 alone, and the DC term of a
    #  pre-emphasised, windowed frame is a sum of 256 cancelling values -- 1e-12 of the frame energy, which fp32 cannot
    #  resolve: |diff| up to 2e-2 in that one log band, the same bit for bit through the generic loader)
    plan = scfeat.get_plan(window=W, hop=H, n_fft=n_fft, n_filt=nf, n_coeffs=nf, preemph_alpha=alpha, window_fn=window_fn)
    want8 = np.stack([_front_end_oracle(a, alpha, window_fn, W, H, n_fft, nf, nf) for a in audio_of(pcm)])
    assert_cepstrum_close(plan.extract_host(pcm), want8)
    assert_cepstrum_close(plan.extract_host(audio_of(pcm)), want8)
    # a large job (three-team kernels) of rolled copies: every clip differs, the oracle is evaluated on a sample
    n = 1664
    rng = np.random.default_rng(5)
    shifts = rng.integers(0, 16000, size=n)
    clips = np.stack([np.roll(pcm[i % 8], shifts[i]) for i in range(n)])
    got = plan.extract_host(clips)
    for i in list(range(0, n, 97)) + [n - 1]:
        assert_cepstrum_close(got[i], _front_end_oracle(audio_of(clips[i]), alpha, window_fn, W, H, n_fft, nf, nf))
    generic = plan.extract_host(clips, lengths=np.full(n, 16000, np.int32), pad=scfeat.plan.PAD_NONE)
    np.testing.assert_array_equal(got, generic)          # the generic loader: same arithmetic, bit for bit
    # short clips, zero-padded in front (common/data_utils.py:77-80), incl. every alignment class of the first sample
    m = 72
    lengths = rng.integers(0, 16001, size=m).astype(np.int32)
    lengths[:10] = [16000, 0, 1, 2, 15999, 16000 - H, 16000 - H - 1, 16000 - W, 16000 - W - 1, 16000 - W + 1]
    got = plan.extract_host(clips[:m], lengths=lengths)
    for i in range(m):
        padded = np.concatenate([np.zeros(16000 - lengths[i], np.float32), audio_of(clips[i][:lengths[i]])])
        assert_cepstrum_close(got[i], _front_end_oracle(padded, alpha, window_fn, W, H, n_fft, nf, nf))
    # odd frame count (the last pair of a clip has no second frame), and the generic loader on the same samples
    L = W + 8 * H                                   # 9 frames
    got = plan.extract_host(clips[:40, :L])
    assert got.shape[1] == 9
    for i in range(0, 40, 7):
        assert_cepstrum_close(got[i], _front_end_oracle(audio_of(clips[i, :L]), alpha, window_fn, W, H, n_fft, nf, nf))
    generic = plan.extract_host(clips[:40, :L], lengths=np.full(40, L, np.int32), pad=scfeat.plan.PAD_NONE)
    np.testing.assert_array_equal(got, generic)          # same arithmetic, bit for bit


def test_custom_bank_matches_builtin(example_pcm):
    _, pcm = example_pcm
    bank = obark.bark_filterbanks(nfilts=24, nfft=1024)
    plan = scfeat.get_plan(n_filt=24, bank=scfeat.plan.BANK_CUSTOM, custom_bank=bank, output=scfeat.plan.OUT_LOG_BANK)
    got = plan.extract_host(pcm[:2])
    want = np.stack([obark.bark_spec(a, 16000, 1024, 512, 1024, 24) for a in audio_of(pcm[:2])])
    assert_log_close(got, want)


# ------------------------------------------------------------------ size-independent properties at full size
def test_properties_at_corpus_shard_size(example_pcm):
    """config 3 per-rank shard (13,229 clips): batch-invariance, determinism and the log-domain scaling law."""
    rng = np.random.default_rng(1000)
    n = 13229
    pcm = rng.integers(-16384, 16384, size=(n, 16000), dtype=np.int16)
    pcm[:8] = example_pcm[1] // 2
    got = scfeat.data_utils.extract_features_batch(pcm)[..., 0]
    assert got.shape == (n, 30, 20) and np.isfinite(got).all()
    again = scfeat.data_utils.extract_features_batch(pcm)[..., 0]
    assert np.array_equal(got, again)                                        # deterministic
    sel = rng.choice(n, 64, replace=False)
    solo = scfeat.data_utils.extract_features_batch(pcm[sel])[..., 0]
    assert np.array_equal(solo, got[sel])                                    # independent of batch position
    want = np.stack([osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 20) for a in audio_of(pcm[sel[:16]])])
    assert_cepstrum_close(got[sel[:16]], want)
    # doubling the samples adds ln 4 to every log band: c0 += ln 4, c1.. unchanged except through the DCT of a constant
    dbl = scfeat.data_utils.extract_features_batch(pcm[:256] * 2)[..., 0]
    np.testing.assert_allclose(dbl[:, :, 0] - got[:256, :, 0], np.log(4.0), atol=2e-5)
    np.testing.assert_allclose(dbl[:, :, 2:], got[:256, :, 2:], atol=2e-4)


def test_long_form_batch_config5_properties():
    """config 5 shape (60 s clips, Bark bank, fft 1024) on a 32-clip slice: 1874 frames per clip, batch rows equal
    single-clip rows, spot rows equal the oracle."""
    rng = np.random.default_rng(3)
    pcm = rng.integers(-32768, 32768, size=(32, 960000), dtype=np.int16)
    plan = scfeat.get_plan(window=1024, hop=512, n_fft=1024, n_filt=26, n_coeffs=13, bank=scfeat.plan.BANK_BARK_REF)
    got = plan.extract_host(pcm)
    assert got.shape == (32, 1874, 13) and np.isfinite(got).all()
    assert np.array_equal(plan.extract_host(pcm[5]), got[5])
    want = obark.bfcc_spec(audio_of(pcm[7][:1024 + 512 * 99]), 16000, 1024, 512, 1024, 26, 13)
    assert_cepstrum_close(got[7, :100], want)
    want_tail = obark.bfcc_spec(audio_of(pcm[31][-(1024 + 512 * 9):]), 16000, 1024, 512, 1024, 26, 13)
    assert_cepstrum_close(got[31, -10:], want_tail)


# ------------------------------------------------------------------ device pointers + DLPack (torch is only plumbing)
def test_device_api_and_dlpack(example_pcm):
    import torch
    _, pcm = example_pcm
    plan = scfeat.get_plan()
    d_in = torch.from_numpy(pcm).cuda()
    d_out = torch.empty((8, 30, 20), dtype=torch.float32, device='cuda')
    st = torch.cuda.current_stream()
    plan.extract_device(d_in.data_ptr(), 8, 16000, d_out.data_ptr(), stream=st.cuda_stream)
    st.synchronize()
    host = plan.extract_host(pcm)
    assert np.array_equal(d_out.cpu().numpy(), host)
    cap = plan.extract_dlpack(d_in.data_ptr(), 8, 16000, stream=st.cuda_stream)
    t = torch.from_dlpack(cap)
    st.synchronize()
    assert t.is_cuda and t.dtype == torch.float32 and tuple(t.shape) == (8, 30, 20)
    assert np.array_equal(t.cpu().numpy(), host)
    del t, cap
    cap2 = plan.extract_dlpack(d_in.data_ptr(), 8, 16000, stream=st.cuda_stream)
    del cap2                                       # an unconsumed capsule frees its buffer
    # strided clips: stride > clip_len
    wide = torch.zeros((8, 16384), dtype=torch.int16, device='cuda')
    wide[:, :16000] = d_in
    plan.extract_device(wide.data_ptr(), 8, 16000, d_out.data_ptr(), clip_stride=16384, stream=st.cuda_stream)
    st.synchronize()
    assert np.array_equal(d_out.cpu().numpy(), host)


def test_async_host_api_matches_sync(example_pcm):
    import torch
    _, pcm = example_pcm
    plan = scfeat.get_plan()
    rng = np.random.default_rng(9)
    batches = [rng.integers(-32768, 32768, size=(64, 16000), dtype=np.int16) for _ in range(5)]
    batches[2][:8] = pcm
    pinned_in = [torch.from_numpy(b).pin_memory().numpy() for b in batches]
    outs = [torch.empty((64, 30, 20), dtype=torch.float32).pin_memory().numpy() for _ in batches]
    for b, o in zip(pinned_in, outs):
        plan.extract_host_async(b, o)
    plan.host_sync()
    for b, o in zip(batches, outs):
        assert np.array_equal(o, plan.extract_host(b))
    lengths = np.full(64, 9000, dtype=np.int32)
    plan.extract_host_async(pinned_in[0], outs[0], lengths=lengths)
    plan.host_sync()
    assert np.array_equal(outs[0], plan.extract_host(batches[0], lengths=lengths))
    with pytest.raises(ValueError):
        plan.extract_host_async(batches[0].astype(np.float32), outs[0])


# ------------------------------------------------------------------ config 4: streaming
@pytest.mark.parametrize('chunk', [1600, 1024, 512, 3000])
def test_stream_matches_listener_oracle(example_pcm, chunk):
    _, pcm = example_pcm
    p = opipe.Params()
    n_streams = 5
    x = np.stack([np.concatenate([pcm[i], pcm[(i + 3) % 8]]) for i in range(n_streams)])
    fs = scfeat.listener.FeatureStream(n_streams, max_chunk=4096)
    oracles = [opipe.ListenerOracle(p) for _ in range(n_streams)]
    for s in range(0, x.shape[1] - chunk + 1, chunk):
        ring, new = fs.push(x[:, s:s + chunk])
        for i, o in enumerate(oracles):
            before = len(o.window_audio)
            want = o.update_vectors(x[i, s:s + chunk].tobytes())[..., 0]
            assert new[i] == (before + chunk - len(o.window_audio)) // p.hop_samples
            scale = max(np.abs(want).max(), 1.0)
            assert np.abs(ring[i] - want).max() <= CEP_REL * scale
    one = scfeat.listener.Listener()
    o = opipe.ListenerOracle(p)
    for s in range(0, 16000 - chunk + 1, chunk):
        got = one.update_vectors(pcm[2, s:s + chunk].tobytes())
        want = o.update_vectors(pcm[2, s:s + chunk].tobytes())
        assert got.shape == want.shape == (30, 20, 1)
        assert np.abs(got - want).max() <= CEP_REL * max(np.abs(want).max(), 1.0)


def test_stream_256_streams_config4():
    rng = np.random.default_rng(2)
    T = 12
    chunks = rng.integers(-32768, 32768, size=(256, T, 1600), dtype=np.int16)
    fs = scfeat.listener.FeatureStream(256, max_chunk=1600)
    total = np.zeros(256, dtype=np.int64)
    for t in range(T):
        ring, new = fs.push(chunks[:, t])
        total += new
    assert (total == (T * 1600 - 1024) // 512 + 1).all()
    for i in (0, 100, 255):
        full = osonopy.mfcc_spec(audio_of(chunks[i].reshape(-1)), 16000, (1024, 512), 1024, 20, 20)
        want = full[total[i] - 30:total[i]]
        assert np.abs(ring[i] - want).max() <= CEP_REL * np.abs(want).max()


def test_stream_device_push_ring_copy_and_small_ring(example_pcm):
    """Device-pointer push (caller's ring copy + new-row counts, no host sync) equals the host-buffer push, also when a
    step emits more frames than the ring has rows (listen.py:107-108 keeps the newest ones) and after a reset."""
    import ctypes
    import torch
    from scfeat import _lib
    _, pcm = example_pcm
    n_streams, chunk = 3, 3000                                   # up to 6 frames per step
    x = np.stack([np.concatenate([pcm[i], pcm[i + 4]]) for i in range(n_streams)])
    plan = scfeat.get_plan()
    for rows in (30, 2):
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().scf_stream_create(plan.handle, n_streams, rows, chunk, ctypes.byref(h)))
        try:
            d_ring = torch.full((n_streams, rows, 20), float('nan'), dtype=torch.float32, device='cuda')
            d_new = torch.zeros((n_streams,), dtype=torch.int32, device='cuda')
            for rep in range(2):                                 # second pass: after a reset
                want = np.zeros((n_streams, rows, 20))
                emitted = np.zeros(n_streams, dtype=np.int64)
                for s in range(0, x.shape[1] - chunk + 1, chunk):
                    d_chunk = torch.from_numpy(np.ascontiguousarray(x[:, s:s + chunk])).cuda()
                    _lib.check(_lib.lib().scf_stream_push_i16(h, d_chunk.data_ptr(), chunk, d_ring.data_ptr(),
                                                              d_new.data_ptr(), None))
                    torch.cuda.synchronize()
                    for i in range(n_streams):
                        full = osonopy.mfcc_spec(audio_of(x[i, :s + chunk]), 16000, (1024, 512), 1024, 20, 20)
                        k = len(full) - emitted[i]
                        assert int(d_new[i]) == k
                        emitted[i] = len(full)
                        want[i] = np.concatenate([want[i], full[len(full) - k:]])[-rows:]
                    got = d_ring.cpu().numpy()
                    assert np.abs(got - want).max() <= CEP_REL * max(np.abs(want).max(), 1.0)
                _lib.check(_lib.lib().scf_stream_reset(h, None))
        finally:
            _lib.lib().scf_stream_destroy(h)


@pytest.mark.parametrize('n_streams,chunk,rows,steps', [(256, 1600, 30, 14), (3, 3000, 2, 10), (64, 512, 30, 24)])
def test_stream_back_to_back_pushes_without_host_sync(n_streams, chunk, rows, steps):
    """include/scfeat.h: pushes of one scf_stream are stream-ordered -- no host synchronisation is needed between them.
    Every push reads the state its predecessor wrote, and consecutive launches overlap through programmatic
    dependent launch, so the kernel has to wait for its predecessor before its first READ of stream state.  All
    pushes are enqueued at once; each step's ring and row count go to their own device buffers and are compared
    with the listen.py:96-114 state machine after ONE synchronisation at the end."""
    import ctypes
    import torch
    from scfeat import _lib
    p = opipe.Params()
    rng = np.random.default_rng(31)
    x = rng.integers(-32768, 32768, size=(n_streams, steps, chunk), dtype=np.int16)
    plan = scfeat.get_plan()
    h = ctypes.c_void_p()
    _lib.check(_lib.lib().scf_stream_create(plan.handle, n_streams, rows, chunk, ctypes.byref(h)))
    try:
        for rep in range(3):                        # repeated: a race does not have to show on the first attempt
            st = torch.cuda.Stream()
            d_chunks = torch.from_numpy(np.ascontiguousarray(x.transpose(1, 0, 2))).cuda()       # [steps][streams][chunk]
            d_rings = torch.full((steps, n_streams, rows, 20), float('nan'), dtype=torch.float32, device='cuda')
            d_new = torch.full((steps, n_streams), -1, dtype=torch.int32, device='cuda')
            torch.cuda.synchronize()
            _lib.check(_lib.lib().scf_stream_reset(h, ctypes.c_void_p(st.cuda_stream)))
            for t in range(steps):
                _lib.check(_lib.lib().scf_stream_push_i16(h, d_chunks[t].data_ptr(), chunk, d_rings[t].data_ptr(),
                                                          d_new[t].data_ptr(), ctypes.c_void_p(st.cuda_stream)))
            st.synchronize()
            rings, new = d_rings.cpu().numpy(), d_new.cpu().numpy()
            check = range(n_streams) if n_streams <= 64 else list(range(0, n_streams, 9)) + [n_streams - 1]
            for i in check:
                o = opipe.ListenerOracle(p)
                o.mfccs = np.zeros((rows, p.n_mfcc))
                for t in range(steps):
                    before = len(o.window_audio)
                    want = o.update_vectors(x[i, t].tobytes())[..., 0]
                    k = (before + chunk - len(o.window_audio)) // p.hop_samples
                    assert new[t, i] == k, (rep, i, t)
                    assert np.abs(rings[t, i] - want).max() <= CEP_REL * max(np.abs(want).max(), 1.0), (rep, i, t)
            # the row counts of ALL streams (cheap): same chunk length -> same count for every stream
            assert (new == new[:, :1]).all()
    finally:
        _lib.lib().scf_stream_destroy(h)


def test_mel_grid_with_repeated_points_vs_compiled_reference_cpp(example_pcm, ref_cpp):
    """n_fft 512 / 256 with 40 mel filters: the grid has repeated points, some filters are one-sided or empty
    (log(eps) = -36.04 in that band).  The reference's C++ twin (mfcc.h:230-264) keeps them; so does the kernel."""
    _, pcm = example_pcm
    for (w, hop, nfft), key in (((512, 256, 512), 'mfcc_512_256_512_40_13_dupgrid'),
                                ((256, 128, 256), 'mfcc_256_128_256_40_13_dupgrid')):
        want = ref_cpp[key]
        got = np.stack([scfeat.sonopy.mfcc_spec(a, 16000, (w, hop), nfft, 40, 13) for a in audio_of(pcm[:2])])
        assert_cepstrum_close(got, want)
        orc = np.stack([osonopy.mfcc_spec(a, 16000, (w, hop), nfft, 40, 13) for a in audio_of(pcm[:2])])
        assert_cepstrum_close(got, orc)
    mel = scfeat.sonopy.mel_spec(audio_of(pcm[0]), 16000, (256, 128), 256, 40)
    assert_log_close(mel, osonopy.mel_spec(audio_of(pcm[0]), 16000, (256, 128), 256, 40))
    np.testing.assert_allclose(mel[:, 0], -36.04365339, atol=2e-5)             # filter 0 is empty


def test_preemphasis_and_hamming_vs_compiled_reference_cpp(example_pcm, ref_cpp):
    """The same front end pinned to the reference itself: mfcc::mfcc<float>(use_preprocess=true) compiled from
    inference/tflite/mfcc.h:394-410 (tests/golden/ref_mfcc_cpp.npz 'mfcc_preproc').  Frame 0 is skipped: its first
    sample reads audio_data[-1] in the reference (out of bounds); the contract here is x[-1] := 0."""
    _, pcm = example_pcm
    want = ref_cpp['mfcc_preproc']
    plan = scfeat.get_plan(window=1024, hop=512, n_fft=1024, n_filt=20, n_coeffs=20, preemph_alpha=0.95, window_fn='hamming')
    got = plan.extract_host(pcm[:2])
    assert got.shape == want.shape == (2, 30, 20)
    assert_cepstrum_close(got[:, 1:], want[:, 1:])
    assert_cepstrum_close(plan.extract_host(audio_of(pcm[:2]))[:, 1:], want[:, 1:])


# ------------------------------------------------------------------ delta features on the device
def test_delta_features_on_device(example_pcm, ref_cpp):
    """scf_config.delta: 'diff' = add_deltas (common/data_utils.py:50-58), 'central' = mfcc::mfcc(use_delta)
    (inference/tflite/mfcc.h:432-441, pinned to the compiled twin's output), 'central2' = ... plus use_delta2
    (mfcc.h:443-453).  The delta columns are exact functions of the float32 base columns the same call returns."""
    _, pcm = example_pcm
    base = scfeat.get_plan().extract_host(pcm)                                   # [8, 30, 20]
    diff = scfeat.get_plan(delta='diff').extract_host(pcm)
    assert diff.shape == (8, 30, 40)
    assert np.array_equal(diff[..., :20], base)
    want = np.zeros_like(base)
    want[:, 1:] = base[:, 1:] - base[:, :-1]
    assert np.array_equal(diff[..., 20:], want)
    for i in range(8):                                                           # ... and equal the oracle's add_deltas
        o = opipe.add_deltas(osonopy.mfcc_spec(audio_of(pcm[i]), 16000, (1024, 512), 1024, 20, 20))
        assert_cepstrum_close(diff[i], o)
    cen = scfeat.get_plan(delta='central').extract_host(pcm)
    nxt, prv = np.minimum(np.arange(30) + 1, 29), np.maximum(np.arange(30) - 1, 0)
    d1 = (base[:, nxt] - base[:, prv]) / 2
    assert cen.shape == (8, 30, 40) and np.array_equal(cen[..., :20], base) and np.array_equal(cen[..., 20:], d1)
    assert_cepstrum_close(cen[0], ref_cpp['mfcc_central_delta'])                 # the reference's own C++ output
    cen2 = scfeat.get_plan(delta='central2').extract_host(pcm)
    assert cen2.shape == (8, 30, 60) and np.array_equal(cen2[..., :40], cen)
    assert np.array_equal(cen2[..., 40:], (d1[:, nxt] - d1[:, prv]) / 2)
    # log-bank output, ragged clips without padding (rows of short clips stay zero), float input
    plan = scfeat.get_plan(output=scfeat.plan.OUT_LOG_BANK, delta='central')
    lengths = np.array([16000, 9000, 1024, 1023, 5000, 16000, 2048, 1536], dtype=np.int32)
    got = plan.extract_host(audio_of(pcm), lengths=lengths, pad=scfeat.plan.PAD_NONE)
    ref = scfeat.get_plan(output=scfeat.plan.OUT_LOG_BANK).extract_host(audio_of(pcm), lengths=lengths, pad=scfeat.plan.PAD_NONE)
    for i, n in enumerate(lengths):
        k = (n - 1024) // 512 + 1 if n >= 1024 else 0
        assert np.array_equal(got[i, :k, :20], ref[i, :k]) and not got[i, k:].any()
        if k:
            nx, pv = np.minimum(np.arange(k) + 1, k - 1), np.maximum(np.arange(k) - 1, 0)
            assert np.array_equal(got[i, :k, 20:], (ref[i, nx] - ref[i, pv]) / 2)
    with pytest.raises(scfeat.ScfError):
        scfeat.get_plan(output=scfeat.plan.OUT_POWER, delta='diff')


def test_stream_with_use_delta_returns_fresh_wide_rings(example_pcm):
    """Listener with pr.use_delta: the returned ring carries delta columns computed from its own base columns (the
    reference re-applies add_deltas to the widened ring every chunk, listen.py:111-112 -- not copied), and every call
    returns a fresh array (listen.py:96-114 builds a new one each time)."""
    _, pcm = example_pcm
    pr = scfeat.params.pr
    try:
        pr.__dict__.update(use_delta=True)
        one = scfeat.listener.Listener()
        o = opipe.ListenerOracle(opipe.Params())
        outs = []
        for s in range(0, 16000 - 1600 + 1, 1600):
            got = one.update_vectors(pcm[1, s:s + 1600].tobytes())
            base = o.update_vectors(pcm[1, s:s + 1600].tobytes())[..., 0]
            want = opipe.add_deltas(base)
            assert got.shape == (30, 40, 1)
            assert np.abs(got[..., 0] - want).max() <= CEP_REL * max(np.abs(want).max(), 1.0)
            outs.append(got)
        assert not np.array_equal(outs[-1], outs[-2]) and outs[0] is not outs[1]          # no aliasing between calls
    finally:
        pr.__dict__.update(use_delta=False)
    # empty audio: front-padded to max_samples like any short clip (common/data_utils.py:79-80) -> all-silence rows
    z = scfeat.data_utils.audio_to_feature(np.zeros(0, dtype=np.float32))
    assert z.shape == (30, 20)
    np.testing.assert_allclose(z[:, 0], -36.04365339, atol=2e-5)


def test_ragged_batch_fast_path_random_lengths(example_pcm):
    """Per-clip lengths with front padding run on the fast kernels (predicated loads only for the pair that straddles
    the padding): every possible alignment of the first valid sample inside a frame pair, int16 and float input."""
    _, pcm = example_pcm
    p = opipe.Params()
    rng = np.random.default_rng(21)
    n = 96
    clips = np.stack([pcm[i % 8] for i in range(n)])
    lengths = rng.integers(0, 16001, size=n).astype(np.int32)
    lengths[:8] = [16000, 0, 1, 15999, 16000 - 512, 16000 - 513, 16000 - 1024, 16000 - 1025]
    got = scfeat.data_utils.extract_features_batch(clips, lengths)[..., 0]
    want = np.stack([opipe.audio_to_feature(audio_of(clips[i][:lengths[i]]), p) if lengths[i] > 0
                     else opipe.audio_to_feature(np.zeros(16000, np.float32), p) for i in range(n)])
    assert_cepstrum_close(got, want)
    # same clips as float audio through the device API
    import torch
    plan = scfeat.get_plan()
    d_in = torch.from_numpy(audio_of(clips)).cuda()
    d_len = torch.from_numpy(lengths).cuda()
    d_out = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
    plan.extract_device(d_in.data_ptr(), n, 16000, d_out.data_ptr(), d_lengths=d_len.data_ptr(), is_f32=True)
    torch.cuda.synchronize()
    assert_cepstrum_close(d_out.cpu().numpy(), want)


def test_dynamic_tile_schedule_equals_round_robin(example_pcm):
    """Launches with more than three tiles per team draw their tiles from a global counter (the team that finishes
    early takes more): the rows must equal, bit for bit, what small launches (fixed round robin) give for the same clips
    -- for the cepstrum and the log-bank output, three times in a row (every launch takes a fresh counter word and
    leaves it zeroed), and for a ragged batch."""
    import torch
    _, pcm = example_pcm
    n = 6000                                            # 11,250 tiles on 444 teams
    rng = np.random.default_rng(11)
    shifts = rng.integers(0, 16000, size=n)
    clips = np.stack([np.roll(pcm[i % 8], shifts[i]) for i in range(n)])
    d_in = torch.from_numpy(clips).cuda()
    lengths = rng.integers(0, 16001, size=n).astype(np.int32)
    d_len = torch.from_numpy(lengths).cuda()
    for kw in (dict(), dict(output=scfeat.plan.OUT_LOG_BANK)):
        plan = scfeat.get_plan(**kw)
        small = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
        for a in range(0, n, 500):                      # 938 tiles per launch: round robin
            plan.extract_device(d_in[a].data_ptr(), min(500, n - a), 16000, small[a].data_ptr())
        torch.cuda.synchronize()
        for rep in range(3):
            big = torch.zeros((n, 30, 20), dtype=torch.float32, device='cuda')
            plan.extract_device(d_in.data_ptr(), n, 16000, big.data_ptr())
            torch.cuda.synchronize()
            assert torch.equal(big, small), (kw, rep)
        small_r = torch.empty((n, 30, 20), dtype=torch.float32, device='cuda')
        for a in range(0, n, 500):
            plan.extract_device(d_in[a].data_ptr(), min(500, n - a), 16000, small_r[a].data_ptr(), d_lengths=d_len[a:].data_ptr())
        big_r = torch.zeros((n, 30, 20), dtype=torch.float32, device='cuda')
        plan.extract_device(d_in.data_ptr(), n, 16000, big_r.data_ptr(), d_lengths=d_len.data_ptr())
        torch.cuda.synchronize()
        assert torch.equal(big_r, small_r), kw
    want = np.stack([osonopy.mfcc_spec(a, 16000, (1024, 512), 1024, 20, 20) for a in audio_of(clips[::997])])
    assert_cepstrum_close(scfeat.get_plan().extract_host(clips)[::997], want)


def test_randomised_self_consistency_stress():
    """tools/stress_schedule.py: random plans (FFT size, bank, filters, coefficients, output kind, front end) and job
    shapes; one big launch (dynamic tile schedule) equals small launches (round robin) bit for bit, streams match the
    batch features of their concatenated audio."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, 'tools', 'stress_schedule.py'), '12', '7'], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0 and 'STRESS_OK' in r.stdout, r.stdout[-3000:] + r.stderr[-2000:]


def test_one_plan_shared_by_threads_and_streams(example_pcm):
    """include/scfeat.h: plans are immutable and may be shared between threads.  Four threads drive the same plan on
    their own CUDA streams (device API) and through the host-buffer API at once; every result must equal the
    single-threaded one bit for bit."""
    import threading
    import torch
    _, pcm = example_pcm
    plan = scfeat.get_plan()
    rng = np.random.default_rng(5)
    batches = [rng.integers(-32768, 32768, size=(64 + 32 * t, 16000), dtype=np.int16) for t in range(4)]
    want = [plan.extract_host(b) for b in batches]
    got_dev, got_host, errs = [None] * 4, [None] * 4, []

    def work(t):
        try:
            st = torch.cuda.Stream()
            d_in = torch.from_numpy(batches[t]).cuda()
            d_out = torch.empty((len(batches[t]), 30, 20), dtype=torch.float32, device='cuda')
            for _ in range(20):
                plan.extract_device(d_in.data_ptr(), len(batches[t]), 16000, d_out.data_ptr(), stream=st.cuda_stream)
            st.synchronize()
            got_dev[t] = d_out.cpu().numpy()
            for _ in range(3):
                got_host[t] = plan.extract_host(batches[t])
        except Exception as e:      # surfaced below: an exception in a thread must fail the test
            errs.append(e)

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errs, errs
    for t in range(4):
        np.testing.assert_array_equal(got_dev[t], want[t])
        np.testing.assert_array_equal(got_host[t], want[t])
