"""Pins the CPU oracle against everything the reference offers for this path (CPU-only tests).

The reference has no tests and no golden vectors, so the pins are:
  * outputs of the unmodified common/bark_feature.py (tests/golden/ref_bark.npz),
  * outputs of the compiled in-tree C++ twin inference/tflite/mfcc.h (tests/golden/ref_mfcc_cpp.npz),
  * the known-answer spot values recorded at survey time.
"""
import numpy as np
import pytest

from oracle import bark, pipeline, sonopy


def audio_of(pcm):
    return pcm.astype(np.float32) / 32768.0


# ---------------------------------------------------------------- bark_feature.py (real file)
@pytest.mark.parametrize('nf,nfft,scale', [(20, 512, 'constant'), (20, 1024, 'constant'), (24, 512, 'constant'),
                                           (24, 1024, 'constant'), (26, 512, 'constant'), (26, 1024, 'constant'),
                                           (22, 512, 'ascendant'), (22, 512, 'descendant')])
def test_bark_bank_matches_reference_file(ref_bark, nf, nfft, scale):
    want = ref_bark['bank_%d_%d_%s' % (nf, nfft, scale)]
    got = bark.bark_filterbanks(nfilts=nf, nfft=nfft, sample_rate=16000, scale=scale)
    assert got.shape == want.shape == (nf, nfft // 2 + 1)
    assert np.array_equal(got != 0, want != 0)
    np.testing.assert_allclose(got, want, rtol=1e-13, atol=0)


def test_bark_bank_known_answers():
    b = bark.bark_filterbanks(nfilts=20, nfft=1024)
    assert b.shape == (20, 513)
    assert list((b != 0).sum(1)) == [11, 12, 12, 13, 14, 16, 17, 20, 21, 24, 28, 32, 37, 41, 48, 55, 63, 72, 83, 96]
    assert np.flatnonzero((b != 0).any(0)).max() == 239
    for nf, nnz in ((20, 715), (24, 811), (26, 852)):
        for nfft in (512, 1024):
            assert int((bark.bark_filterbanks(nfilts=nf, nfft=nfft) != 0).sum()) == nnz


def test_power_spec_matches_reference_file(example_pcm, ref_bark):
    _, pcm = example_pcm
    a = audio_of(pcm)
    got = np.stack([sonopy.power_spec(x, (1024, 512), 1024) for x in a[:2]])
    np.testing.assert_allclose(got, ref_bark['power_1024_512_1024'], rtol=1e-12, atol=1e-18)
    np.testing.assert_allclose(sonopy.power_spec(a[0], (160, 80), 512), ref_bark['power_160_80_512'],
                               rtol=1e-12, atol=1e-18)
    # window > fft_size: np.fft.rfft(n=) crops the frame
    np.testing.assert_allclose(sonopy.power_spec(a[0], (1200, 400), 1024), ref_bark['power_1200_400_1024'],
                               rtol=1e-12, atol=1e-18)


@pytest.mark.parametrize('key,args', [('bfcc_1024_512_1024_20_20', (1024, 512, 1024, 20, 20)),
                                      ('bfcc_1024_512_1024_26_13', (1024, 512, 1024, 26, 13)),
                                      ('bfcc_512_256_512_26_13', (512, 256, 512, 26, 13))])
def test_bfcc_matches_reference_file(example_pcm, ref_bark, key, args):
    _, pcm = example_pcm
    got = np.stack([bark.bfcc_spec(x, 16000, *args) for x in audio_of(pcm)])
    np.testing.assert_allclose(got, ref_bark[key], rtol=0, atol=1e-10)


@pytest.mark.parametrize('key,args', [('bark_1024_512_1024_24', (1024, 512, 1024, 24)),
                                      ('bark_1024_512_1024_20', (1024, 512, 1024, 20)),
                                      ('bark_512_256_512_24', (512, 256, 512, 24))])
def test_bark_spec_matches_reference_file(example_pcm, ref_bark, key, args):
    _, pcm = example_pcm
    got = np.stack([bark.bark_spec(x, 16000, *args) for x in audio_of(pcm)])
    np.testing.assert_allclose(got, ref_bark[key], rtol=0, atol=1e-10)


def test_bark_long_and_synthetic(ref_bark):
    lf = audio_of(ref_bark['long_pcm'])
    np.testing.assert_allclose(bark.bfcc_spec(lf, 16000, 1024, 512, 1024, 26, 13),
                               ref_bark['long_bfcc_1024_512_1024_26_13'], rtol=0, atol=1e-10)
    np.testing.assert_allclose(bark.bark_spec(lf, 16000, 1024, 512, 1024, 24),
                               ref_bark['long_bark_1024_512_1024_24'], rtol=0, atol=1e-10)
    got = np.stack([bark.bfcc_spec(x, 16000, 1024, 512, 1024, 26, 13) for x in audio_of(ref_bark['synth_pcm'])])
    np.testing.assert_allclose(got, ref_bark['synth_bfcc_1024_512_1024_26_13'], rtol=0, atol=1e-10)


def test_bark_known_answers(example_pcm):
    names, pcm = example_pcm
    a = audio_of(pcm[names.index('right_1')])
    c = bark.bfcc_spec(a, 16000, 1024, 512, fft_size=1024, num_filt=20, num_coeffs=20)
    np.testing.assert_allclose(c[0, :5], [-1.71329153, 0.21673045, 3.70136894, 3.03721111, 2.35450541], atol=5e-8)
    s = bark.bark_spec(a, 16000, 1024, 512, fft_size=1024, num_filt=20)
    assert abs(s.min() - -8.58198) < 1e-5 and abs(s.max() - 4.17898) < 1e-5
    assert bark.bfcc_spec(a[:100], 16000, 1024, 512).shape == (0, 13)


# ---------------------------------------------------------------- sonopy restatement vs mfcc.h
def test_mel_grid_known_answer():
    assert sonopy.mel_grid(16000, 20, 513) == [0, 3, 7, 12, 18, 25, 33, 42, 52, 64, 79, 95, 115, 137, 163, 193,
                                               229, 270, 317, 373, 437, 513]
    assert int((sonopy.filterbanks(16000, 20, 513) != 0).sum()) == 927         # SURVEY.md section 8 a6 (each rising edge starts at exactly 0)


def test_mel_bank_matches_cpp_twin(ref_cpp):
    np.testing.assert_allclose(sonopy.filterbanks(16000, 20, 513), ref_cpp['bank_16000_20_1024'], rtol=1e-14, atol=0)


def test_mel_bank_with_repeated_grid_points_matches_cpp_twin(ref_cpp):
    # n_fft 512 / 40 filters: the raw grid starts 0, 0, 1, 2, 2, ... -- the twin (mfcc.h:230-264) keeps the repeats,
    # so some filters are empty or one-sided; a de-duplicating restatement differs in 20 of the 40 rows
    # (n_fft 256 / 40 filters starts 0, 0, 0, 1, 2, 2: filter 0 is empty altogether)
    for bins, key, n_empty in ((257, 'bank_16000_40_512', 0), (129, 'bank_16000_40_256', 1)):
        grid = sonopy.mel_grid(16000, 40, bins)
        assert len(set(grid)) < len(grid)
        want = ref_cpp[key]
        got = sonopy.filterbanks(16000, 40, bins)
        assert np.array_equal(got != 0, want != 0) and int((want.sum(axis=1) == 0).sum()) >= n_empty
        np.testing.assert_allclose(got, want, rtol=1e-14, atol=0)


def test_mfcc_matches_cpp_twin(example_pcm, ref_cpp, ref_bark):
    _, pcm = example_pcm
    got = np.stack([sonopy.mfcc_spec(x, 16000, (1024, 512), 1024, 20, 20) for x in audio_of(pcm)])
    assert got.shape == (8, 30, 20)
    assert np.abs(got - ref_cpp['mfcc_params_json']).max() < 1e-6        # C++ rounds its output to fp32
    got = np.stack([sonopy.mfcc_spec(x, 16000, (1024, 512), 1024, 20, 20) for x in audio_of(ref_bark['synth_pcm'])])
    assert np.abs(got - ref_cpp['mfcc_params_json_synth']).max() < 1e-6
    got = np.stack([sonopy.mfcc_spec(x, 16000, (512, 256), 512, 20, 13) for x in audio_of(pcm[:2])])
    assert np.abs(got - ref_cpp['mfcc_512_256_512_20_13']).max() < 1e-6
    # repeated grid points: empty filters give log(eps) = -36.04 in both
    got = np.stack([sonopy.mfcc_spec(x, 16000, (512, 256), 512, 40, 13) for x in audio_of(pcm[:2])])
    assert np.abs(got - ref_cpp['mfcc_512_256_512_40_13_dupgrid']).max() < 5e-6       # |values| reach 60: fp32 output rounding
    got = np.stack([sonopy.mfcc_spec(x, 16000, (256, 128), 256, 40, 13) for x in audio_of(pcm[:2])])
    assert np.abs(got - ref_cpp['mfcc_256_128_256_40_13_dupgrid']).max() < 5e-6


def test_mfcc_known_answers(example_pcm):
    names, pcm = example_pcm
    m = sonopy.mfcc_spec(audio_of(pcm[names.index('right_1')]), 16000, (1024, 512), 1024, 20, 20)
    np.testing.assert_allclose(m[0, :5], [-1.71329153, 1.68849206, 2.06822109, 3.96351266, 2.23058176], atol=6e-7)
    np.testing.assert_allclose(m[1, :5], [-1.67023587, 1.37456262, 1.89936841, 4.11730099, 2.49116015], atol=6e-7)
    # c0 is the log frame energy for both banks (bark_feature.py:173)
    b = bark.bfcc_spec(audio_of(pcm[names.index('right_1')]), 16000, 1024, 512, 1024, 20, 20)
    np.testing.assert_allclose(m[:, 0], b[:, 0], rtol=0, atol=1e-12)


def test_oracle_regression_pin(example_pcm, oracle_pin):
    _, pcm = example_pcm
    a = audio_of(pcm)
    got = np.stack([sonopy.mfcc_spec(x, 16000, (1024, 512), 1024, 20, 20) for x in a])
    np.testing.assert_allclose(got, oracle_pin['mfcc_params_json'], rtol=0, atol=1e-10)
    got = np.stack([sonopy.mel_spec(x, 16000, (1024, 512), 1024, 20) for x in a])
    np.testing.assert_allclose(got, oracle_pin['mel_params_json'], rtol=0, atol=1e-10)


def test_dct_matrix_equals_scipy():
    from scipy.fftpack import dct
    x = np.random.default_rng(0).normal(size=(7, 26))
    np.testing.assert_allclose(x @ sonopy.dct2_ortho_matrix(26, 13), dct(x, norm='ortho')[:, :13], atol=1e-12)
    np.testing.assert_allclose(x[:, :20] @ sonopy.dct2_ortho_matrix(20, 40), dct(x[:, :20], norm='ortho'), atol=1e-12)


def test_all_zero_clip_and_short_input():
    z = sonopy.mfcc_spec(np.zeros(16000, np.float32), 16000, (1024, 512), 1024, 20, 20)
    assert z.shape == (30, 20)
    np.testing.assert_allclose(z[:, 0], -36.04365339, atol=1e-7)
    assert np.abs(z[:, 1:]).max() < 1e-12
    assert sonopy.mfcc_spec(np.zeros(1023), 16000, (1024, 512), 1024, 20, 20).shape == (0, 20)
    assert sonopy.mfcc_spec(np.zeros(1023), 16000, (1024, 512), 1024, 20, 30).shape == (0, 20)
    assert sonopy.mfcc_spec(np.ones(2048), 16000, (1024, 512), 1024, 20, 30).shape == (3, 20)


# ---------------------------------------------------------------- callers: params, pad/crop, stream
def test_params_derived_sizes():
    p = pipeline.Params()
    assert (p.window_samples, p.hop_samples, p.max_samples, p.buffer_samples) == (1024, 512, 16000, 15872)
    assert (p.n_features, p.feature_size) == (30, 20)


def test_audio_to_feature_crop_and_front_pad(example_pcm):
    _, pcm = example_pcm
    p = pipeline.Params()
    a = audio_of(pcm[0])
    full = pipeline.audio_to_feature(a, p)
    assert full.shape == (30, 20)
    longer = np.concatenate([a, a[:5000]])
    np.testing.assert_array_equal(pipeline.audio_to_feature(longer, p), full)       # keeps the head
    short = pipeline.audio_to_feature(a[:9000], p)
    padded = np.concatenate([np.zeros(7000), a[:9000]])
    np.testing.assert_array_equal(short, pipeline.vectorize_raw(padded, p))          # pads in front
    with pytest.raises(ValueError):
        pipeline.vectorize_raw(np.zeros(0), p)


@pytest.mark.parametrize('chunk', [1600, 1024, 512, 3000])
def test_stream_oracle_reproduces_batch_features(example_pcm, chunk):
    _, pcm = example_pcm
    p = pipeline.Params()
    x = np.concatenate([pcm[0], pcm[1]])
    lo = pipeline.ListenerOracle(p)
    emitted = []
    for s in range(0, len(x) - chunk + 1, chunk):
        before = len(lo.window_audio)
        out = lo.update_vectors(x[s:s + chunk].tobytes())
        assert out.shape == (30, 20, 1)
        emitted.append((before + chunk - len(lo.window_audio)) // p.hop_samples)
    n_done = sum(emitted)
    batch = sonopy.mfcc_spec(audio_of(x), 16000, (1024, 512), 1024, 20, 20)
    np.testing.assert_allclose(lo.mfccs, batch[n_done - 30:n_done], rtol=0, atol=2e-7)
    if chunk == 1600:
        assert emitted[:10] == [2, 3, 3, 3, 3, 3, 3, 4, 3, 3]
    if chunk == 1024:
        assert emitted[:4] == [1, 2, 2, 2]
