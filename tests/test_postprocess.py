"""Streaming post-processing: ThresholdDecoder / TriggerDetector.

CPU part: the float64 numpy restatement in oracle/postprocess.py against outputs of the reference's own classes
(extracted verbatim from listen.py by tests/golden/make_golden.py -> tests/golden/ref_postprocess.npz), and the
library's host-side table builder (scf_post_build_cd) against it.
GPU part (-m gpu): the device implementation behind scf_post_* (csrc/scfeat_post.cu) against the same fixtures --
decode within 1e-12, trigger decisions bit-exact -- and a 256-stream run of the fused step against the oracle."""
import os

import numpy as np
import pytest

import scfeat
from oracle.postprocess import BatchTriggerDetector, ThresholdDecoder, TriggerDetector

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_postprocess.npz'))
CFGS = {'default': (((6, 4),), 0.2), 'two': (((6, 4), (-2, 3)), 0.5), 'narrow': (((0, 0.1),), 0.5)}
CLASSES = ['background', 'up', 'down', 'left']


# ------------------------------------------------------------------------------------------------ CPU: the oracle
@pytest.mark.parametrize('name', sorted(CFGS))
def test_threshold_decoder_matches_reference(name):
    cfg, center = CFGS[name]
    d = ThresholdDecoder(cfg, center)
    np.testing.assert_allclose(d.decode_batch(G['raw']), G['decode_' + name], rtol=0, atol=1e-15)
    assert d.decode(0.0) == 0.0 and d.decode(1.0) == 1.0
    for r, w in zip(G['raw'][:20], G['decode_' + name][:20]):
        assert d.decode(float(r)) == pytest.approx(w, abs=1e-15)
    got = np.array([d.encode(float(t)) for t in np.linspace(0.01, 0.99, 50)])
    np.testing.assert_allclose(got, G['encode_' + name], rtol=0, atol=1e-15)


@pytest.mark.parametrize('chunk', [1024, 1600, 4096])
def test_trigger_detector_matches_reference(chunk):
    idx, score = G['trigger_idx'], G['trigger_score']
    det = TriggerDetector(chunk, CLASSES, 0.5, 3)
    got = np.array([det.update(int(i), float(s)) for i, s in zip(idx, score)])
    assert np.array_equal(got, G['trigger_%d' % chunk])
    assert got.sum() > 0
    # batch form: stream j runs the same sequence delayed by j steps
    n = 5
    b = BatchTriggerDetector(n, chunk, CLASSES, 0.5, 3)
    fired = np.zeros((len(idx) + n, n), dtype=bool)
    for t in range(len(idx) + n):
        ii = np.array([idx[t - j] if 0 <= t - j < len(idx) else 0 for j in range(n)])
        ss = np.array([score[t - j] if 0 <= t - j < len(idx) else 0.0 for j in range(n)])
        fired[t] = b.update(ii, ss)
    for j in range(n):
        assert np.array_equal(fired[j:j + len(idx), j], G['trigger_%d' % chunk])


@pytest.mark.parametrize('name', sorted(CFGS))
def test_library_decoder_table_equals_oracle(name):
    """scf_post_build_cd (host arithmetic, no GPU): min_out / max_out exact, cumulative table within a few ulps."""
    cfg, center = CFGS[name]
    d = ThresholdDecoder(cfg, center)
    lo, hi, cd = scfeat.postprocess.build_cd(cfg)
    assert (lo, hi) == (d.min_out, d.max_out) and cd.shape == d.cd.shape
    np.testing.assert_allclose(cd, d.cd, rtol=1e-13, atol=1e-15)


# ------------------------------------------------------------------------------------------------ GPU: scf_post_*
@pytest.mark.gpu
@pytest.mark.parametrize('name', sorted(CFGS))
def test_device_decode_matches_reference(name):
    cfg, center = CFGS[name]
    d = scfeat.postprocess.ThresholdDecoder(cfg, center)
    got = d.decode_batch(G['raw'])
    assert np.abs(got - G['decode_' + name]).max() <= 1e-12
    assert d.decode(0.0) == 0.0 and d.decode(1.0) == 1.0
    assert abs(d.decode(float(G['raw'][11])) - G['decode_' + name][11]) <= 1e-12
    enc = np.array([d.encode(float(t)) for t in np.linspace(0.01, 0.99, 50)])
    np.testing.assert_allclose(enc, G['encode_' + name], rtol=0, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize('chunk', [1024, 1600, 4096])
def test_device_trigger_detector_matches_reference(chunk):
    idx, score = G['trigger_idx'], G['trigger_score']
    det = scfeat.postprocess.TriggerDetector(chunk, CLASSES, 0.5, 3)
    got = np.array([det.update(int(i), float(s)) for i, s in zip(idx, score)])
    assert np.array_equal(got, G['trigger_%d' % chunk])
    # 256 streams at once, stream j delayed by j steps: every column reproduces the reference sequence
    n = 256
    b = scfeat.postprocess.BatchTriggerDetector(n, chunk, CLASSES, 0.5, 3)
    T = len(idx)
    fired = np.zeros((T + n, n), dtype=bool)
    jj = np.arange(n)
    for t in range(T + n):
        k = t - jj
        ok = (k >= 0) & (k < T)
        ii = np.where(ok, idx[np.clip(k, 0, T - 1)], 0)
        ss = np.where(ok, score[np.clip(k, 0, T - 1)], 0.0)
        fired[t] = b.update(ii, ss)
    for j in (0, 1, 17, 255):
        assert np.array_equal(fired[j:j + T, j], G['trigger_%d' % chunk])


@pytest.mark.gpu
def test_device_fused_step_256_streams_vs_oracle():
    """scf_post_step: arg-max + max + decode of non-background scores + trigger update for 256 streams in one launch,
    against the listen.py:411-425 loop written with the oracle classes (float32 model scores, as in the reference)."""
    rng = np.random.default_rng(5)
    n, T, chunk = 256, 300, 1600
    cfg, center = CFGS['default']
    post = scfeat.postprocess.PostProcessor(n, CLASSES, cfg, center, chunk_size=chunk, sensitivity=0.5, trigger_level=3)
    dec = ThresholdDecoder(cfg, center)
    dets = [TriggerDetector(chunk, CLASSES, 0.5, 3) for _ in range(n)]
    # softmax-like scores with long runs of one confident class per stream, so that triggers do fire
    logits = rng.normal(0, 1, size=(T, n, len(CLASSES))).astype(np.float32)
    for s in range(n):
        for start in rng.integers(0, T - 40, size=4):
            logits[start:start + rng.integers(3, 40), s, rng.integers(0, len(CLASSES))] += 9.0
    e = np.exp(logits - logits.max(axis=-1, keepdims=True))
    probs = (e / e.sum(axis=-1, keepdims=True)).astype(np.float32)
    probs[5, :, :] = np.float32(0.25)                     # ties: arg-max takes the first
    probs[6, :3] = [[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, 0, 1]]          # exact 0 / 1 scores pass through decode
    n_fired = 0
    for t in range(T):
        idx, score, fired = post.step(probs[t])
        want_idx = probs[t].argmax(axis=-1)
        assert np.array_equal(idx, want_idx)
        for s in range(n):
            sc = probs[t, s, want_idx[s]]                 # numpy float32 scalar, as np.max(output) gives the reference
            if CLASSES[want_idx[s]] != 'background':
                # the reference's asigmoid evaluates 1 / x - 1 in float32 on the model's float32 score (listen.py:484)
                x = float(sc)
                if x == 1.0 or x == 0.0:
                    w = x
                else:
                    arg = np.float32(1) / sc - np.float32(1)
                    logit = -np.log(np.float64(arg)) if 0 < x < 1 else -10.0
                    ratio = min(max((logit - dec.min_out) / dec.out_range, 0.0), 1.0)
                    cp = dec.cd[int(ratio * (len(dec.cd) - 1) + 0.5)]
                    w = 0.5 * cp / dec.center if cp < dec.center else 0.5 + 0.5 * (cp - dec.center) / (1 - dec.center)
            else:
                w = float(sc)
            assert abs(score[s] - w) <= 1e-12, (t, s)
            f = dets[s].update(int(want_idx[s]), w)
            assert bool(fired[s]) == f, (t, s)
            n_fired += f
    assert n_fired > 50
    act, rec = post.state()
    assert np.array_equal(act, [d.activation for d in dets])
    post.reset()
    assert not post.state()[0].any() and (post.state()[1] == -1).all()
