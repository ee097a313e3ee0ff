"""ThresholdDecoder / TriggerDetector mirrors against outputs of the reference's own classes
(extracted verbatim from listen.py by tests/golden/make_golden.py -> tests/golden/ref_postprocess.npz)."""
import os

import numpy as np
import pytest

from scfeat.postprocess import BatchTriggerDetector, ThresholdDecoder, TriggerDetector

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'ref_postprocess.npz'))
CFGS = {'default': (((6, 4),), 0.2), 'two': (((6, 4), (-2, 3)), 0.5), 'narrow': (((0, 0.1),), 0.5)}


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
    classes = ['background', 'up', 'down', 'left']
    idx, score = G['trigger_idx'], G['trigger_score']
    det = TriggerDetector(chunk, classes, 0.5, 3)
    got = np.array([det.update(int(i), float(s)) for i, s in zip(idx, score)])
    assert np.array_equal(got, G['trigger_%d' % chunk])
    assert got.sum() > 0
    # batch form: stream j runs the same sequence delayed by j steps
    n = 5
    b = BatchTriggerDetector(n, chunk, classes, 0.5, 3)
    fired = np.zeros((len(idx) + n, n), dtype=bool)
    for t in range(len(idx) + n):
        ii = np.array([idx[t - j] if 0 <= t - j < len(idx) else 0 for j in range(n)])
        ss = np.array([score[t - j] if 0 <= t - j < len(idx) else 0.0 for j in range(n)])
        fired[t] = b.update(ii, ss)
    for j in range(n):
        assert np.array_equal(fired[j:j + len(idx), j], G['trigger_%d' % chunk])
