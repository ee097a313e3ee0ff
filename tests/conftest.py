import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session', autouse=True)
def built_library():
    """libscfeat.so is a build artefact (git-ignored): compile it with nvcc when a fresh checkout runs the tests
    (nvcc cross-compiles sm_100a without a GPU).  On the GPU box the prebuilt file travels with the snapshot."""
    import scfeat
    if not os.path.exists(scfeat._lib.LIB_PATH):
        scfeat.build()
    return scfeat._lib.LIB_PATH


@pytest.fixture(scope='session')
def example_pcm():
    z = np.load(os.path.join(GOLDEN, 'example_pcm.npz'))
    return [str(n) for n in z['names']], z['pcm']


@pytest.fixture(scope='session')
def ref_bark():
    return np.load(os.path.join(GOLDEN, 'ref_bark.npz'))


@pytest.fixture(scope='session')
def ref_cpp():
    return np.load(os.path.join(GOLDEN, 'ref_mfcc_cpp.npz'))


@pytest.fixture(scope='session')
def oracle_pin():
    return np.load(os.path.join(GOLDEN, 'oracle_mfcc.npz'))
