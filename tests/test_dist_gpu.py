"""Fused extract + all-gather (scf_extract_i16_gather).  On the single-GPU test box several ranks are emulated by
sequential launches on one device (the kernel never waits on another rank: it only stores finished rows through the
peer table), with one cache per emulated rank.  A real 2-GPU run over CUDA IPC / NVLink is tools/dist_check.py."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

import scfeat
from oracle import sonopy as osonopy
from scfeat import _lib
from scfeat.dist import FeatureCacheGather, shard_range

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize('n_clips,world', [(37, 2), (100, 3), (8, 8)])
def test_emulated_ranks_fill_every_cache(example_pcm, n_clips, world):
    import torch
    rng = np.random.default_rng(5)
    pcm = rng.integers(-32768, 32768, size=(n_clips, 16000), dtype=np.int16)
    pcm[:8] = example_pcm[1]
    plan = scfeat.get_plan()
    L = _lib.lib()
    per = -(-n_clips // world)
    rows = world * per
    nbytes = rows * 30 * 20 * 4
    caches = []
    for _ in range(world):
        p = ctypes.c_void_p()
        _lib.check(L.scf_device_malloc(0, nbytes, ctypes.byref(p)))
        caches.append(p)
    table = (ctypes.c_void_p * world)(*[c.value for c in caches])
    d_pcm = torch.from_numpy(pcm).cuda()
    st = torch.cuda.current_stream()
    for r in range(world):
        start, count, _ = shard_range(n_clips, world, r)
        if count:
            _lib.check(L.scf_extract_i16_gather(plan.handle, d_pcm[start].data_ptr(), count, 16000, 16000, table, world, r,
                                                per, st.cuda_stream))
    st.synchronize()
    want = plan.extract_host(pcm)
    ref = np.stack([osonopy.mfcc_spec(c.astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20, 20) for c in pcm[:8]])
    assert np.abs(want[:8] - ref).max() <= 1e-4 * np.abs(ref).max()
    for c in caches:
        got = np.empty((rows, 30, 20), dtype=np.float32)
        _lib.check(L.scf_memcpy(0, got.ctypes.data, c, nbytes, 1, None))
        assert np.array_equal(got[:n_clips], want)            # every rank holds all rows, in clip order
        _lib.check(L.scf_device_free(0, c))


def test_feature_cache_gather_single_process(example_pcm):
    import torch
    plan = scfeat.get_plan()
    pcm = np.tile(example_pcm[1], (3, 1))
    g = FeatureCacheGather(plan, len(pcm), 16000, 1, 0, 0)
    d = torch.from_numpy(pcm).cuda()
    g.extract_and_gather(d.data_ptr(), stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    want = plan.extract_host(pcm)
    assert np.array_equal(g.to_host(), want)
    # the gathered cache handed to the framework through DLPack: [N, 30, 20, 1] on the device, kept alive by the tensor
    t = torch.from_dlpack(g.to_dlpack())
    assert t.is_cuda and tuple(t.shape) == (len(pcm), 30, 20, 1) and t.data_ptr() == g.ptr
    del g
    import gc
    gc.collect()
    assert np.array_equal(t.cpu().numpy()[..., 0], want)          # still valid: the capsule owns a reference to the cache
    del t


def _dist_check(nproc, port, *extra):
    return subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node=%d' % nproc,
                           '--master-addr', '127.0.0.1', '--master-port', str(port), os.path.join(ROOT, 'tools', 'dist_check.py'),
                           '--clips', '4001'] + list(extra), capture_output=True, text=True, timeout=600)


def test_two_gpus_over_ipc_when_available():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs (run tools/dist_check.py under gpurun --gpus 2)')
    r = _dist_check(2, 29517)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DIST_CHECK_OK' in r.stdout


def test_multicast_gather_when_available():
    """scf_extract_i16_gather_multicast: one multimem.st per row segment to the multicast address of a torch
    symmetric-memory cache.  Runs with every visible GPU (one process each, at most 2); a fabric or torch build without a
    multicast mapping skips (the peer-store path above is the portable one)."""
    import torch
    n = min(torch.cuda.device_count(), 2)
    r = _dist_check(n, 29518, '--multicast')
    if r.returncode != 0 and ('no multicast address' in r.stdout + r.stderr or 'multicast' in (r.stdout + r.stderr).lower()
                              and 'DIST_CHECK_FAILED' not in r.stdout):
        pytest.skip('no multicast mapping on this box: ' + (r.stdout + r.stderr)[-300:].replace('\n', ' '))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert 'DIST_CHECK_OK' in r.stdout and '"multicast": true' in r.stdout
