"""World-size-2/3 gloo jobs on the CPU covering the host logic of the multi-GPU feature-cache path
(sharding, padding, rank-major layout, trimming) -- see tests/dist_worker_cpu.py."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import sonopy as osonopy
from scfeat.dist import shard_range

HERE = os.path.dirname(os.path.abspath(__file__))


def free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


@pytest.mark.parametrize('n,world', [(105829, 8), (512, 8), (7, 2), (1, 4), (13229, 1), (9, 4), (0, 3)])
def test_shard_range_partitions_the_clips(n, world):
    per = -(-n // world)
    seen = []
    for r in range(world):
        start, count, per_rank = shard_range(n, world, r)
        assert per_rank == per and 0 <= count <= per_rank
        seen += list(range(start, start + count))
        if count:
            assert start == r * per_rank
    assert seen == list(range(n))             # rank-major blocks of per_rank rows: the gathered cache is in clip order
    if n == 105829 and world == 8:
        assert per == 13229


@pytest.mark.parametrize('n_clips,world', [(7, 2), (5, 3)])
def test_gloo_allgather_assembles_the_cache(tmp_path, n_clips, world):
    port = free_port()
    procs = []
    for r in range(world):
        env = dict(os.environ, RANK=str(r), WORLD_SIZE=str(world), MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port),
                   OMP_NUM_THREADS='1')
        procs.append(subprocess.Popen([sys.executable, os.path.join(HERE, 'dist_worker_cpu.py'), str(tmp_path), str(n_clips)],
                                      env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT))
    for p in procs:
        out, _ = p.communicate(timeout=240)
        assert p.returncode == 0, out.decode()[-2000:]
    rng = np.random.default_rng(123)
    pcm = rng.integers(-32768, 32768, size=(n_clips, 4096), dtype=np.int16)
    want = np.stack([osonopy.mfcc_spec(c.astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20, 20) for c in pcm])
    for r in range(world):
        got = np.load(tmp_path / ('rank%d.npy' % r))
        assert got.shape == (n_clips, 7, 20)
        np.testing.assert_allclose(got, want.astype(np.float32), rtol=0, atol=0)
