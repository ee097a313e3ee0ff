"""Worker for tests/test_dist_cpu.py: one rank of a world-size-2 (or more) gloo job on the CPU.

Host-side logic of the multi-GPU path under test: shard arithmetic (scfeat.dist.shard_range), padding of the last
rank, rank-major cache layout and trimming.  The per-clip features come from the CPU oracle here (there is no GPU);
the all-gather is gloo's.  Writes <out_dir>/rank<r>.npy with the assembled cache of this rank."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import sonopy as osonopy  # noqa: E402
from scfeat.dist import shard_range  # noqa: E402


def main():
    out_dir, n_clips = sys.argv[1], int(sys.argv[2])
    rank, world = int(os.environ['RANK']), int(os.environ['WORLD_SIZE'])
    dist.init_process_group('gloo', rank=rank, world_size=world)
    rng = np.random.default_rng(123)
    pcm = rng.integers(-32768, 32768, size=(n_clips, 4096), dtype=np.int16)     # 7 frames per clip
    start, count, per_rank = shard_range(n_clips, world, rank)
    local = np.zeros((per_rank, 7, 20), dtype=np.float32)
    for i in range(count):
        local[i] = osonopy.mfcc_spec(pcm[start + i].astype(np.float32) / 32768.0, 16000, (1024, 512), 1024, 20, 20)
    full = torch.zeros((world * per_rank, 7, 20), dtype=torch.float32)
    dist.all_gather_into_tensor(full, torch.from_numpy(local))
    np.save(os.path.join(out_dir, 'rank%d.npy' % rank), full.numpy()[:n_clips])
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
