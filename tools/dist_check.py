#!/usr/bin/env python3
"""Multi-GPU check + timing of the fused extract/all-gather (config 3 shape), one process per GPU:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
      tools/dist_check.py [--clips 105829] [--reps 5]

Every rank extracts its shard and the kernel epilogue stores the rows into all ranks' caches over NVLink (CUDA IPC
peer mappings).  Checks that every rank's cache equals the single-GPU result, then times (max over ranks, CUDA
events) extract-only, fused extract+gather, and extract + NCCL all-gather."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import scfeat
from scfeat.dist import FeatureCacheGather, shard_range


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--clips', type=int, default=105829)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--multicast', action='store_true', help='one multimem.st per row segment to the multicast address '
                    'of a torch symmetric-memory cache instead of world - 1 peer stores')
    args = ap.parse_args()
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    plan = scfeat.get_plan(device=local)
    n = args.clips
    start, count, per = shard_range(n, world, rank)
    g = torch.Generator(device='cuda')
    g.manual_seed(1000 + rank)
    d_pcm = torch.randint(-32768, 32768, (max(count, 1), 16000), dtype=torch.int16, device='cuda', generator=g)
    cache = FeatureCacheGather(plan, n, 16000, world, rank, local, group=dist.group.WORLD, multicast=args.multicast)
    st = torch.cuda.current_stream()
    dist.barrier()
    cache.extract_and_gather(d_pcm.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    dist.barrier()
    # --- correctness: my cache's block of every rank r equals what rank r computes locally --------------
    local_out = torch.empty((max(count, 1), 30, 20), dtype=torch.float32, device='cuda')
    plan.extract_device(d_pcm.data_ptr(), count, 16000, local_out.data_ptr(), stream=st.cuda_stream)
    torch.cuda.synchronize()
    blocks = [torch.zeros((per, 30, 20), dtype=torch.float32, device='cuda') for _ in range(world)]
    mine = torch.zeros((per, 30, 20), dtype=torch.float32, device='cuda')
    mine[:count] = local_out[:count]
    dist.all_gather(blocks, mine)
    want = torch.cat(blocks)[:n].cpu().numpy()
    got = cache.to_host()
    ok = bool(np.array_equal(got, want)) and bool(np.isfinite(got).all())
    flag = torch.tensor([1 if ok else 0], device='cuda')
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)

    def timed(fn):
        best = None
        for _ in range(args.reps):
            dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            t = torch.tensor([e0.elapsed_time(e1)], device='cuda')
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            best = float(t) if best is None else min(best, float(t))
        return best

    full = torch.empty((world * per, 30, 20), dtype=torch.float32, device='cuda')
    t_extract = timed(lambda: plan.extract_device(d_pcm.data_ptr(), count, 16000, local_out.data_ptr(), stream=st.cuda_stream))
    t_fused = timed(lambda: cache.extract_and_gather(d_pcm.data_ptr(), stream=st.cuda_stream))

    def nccl_path():
        plan.extract_device(d_pcm.data_ptr(), count, 16000, mine.data_ptr(), stream=st.cuda_stream)
        dist.all_gather_into_tensor(full, mine)
    t_nccl = timed(nccl_path)
    if rank == 0:
        print(json.dumps({'world': world, 'clips': n, 'per_rank': per, 'ok': bool(int(flag)), 'multicast': bool(args.multicast),
                          'extract_only_ms': t_extract, 'fused_extract_gather_ms': t_fused, 'extract_plus_nccl_allgather_ms': t_nccl,
                          'fused_clips_per_s': n / (t_fused * 1e-3), 'extract_only_clips_per_s': n / (t_extract * 1e-3)}))
        print('DIST_CHECK_OK' if int(flag) else 'DIST_CHECK_FAILED')
    cache.close()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag) else 1)


if __name__ == '__main__':
    main()
