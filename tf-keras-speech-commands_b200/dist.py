"""Multi-GPU feature-cache assembly (the reference itself is single-process).

Every clip is independent, so ranks own contiguous clip ranges and there is no collective during
extraction.  What the training loop needs (classifier/data.py:101-114 loads ALL features into memory on the
process that calls model.fit) is the full [N, frames, cols] tensor on every rank: FeatureCacheGather gives
every rank a peer-mapped view of every rank's cache (CUDA IPC over NVLink) and the extraction kernel's epilogue
stores each finished row into all of them (scf_extract_i16_gather) -- the all-gather is fused into the compute
kernel.  torch.distributed is used only to exchange the 64-byte IPC handles and for the closing barrier.

With ``multicast=True`` the cache is allocated as torch symmetric memory and the epilogue issues ONE ``multimem.st`` per
row segment to the multicast address of all ranks' caches (scf_extract_i16_gather_multicast): the NVSwitch replicates
the row, so it leaves the GPU once instead of world - 1 times.
"""
import ctypes

import numpy as np

from . import _lib
from ._lib import check


def shard_range(n_clips, world, rank):
    """Contiguous shard of rank `rank`: (start, count, per_rank) with per_rank = ceil(n/world); the last
    ranks may own fewer (or zero) real clips and are padded to per_rank rows inside the cache."""
    per_rank = -(-int(n_clips) // int(world))
    start = min(rank * per_rank, n_clips)
    count = max(0, min(n_clips, start + per_rank) - start)
    return start, count, per_rank


class FeatureCacheGather:
    """Per-rank cache [world * per_rank, frames, cols] float32, peer-mapped on every rank.

    `group` is a torch.distributed process group (or None for a single process); `local_peers` lets one
    process emulate several ranks on one GPU (tests): a list of device pointers standing for the ranks' caches.
    """

    def __init__(self, plan, n_clips, clip_len, world, rank, device, group=None, multicast=False):
        self.plan, self.n_clips, self.clip_len = plan, int(n_clips), int(clip_len)
        self.world, self.rank, self.device = int(world), int(rank), int(device)
        self.frames = plan.frames(clip_len)
        self.cols = plan.out_cols
        self.start, self.count, self.per_rank = shard_range(n_clips, world, rank)
        self.rows = self.world * self.per_rank
        self.bytes = self.rows * self.frames * self.cols * 4
        self._own = ctypes.c_void_p()
        self._imported = []
        self._symm = None                  # (tensor, handle) when the cache is torch symmetric memory
        self.multicast_ptr = 0
        if multicast:
            self._init_multicast(group)
            return
        check(_lib.lib().scf_device_malloc(self.device, self.bytes, ctypes.byref(self._own)))
        self._peers = [None] * self.world
        self._peers[self.rank] = self._own.value
        if self.world > 1:
            if group is None:
                raise ValueError('world > 1 needs a torch.distributed group to exchange IPC handles')
            import torch.distributed as dist
            handle = (ctypes.c_uint8 * 64)()
            check(_lib.lib().scf_ipc_export(self.device, self._own, handle))
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(handle), group=group)
            for r, h in enumerate(handles):
                if r == self.rank:
                    continue
                buf = (ctypes.c_uint8 * 64).from_buffer_copy(h)
                p = ctypes.c_void_p()
                check(_lib.lib().scf_ipc_import(self.device, buf, ctypes.byref(p)))
                self._peers[r] = p.value
                self._imported.append(p.value)
        self._table = (ctypes.c_void_p * self.world)(*self._peers)

    def _init_multicast(self, group):
        """Cache in torch symmetric memory; needs a process group and multicast support on the fabric."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem
        if group is None:
            group = dist.group.WORLD
        t = symm_mem.empty(self.bytes // 4, dtype=torch.float32, device=torch.device('cuda', self.device))
        hdl = symm_mem.rendezvous(t, group=group)
        mc = int(getattr(hdl, 'multicast_ptr', 0) or 0)
        if mc == 0:
            raise _lib.ScfError(-4, 'no multicast address for the symmetric cache (fabric without NVLink SHARP multicast?)')
        self._symm = (t, hdl)
        self._own = ctypes.c_void_p(t.data_ptr())
        self.multicast_ptr = mc
        self._peers = [int(p) for p in hdl.buffer_ptrs]
        self._table = (ctypes.c_void_p * self.world)(*self._peers)

    @property
    def ptr(self):
        return self._own.value

    def extract_and_gather(self, d_pcm_local, stream=0, clip_stride=None):
        """Extracts this rank's `count` clips (device int16 [count, clip_stride]) and stores every row into
        all ranks' caches.  Stream-ordered; call barrier() (or torch.distributed.barrier) before reading."""
        if self.count == 0:
            return
        if self.multicast_ptr:
            check(_lib.lib().scf_extract_i16_gather_multicast(self.plan.handle, d_pcm_local, self.count,
                                                              self.clip_len if clip_stride is None else clip_stride,
                                                              self.clip_len, self.multicast_ptr, self.world, self.rank,
                                                              self.per_rank, stream))
            return
        check(_lib.lib().scf_extract_i16_gather(self.plan.handle, d_pcm_local, self.count,
                                                self.clip_len if clip_stride is None else clip_stride,
                                                self.clip_len, self._table, self.world, self.rank, self.per_rank, stream))

    def to_host(self):
        """The first n_clips rows of this rank's cache as float32 numpy [n_clips, frames, cols]
        (rank r's rows start at r * per_rank; the last rank's padding rows come last, so clip order holds)."""
        out = np.empty((self.rows, self.frames, self.cols), dtype=np.float32)
        check(_lib.lib().scf_memcpy(self.device, out.ctypes.data, self._own, self.bytes, 1, None))
        return out[:self.n_clips]

    def to_dlpack(self):
        """The gathered cache as a DLPack capsule [n_clips, frames, cols, 1] float32 on this rank's GPU -- what
        classifier/data.py:97-120 assembles from 105k .npy files, handed to the framework without leaving the device
        (tf.experimental.dlpack.from_dlpack / torch.from_dlpack).  Rank r's rows start at r * per_rank and the padding
        rows of the last rank come last, so the first n_clips rows are the clips in order.  The cache stays alive
        until the consumer releases the tensor; call after the closing barrier."""
        from .plan import dlpack_wrap
        return dlpack_wrap(self, self._own.value, (self.n_clips, self.frames, self.cols, 1), self.device)

    def close(self):
        for p in self._imported:
            _lib.lib().scf_ipc_close(self.device, p)
        self._imported = []
        if self._symm is not None:           # torch owns the symmetric allocation
            self._symm = None
            self._own = ctypes.c_void_p()
            self.multicast_ptr = 0
        elif self._own:
            _lib.lib().scf_device_free(self.device, self._own)
            self._own = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
