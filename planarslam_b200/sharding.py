"""Multi-GPU plumbing for the per-frame hot path: one process per GPU, frames sharded across ranks with no
data-path collective, and ONE exchange step — an all-gather of fixed-capacity keyframe descriptor blocks — feeding the
brute-force Hamming matcher used by relocalisation / loop-closure style queries.

torch.distributed is plumbing only (NCCL over NVLink on the GPU box, gloo in the CPU tests); the matching itself is the
CUDA kernel behind pslam_hamming_knn2_batch_dev.
"""
from __future__ import annotations

import numpy as np


def shard_frames(n_frames: int, rank: int, world: int) -> np.ndarray:
    """Frame i goes to rank i mod world (round-robin keeps temporal neighbours on different GPUs)."""
    if not (0 <= rank < world):
        raise ValueError("rank outside [0, world)")
    return np.arange(rank, n_frames, world, dtype=np.int64)


def frames_per_rank(n_frames: int, world: int) -> list[int]:
    return [len(range(r, n_frames, world)) for r in range(world)]


class KeyframeDescriptorExchange:
    """All-gather of per-rank keyframe descriptor blocks.

    Each rank owns up to `cap` 32-byte descriptors (its keyframes' rBRIEF rows) and their count.  `gather()` returns the
    concatenation over ranks as one [world*cap, 32] buffer plus the counts, on the same device as the input (a single
    `all_gather_into_tensor` of a fixed-size block: latency-bound at these sizes).
    `to_global` / `from_global` translate (rank, local row) <-> row of the gathered train set after compaction."""

    def __init__(self, cap: int, group=None):
        import torch.distributed as dist
        self.dist, self.group, self.cap = dist, group, int(cap)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)

    def gather(self, desc, count, sync_counts: bool = True):
        """count: a Python int, or a one-element integer tensor on desc's device (then nothing crosses to the host on the way in).  With sync_counts=False the
        per-rank counts come back as an int64 tensor on the device as well: the call is then free of host round trips (the block and the gathered buffer are
        allocated once and reused, so the returned views are valid until the next call; this is the NCCL comparison path - the product exchange is
        PeerDescriptorExchange below)."""
        import torch
        if desc.shape != (self.cap, 32) or desc.dtype != torch.uint8:
            raise ValueError("desc must be a [cap, 32] uint8 tensor")
        n = self.cap * 32
        if getattr(self, "_block", None) is None or self._block.device != desc.device:
            self._block = torch.zeros(n + 8, dtype=torch.uint8, device=desc.device)
            self._out = torch.empty(self.world * (n + 8), dtype=torch.uint8, device=desc.device)
        block, out = self._block, self._out
        block[:n] = desc.reshape(-1)
        if torch.is_tensor(count):
            block[n:] = count.reshape(1).to(torch.int64).view(torch.uint8)
        else:
            block[n:] = torch.tensor([int(count)], dtype=torch.int64).view(torch.uint8).to(desc.device)
        self.dist.all_gather_into_tensor(out, block, group=self.group)
        rows = out.view(self.world, -1)
        counts = rows[:, n:].contiguous().view(torch.int64).reshape(self.world)
        gathered = rows[:, :n].reshape(self.world, self.cap, 32)
        return gathered, (counts.cpu().numpy() if sync_counts else counts)

    @staticmethod
    def compact(gathered, counts):
        """[world, cap, 32] + counts -> ([sum(counts), 32] train set, offsets[world+1])."""
        import torch
        offs = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
        train = torch.cat([gathered[r, : int(counts[r])] for r in range(len(counts))], 0) if offs[-1] else gathered.new_zeros((0, 32))
        return train, offs

    @staticmethod
    def from_global(row: int, offs: np.ndarray):
        r = int(np.searchsorted(offs, row, side="right") - 1)
        return r, int(row - offs[r])

    @staticmethod
    def to_global(rank: int, local: int, offs: np.ndarray) -> int:
        return int(offs[rank] + local)


class PeerDescriptorExchange:
    """The exchange step on hardware (include/pslam_abi.h pslam_exchange_*): every rank publishes key-frame ORB blocks into records in its own HBM, the peers map
    them through CUDA IPC (NVLink / NVSwitch P2P), and `match` runs the fused wait-on-flag + Hamming k = 2 kernel that reads the peers' records in place.
    torch.distributed only carries the 64-byte IPC handles (once) and the epoch barriers that guard slot reuse."""

    def __init__(self, ctx, cap: int, slots: int = 1, group=None):
        import ctypes as C
        import torch
        self.C, self.torch, self.ctx, self.cap, self.slots, self.group = C, torch, ctx, int(cap), int(slots), group
        L = ctx.L
        L.pslam_exchange_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pslam_exchange_attach.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.pslam_exchange_publish_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.pslam_exchange_match_dev.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        handle = np.zeros(64, np.uint8)
        ctx.check(L.pslam_exchange_create(ctx.h, self.cap, self.slots, handle.ctypes.data))
        import torch.distributed as dist
        self.dist = dist if dist.is_available() and dist.is_initialized() else None
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        handles = handle[None]
        if self.world > 1:
            dev = torch.device("cuda", ctx.cfg.device)
            mine = torch.from_numpy(handle).to(dev)
            allh = torch.empty(self.world * 64, dtype=torch.uint8, device=dev)
            self.dist.all_gather_into_tensor(allh, mine, group=group)
            handles = allh.cpu().numpy().reshape(self.world, 64)
        handles = np.ascontiguousarray(handles)
        ctx.check(L.pslam_exchange_attach(ctx.h, self.world, self.rank, handles.ctypes.data))

    def publish(self, slot: int, d_desc, d_n, epoch: int, d_kps=None):
        """d_desc [cap, 32] uint8 and d_n int32 [1] device tensors (the ORB output of a key frame); enqueued on the context's stream."""
        self.ctx.check(self.ctx.L.pslam_exchange_publish_dev(self.ctx.h, slot, d_desc.data_ptr(), d_kps.data_ptr() if d_kps is not None else None, d_n.data_ptr(), epoch))

    def match(self, slot: int, epoch: int, d_q, d_nq, d_idx, d_dist):
        """k = 2 nearest rows of every query over the concatenation of all ranks' records of `slot` (waits inside the kernel for each peer's epoch flag)."""
        self.ctx.check(self.ctx.L.pslam_exchange_match_dev(self.ctx.h, slot, epoch, d_q.data_ptr(), d_nq.data_ptr(), int(d_q.shape[0]), d_idx.data_ptr(), d_dist.data_ptr()))

    def barrier(self):
        """Call before re-publishing a slot: every rank must have finished matching the old epoch."""
        self.torch.cuda.synchronize()
        if self.dist and self.world > 1:
            self.dist.barrier(group=self.group)
