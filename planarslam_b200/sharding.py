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

    def gather(self, desc, count: int):
        import torch
        if desc.shape != (self.cap, 32) or desc.dtype != torch.uint8:
            raise ValueError("desc must be a [cap, 32] uint8 tensor")
        block = torch.zeros(self.cap * 32 + 8, dtype=torch.uint8, device=desc.device)
        block[: self.cap * 32] = desc.reshape(-1)
        block[self.cap * 32:] = torch.tensor([count], dtype=torch.int64).view(torch.uint8).to(desc.device)
        out = torch.empty(self.world * block.numel(), dtype=torch.uint8, device=desc.device)
        self.dist.all_gather_into_tensor(out, block, group=self.group)
        out = out.view(self.world, -1)
        counts = out[:, self.cap * 32:].contiguous().view(torch.int64).reshape(self.world).cpu().numpy()
        return out[:, : self.cap * 32].reshape(self.world, self.cap, 32), counts

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
