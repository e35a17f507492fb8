"""GPU (one device): the peer-memory key-frame exchange + fused Hamming matcher with world = 1 (the rank's own record): identical to the batched k = 2 matcher
and to cv2's BFMatcher golden vectors.  The multi-GPU path (records of other ranks mapped through CUDA IPC over NVLink) is checked by tools/exchange_check.py
under torchrun on 2+ GPUs: same kernel, same comparison against pslam_hamming_knn2 over the concatenated set."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_exchange_single_rank_matches_knn2_and_cv2_golden():
    import torch
    from planarslam_b200._lib import Context
    from planarslam_b200.matcher import ORBmatcher
    from planarslam_b200.sharding import PeerDescriptorExchange
    g = np.load(os.path.join(GOLD, "bfmatcher_cv2_4_13.npz"))
    ctx = Context(640, 480, 1)
    cap = 512
    ex = PeerDescriptorExchange(ctx, cap, slots=2)
    dev = torch.device("cuda", 0)
    rng = np.random.default_rng(4)
    for epoch, (q, t) in enumerate([(g["q"], g["t"]), (rng.integers(0, 256, (300, 32), dtype=np.uint8), rng.integers(0, 256, (512, 32), dtype=np.uint8)),
                                     (rng.integers(0, 256, (40, 32), dtype=np.uint8), np.zeros((0, 32), np.uint8))], start=1):
        slot = epoch % 2
        tt = np.zeros((cap, 32), np.uint8)
        tt[:len(t)] = t
        d_t, d_nt = torch.from_numpy(tt).to(dev), torch.tensor([len(t)], dtype=torch.int32, device=dev)
        d_q, d_nq = torch.from_numpy(np.ascontiguousarray(q)).to(dev), torch.tensor([len(q)], dtype=torch.int32, device=dev)
        d_idx, d_dist = torch.empty((len(q), 2), dtype=torch.int32, device=dev), torch.empty((len(q), 2), dtype=torch.int32, device=dev)
        ex.publish(slot, d_t, d_nt, epoch)
        ex.match(slot, epoch, d_q, d_nq, d_idx, d_dist)
        torch.cuda.synchronize()
        idx, dist, _ = ORBmatcher(ctx=ctx).knn2(q, t) if len(t) else (np.full((len(q), 2), -1, np.int32), np.full((len(q), 2), 256, np.int32), None)
        assert np.array_equal(d_idx.cpu().numpy(), idx) and np.array_equal(d_dist.cpu().numpy(), dist), epoch
        if epoch == 1:
            assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
        ex.barrier()
