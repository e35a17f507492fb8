#!/usr/bin/env python3
"""Multi-GPU check of the peer-memory key-frame exchange (run under torchrun, one rank per GPU):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 tools/exchange_check.py
Every rank extracts ORB from its own frames on its own GPU, publishes the blocks, and matches its descriptors against the records of ALL ranks read in place
over NVLink (pslam_exchange_match_dev).  The result must equal pslam_hamming_knn2 over the concatenated set, which each rank rebuilds here from an NCCL
all_gather of the raw blocks (the comparison path only).  Prints one JSON line per rank 0 with timings of the fused kernel and of all_gather + knn2."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    from planarslam_b200 import synth
    from planarslam_b200._lib import Context
    from planarslam_b200.sharding import PeerDescriptorExchange
    W, H, n_kf = (1280, 960, 4) if os.environ.get("PSLAM_CONFIG") == "5" else (640, 480, 8)
    nfeat = 2000 if W == 1280 else 1000
    ctx = Context(W, H, n_kf, device=local, nfeatures=nfeat)
    L = ctx.L
    ctx.set_stream(torch.cuda.current_stream(dev).cuda_stream)          # one stream: the CUDA events below bracket the library's launches
    cap = int(L.pslam_orb_max_keypoints(ctx.h))
    frames = np.stack([synth.render_frame(2, 8 * (rank * n_kf + k) % 64, W, H)[0] for k in range(n_kf)])
    d_gray = torch.from_numpy(frames).to(dev)
    d_kps = torch.empty((n_kf, cap, 28), dtype=torch.uint8, device=dev)
    d_desc = torch.empty((n_kf, cap, 32), dtype=torch.uint8, device=dev)
    d_n = torch.zeros(n_kf, dtype=torch.int32, device=dev)
    ctx.check(L.pslam_orb_extract_batch_dev(ctx.h, d_gray.data_ptr(), n_kf, d_kps.data_ptr(), d_desc.data_ptr(), cap, d_n.data_ptr()))
    ex = PeerDescriptorExchange(ctx, cap, slots=n_kf)
    d_idx = torch.empty((n_kf, cap, 2), dtype=torch.int32, device=dev)
    d_dist = torch.empty((n_kf, cap, 2), dtype=torch.int32, device=dev)
    ok = True
    t_fused, t_base = [], []
    for epoch in (1, 2, 3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ex.barrier()
        e0.record()
        for k in range(n_kf):
            ex.publish(k, d_desc[k], d_n[k:k + 1], epoch, d_kps[k])
        for k in range(n_kf):
            ex.match(k, epoch, d_desc[k], d_n[k:k + 1], d_idx[k], d_dist[k])
        e1.record()
        torch.cuda.synchronize()
        t_fused.append(e0.elapsed_time(e1))
        # comparison path: NCCL all_gather of the raw blocks + the single-GPU matcher over the concatenation
        b0, b1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        b0.record()
        if world > 1:
            all_desc = torch.empty((world,) + tuple(d_desc.shape), dtype=torch.uint8, device=dev)
            all_n = torch.empty((world, n_kf), dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(all_desc, d_desc)
            dist.all_gather_into_tensor(all_n, d_n)
        else:
            all_desc, all_n = d_desc[None], d_n[None]
        b1.record()
        torch.cuda.synchronize()
        t_base.append(b0.elapsed_time(b1))
        nn = all_n.cpu().numpy()
        for k in range(n_kf):
            train = torch.cat([all_desc[r, k, :int(nn[r, k])] for r in range(world)], 0).contiguous()
            nt = torch.tensor([train.shape[0]], dtype=torch.int32, device=dev)
            r_idx = torch.empty((cap, 2), dtype=torch.int32, device=dev)
            r_dist = torch.empty((cap, 2), dtype=torch.int32, device=dev)
            ctx.check(L.pslam_hamming_knn2_batch_dev(ctx.h, d_desc[k].data_ptr(), d_n[k:k + 1].data_ptr(), cap, train.data_ptr(), nt.data_ptr(), int(train.shape[0]), 1,
                                                     r_idx.data_ptr(), r_dist.data_ptr(), None, None))
            torch.cuda.synchronize()
            nq = int(nn[rank, k])
            same = torch.equal(d_idx[k, :nq], r_idx[:nq]) and torch.equal(d_dist[k, :nq], r_dist[:nq])
            ok = ok and same and nq > 0.5 * nfeat
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"exchange_check": "ok" if int(flag.item()) else "MISMATCH", "world": world, "key_frames_per_rank": n_kf, "rows_per_key_frame": int(d_n[0].item()),
                          "fused_publish_match_ms": [round(v, 3) for v in t_fused], "nccl_all_gather_ms": [round(v, 3) for v in t_base], "size": [W, H]}))
    if world > 1:
        dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
