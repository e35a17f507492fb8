"""CPU: the N>1 host logic (frame sharding, keyframe-descriptor all-gather) over gloo with world_size 2."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from planarslam_b200.sharding import KeyframeDescriptorExchange, shard_frames
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        cap = 64
        rng = np.random.default_rng(100 + rank)
        cnt = 40 + 7 * rank
        desc = torch.zeros((cap, 32), dtype=torch.uint8)
        desc[:cnt] = torch.from_numpy(rng.integers(0, 256, (cnt, 32), dtype=np.uint8))
        ex = KeyframeDescriptorExchange(cap)
        gathered, counts = ex.gather(desc, cnt)
        train, offs = ex.compact(gathered, counts)
        mine = shard_frames(11, rank, world)
        q.put((rank, counts.tolist(), offs.tolist(), train.numpy().tobytes(), desc[:cnt].numpy().tobytes(), mine.tolist(),
               ex.from_global(int(offs[1]) + 3, offs), ex.to_global(1, 3, offs)))
    finally:
        dist.destroy_process_group()


def test_descriptor_exchange_and_sharding_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (r0, c0, o0, t0, d0, m0, fg0, tg0), (r1, c1, o1, t1, d1, m1, fg1, tg1) = res
    assert c0 == c1 == [40, 47] and o0 == o1 == [0, 40, 87]
    assert t0 == t1 == d0 + d1                                # every rank sees rank 0's rows followed by rank 1's
    assert m0 == [0, 2, 4, 6, 8, 10] and m1 == [1, 3, 5, 7, 9]    # frames sharded round-robin, disjoint and complete
    assert fg0 == (1, 3) and tg0 == 43


def test_shard_frames_properties():
    sys.path.insert(0, ROOT)
    from planarslam_b200.sharding import frames_per_rank, shard_frames
    for n in (0, 1, 7, 64, 513):
        for w in (1, 2, 4, 8):
            parts = [shard_frames(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(n))
            assert [len(p) for p in parts] == frames_per_rank(n, w)
    with pytest.raises(ValueError):
        shard_frames(4, 2, 2)
