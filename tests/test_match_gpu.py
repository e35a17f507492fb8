"""GPU: brute-force Hamming matching through the C ABI vs the CPU oracle (bit-exact indices and distances)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib

pytestmark = pytest.mark.gpu


def _oracle(q, t):
    L = oracle_lib.lib()
    L.orc_bf_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_match_gate.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    nq = len(q)
    idx, dist = np.zeros((max(nq, 1), 2), np.int32), np.zeros((max(nq, 1), 2), np.int32)
    L.orc_bf_knn2(q.ctypes.data, nq, t.ctypes.data, len(t), idx.ctypes.data, dist.ctypes.data)
    d1 = np.ascontiguousarray(dist[:nq, 0])
    keep = np.zeros(max(nq, 1), np.int32)
    # an empty train set yields no DMatch at all (cv::BFMatcher::match, src/ORBmatcher.cc:1346-1366): MatchORBPoints keeps nothing
    n = L.orc_match_gate(d1.ctypes.data, nq, keep.ctypes.data) if len(t) else 0
    return idx[:nq], dist[:nq], keep[:n]


def test_knn2_and_gate_match_oracle():
    from planarslam_b200.matcher import ORBmatcher
    m = ORBmatcher()
    rng = np.random.default_rng(9)
    for (nq, nt, nbits) in [(1000, 1000, 256), (2000, 2000, 256), (40, 40, 256), (1005, 3, 256), (500, 700, 5), (7, 1, 256), (300, 0, 256)]:
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nbits < 256:                      # many exact ties
            q[:, 1:] = 0; t[:, 1:] = 0
            q[:, 0] &= (1 << nbits) - 1; t[:, 0] &= (1 << nbits) - 1
        if nt > 10:                          # plant true correspondences so the gate keeps a realistic subset
            sel = rng.choice(nq, nq // 2, replace=False)
            t_sel = rng.integers(0, nt, len(sel))
            q[sel] = t[t_sel] ^ (rng.random((len(sel), 32)) < 0.02).astype(np.uint8)
        idx, dist, good = m.knn2(q, t, gate=True)
        oi, od, og = _oracle(q, t)
        assert np.array_equal(idx, oi) and np.array_equal(dist, od), (nq, nt)
        assert np.array_equal(good, og), (nq, nt)


def test_orb_descriptors_end_to_end():
    """MatchORBPoints on real ORB descriptors of two nearby synthetic frames."""
    from planarslam_b200 import synth
    from planarslam_b200.orb import ORBextractor
    from planarslam_b200.matcher import ORBmatcher
    ext = ORBextractor()
    k1, d1 = ext(synth.render_frame(2, 10)[0])
    k2, d2 = ext(synth.render_frame(2, 11)[0])
    n, good = ORBmatcher().MatchORBPoints(d2, d1)
    oi, od, og = _oracle(d2, d1)
    assert n == len(og) and [g[0] for g in good] == og.tolist() and [g[1] for g in good] == oi[og, 0].tolist()
    assert n > 50
