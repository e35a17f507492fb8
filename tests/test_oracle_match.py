"""CPU: pin the oracle's brute-force Hamming matching against cv2.BFMatcher (the reference calls cv::BFMatcher,
src/ORBmatcher.cc:1346-1347, src/LSDmatcher.cpp:249-254)."""
import ctypes as C

import numpy as np
import pytest

import oracle_lib

cv2 = pytest.importorskip("cv2")


def _lib():
    L = oracle_lib.lib()
    L.orc_bf_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_bf_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.orc_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    return L


def test_bf_match_and_knn_match_cv2():
    L = _lib()
    rng = np.random.default_rng(5)
    for (nq, nt, nbits) in [(1000, 1000, 256), (40, 40, 256), (300, 7, 256), (500, 500, 6), (64, 1, 256)]:
        # few distinct bits -> many exact distance ties, exercises the tie-breaking order
        q = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
        t = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
        if nbits < 256:
            q[:, 1:] = 0
            t[:, 1:] = 0
            q[:, 0] &= (1 << nbits) - 1
            t[:, 0] &= (1 << nbits) - 1
        idx, dist = np.zeros(nq, np.int32), np.zeros(nq, np.int32)
        L.orc_bf_match(q.ctypes.data, nq, t.ctypes.data, nt, idx.ctypes.data, dist.ctypes.data)
        m = cv2.BFMatcher(cv2.NORM_HAMMING).match(q, t)
        assert [x.trainIdx for x in m] == idx.tolist() and [int(x.distance) for x in m] == dist.tolist()
        idx2, dist2 = np.zeros((nq, 2), np.int32), np.zeros((nq, 2), np.int32)
        L.orc_bf_knn2(q.ctypes.data, nq, t.ctypes.data, nt, idx2.ctypes.data, dist2.ctypes.data)
        k = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(q, t, 2)
        for i, pair in enumerate(k):
            for r, x in enumerate(pair):
                assert x.trainIdx == idx2[i, r] and int(x.distance) == dist2[i, r], (nq, nt, i, r)
            for r in range(len(pair), 2):
                assert idx2[i, r] == -1


def test_descriptor_distance_is_popcount():
    L = _lib()
    rng = np.random.default_rng(6)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8)
        b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert L.orc_descriptor_distance(a.ctypes.data, b.ctypes.data) == int(np.unpackbits(a ^ b).sum())
