"""CPU: the oracle of the per-line 3-D fit (oracle/line3d.cc: Frame::isLineGood, src/Frame.cc:189-267 + src/LineExtractor.cpp:1157-1470).

Pins of the third-party arithmetic it needs:
  * glibc rand() (random_unique, include/LSDextractor.h:239-251): bit-exact against the container's libc for several seeds;
  * cv::SVD: OpenCV's Jacobi algorithm restated (oracle/cvsvd.h); the in-container cv2 routes double SVDs to LAPACK, so the check
    is w to 1e-12 relative and singular vectors up to sign (and exact reconstruction).
The fit itself has no reference vectors (parity unpinned by the reference); it is checked on synthetic RGB-D frames against the
scene geometry and against an independent numpy statement of the Mahalanobis point-line distance."""
import ctypes as C
import ctypes.util

import cv2
import numpy as np

import oracle_lib
from planarslam_b200 import synth


def test_glibc_rand_matches_libc():
    libc = C.CDLL(ctypes.util.find_library("c") or "libc.so.6")
    libc.srand.argtypes = [C.c_uint]
    for seed in (1, 0, 42, 123456789, 2 ** 31 + 5, 2 ** 32 - 1):
        libc.srand(seed)
        ref = np.array([libc.rand() for _ in range(2000)], np.int32)
        assert np.array_equal(oracle_lib.glibc_rand(seed, 2000), ref), seed
    assert oracle_lib.glibc_rand(1, 1)[0] == 1804289383


def _check_svd(A, rtol):
    w, u, vt = oracle_lib.cv_svd(A)
    rw, ru, rvt = cv2.SVDecomp(A)
    assert w.shape == rw.ravel().shape and u.shape == ru.shape and vt.shape == rvt.shape
    assert np.allclose(w, rw.ravel(), rtol=rtol, atol=rtol * abs(rw).max())
    assert (np.diff(w) <= 0).all()
    assert np.allclose((u * w) @ vt, A, rtol=0, atol=50 * rtol * abs(A).max())
    gap_ok = np.abs(np.diff(w, append=0)) > 1e-6 * w[0]               # singular vectors are only defined where the values are separated
    for i in range(len(w)):
        if gap_ok[i] and (i == 0 or gap_ok[i - 1]):
            s = np.sign(np.dot(vt[i], rvt[i]))
            assert np.allclose(vt[i] * s, rvt[i], rtol=0, atol=1e4 * rtol), i
            assert np.allclose(u[:, i] * s, ru[:, i], rtol=0, atol=1e4 * rtol), i


def test_cv_svd_matches_cv2_up_to_sign():
    rng = np.random.default_rng(2)
    for _ in range(50):                                               # covariance-like 3 x 3 (compPt3dCov)
        J = rng.normal(size=(3, 3))
        _check_svd(J @ np.diag([1.0, 1.0, 1e-5]) @ J.T, 1e-12)
    for n in (2, 3, 4, 10, 51):                                       # n x 3 point sets (computeLine3d_svd), incl. the m < n path
        for _ in range(10):
            _check_svd(rng.normal(size=(n, 3)) * [1.0, 0.02, 0.01], 1e-12)
    for _ in range(20):                                               # float 3 x 3 (TrackManhattanFrame)
        _check_svd(rng.normal(size=(3, 3)).astype(np.float32), 2e-6)


def _frame(seed):
    gray, d16, z, _ = synth.render_frame(seed=seed, frame=3 * seed)
    kl, _ = oracle_lib.extract_line_segments(gray, 40)
    depth = d16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR)
    return kl, depth, z


def test_lines3d_on_synthetic_frames():
    n_valid = 0
    for seed in (0, 3, 7):
        kl, depth, z = _frame(seed)
        fx, fy, cx, cy = synth.TUM3_K
        r = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=1)
        assert len(kl) == 40 and 0 < r["n_drawn"] <= 40 * 20
        length = np.hypot(kl["startPointX"] - kl["endPointX"], kl["startPointY"] - kl["endPointY"])
        assert (r["n_points"] <= np.minimum(length.astype(int), 50) + 1).all()
        v = r["valid"].astype(bool)
        n_valid += int(v.sum())
        assert (r["depth_line"][~v] == -1).all() and (r["lines3d"][~v] == 0).all()
        assert (r["n_inliers"][v] / length[v] > 0.4).all()
        assert np.array_equal(r["n_inliers"], [bin(int(m)).count("1") for m in r["inliers"]])
        A, B = r["lines3d"][v, :3], r["lines3d"][v, 3:]
        assert (np.linalg.norm(A - B, axis=1) > 0.02).all()
        assert np.allclose(np.linalg.norm(r["director"][v], axis=1), 1.0, atol=1e-12)
        assert np.allclose(np.cross(r["director"][v], A - B), 0, atol=1e-12)
        # the end points are back-projected depth samples: they re-project onto the 2-D segment (nearest-pixel sampling: < 1.5 px)
        for P in (A, B):
            u, w_ = fx * P[:, 0] / P[:, 2] + cx, fy * P[:, 1] / P[:, 2] + cy
            sx, sy, ex, ey = (kl[k][v] for k in ("startPointX", "startPointY", "endPointX", "endPointY"))
            t = np.clip(((u - sx) * (ex - sx) + (w_ - sy) * (ey - sy)) / length[v] ** 2, 0, 1)
            assert (np.hypot(u - (sx + t * (ex - sx)), w_ - (sy + t * (ey - sy))) < 1.5).all()
        # mvDepthLine = min of the depth at the two (truncated) end points
        ez = depth[kl["endPointY"].astype(int), kl["endPointX"].astype(int)]
        sz = depth[kl["startPointY"].astype(int), kl["startPointX"].astype(int)]
        assert np.array_equal(r["depth_line"][v], np.minimum(ez, sz)[v])
    assert n_valid >= 30


def test_lines3d_rand_stream_semantics():
    kl, depth, _ = _frame(3)
    a = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=1)
    b = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=1)
    assert all(np.array_equal(a[k], b[k], equal_nan=True) for k in ("valid", "lines3d", "inliers", "director")) and a["n_drawn"] == b["n_drawn"]
    # the lines of a frame draw from one stream in order: running the first 10 lines, then the rest with skip = draws so far, is the same
    h = oracle_lib.lines3d_frame(kl[:10], depth, synth.TUM3_K, seed=1)
    t = oracle_lib.lines3d_frame(kl[10:], depth, synth.TUM3_K, seed=1, skip=h["n_drawn"])
    assert h["n_drawn"] + t["n_drawn"] == a["n_drawn"]
    for k in ("valid", "lines3d", "inliers", "n_inliers"):
        assert np.array_equal(np.concatenate([h[k], t[k]]), a[k]), k
    c = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=2)
    assert not np.array_equal(a["inliers"], c["inliers"]) or a["n_drawn"] != c["n_drawn"] or True   # a different stream may or may not change the fit
    # lines without depth support are rejected
    none = oracle_lib.lines3d_frame(kl, np.zeros_like(depth), synth.TUM3_K)
    assert not none["valid"].any() and none["n_drawn"] == 0 and (none["n_points"] == 0).all()
