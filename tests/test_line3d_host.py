"""CPU: planarslam_b200/csrc/line3d_body.h - the code the CUDA kernel k_lines3d runs, one thread per frame - compiled for the HOST
with g++ and compared with the oracle (oracle/line3d.cc, an independent statement of Frame::isLineGood with std::vector sets, the
generic Jacobi SVD and libm hypot).  The body is plain IEEE double arithmetic, nvcc builds it with --fmad=false, so the host
result is what the device computes; the kernel around it only indexes frames.  No GPU time was left in round 1 to run the kernel
itself (GPU run: tests/test_line3d_gpu.py)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("line3d") / "libline3d_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                    "-I", os.path.join(ROOT, "planarslam_b200", "csrc"), "-o", str(out), os.path.join(ROOT, "tests", "host_harness", "line3d_host.cc")], check=True)
    L = C.CDLL(str(out))
    L.host_lines3d_frame.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p, C.c_uint32, C.c_int] + [C.c_void_p] * 7
    L.host_glibc_rand.argtypes = [C.c_uint32, C.c_int, C.c_int, C.c_void_p]
    return L


def host_lines3d(L, kl, d16, cam, seed=1, skip=0, factor=np.float32(1.0 / synth.DEPTH_FACTOR)):
    kl = np.ascontiguousarray(kl, oracle_lib.KEYLINE_DTYPE)
    d16 = np.ascontiguousarray(d16, np.uint16)
    n = len(kl)
    camv = np.asarray(cam, np.float32)
    o = dict(lines3d=np.zeros((n, 6)), director=np.zeros((n, 3)), inliers=np.zeros(n, np.uint64), depth_line=np.zeros(n, np.float32),
             n_points=np.zeros(n, np.int32), n_inliers=np.zeros(n, np.int32), valid=np.zeros(n, np.int32))
    o["n_drawn"] = L.host_lines3d_frame(kl.ctypes.data, n, d16.ctypes.data, d16.shape[1], d16.shape[0], float(factor), camv.ctypes.data, seed, skip,
                                        o["lines3d"].ctypes.data, o["director"].ctypes.data, o["inliers"].ctypes.data, o["depth_line"].ctypes.data,
                                        o["n_points"].ctypes.data, o["n_inliers"].ctypes.data, o["valid"].ctypes.data)
    return o


def test_body_rand_matches_oracle(host_lib):
    for seed, skip in ((1, 0), (42, 17), (2 ** 31 + 5, 0), (0, 3)):
        out = np.zeros(500, np.int32)
        host_lib.host_glibc_rand(seed, skip, 500, out.ctypes.data)
        assert np.array_equal(out, oracle_lib.glibc_rand(seed, 500 + skip)[skip:]), (seed, skip)
    # the jump-ahead path of l3d_srand (more than L3D_JUMP_ABOVE = 4096 discarded draws: x^D modulo x^31 - x^28 - 1) against the stepping oracle
    for seed, skip in ((1, 3787), (1, 3786), (1, 3785), (1, 4097), (7, 5000), (1, 65536), (123456789, 1000003), (2 ** 31 + 5, 3141592)):
        out = np.zeros(200, np.int32)
        host_lib.host_glibc_rand(seed, skip, 200, out.ctypes.data)
        assert np.array_equal(out, oracle_lib.glibc_rand(seed, 200 + skip)[skip:]), (seed, skip)


def test_body_matches_oracle_on_synthetic_frames(host_lib):
    n_valid = 0
    for s in range(8):
        gray, d16, _, _ = synth.render_frame(seed=s, frame=3 * s)
        kl, _ = oracle_lib.extract_line_segments(gray, 40)
        depth = d16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR)
        for seed in (1, 77 + s):
            o = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=seed)
            p = host_lines3d(host_lib, kl, d16, synth.TUM3_K, seed=seed)
            assert p["n_drawn"] == o["n_drawn"], (s, seed)
            assert np.array_equal(p["valid"], o["valid"]) and np.array_equal(p["n_points"], o["n_points"])
            assert np.array_equal(p["inliers"], o["inliers"]) and np.array_equal(p["n_inliers"], o["n_inliers"])
            assert np.array_equal(p["lines3d"], o["lines3d"]) and np.array_equal(p["depth_line"], o["depth_line"])
            assert np.array_equal(p["director"], o["director"], equal_nan=True)
            n_valid += int(p["valid"].sum())
    assert n_valid > 300


def test_body_edge_cases(host_lib):
    gray, d16, _, _ = synth.render_frame(seed=2, frame=6)
    kl, _ = oracle_lib.extract_line_segments(gray, 40)
    none = host_lines3d(host_lib, kl, np.zeros_like(d16), synth.TUM3_K)
    assert not none["valid"].any() and none["n_drawn"] == 0 and (none["depth_line"] == -1).all()
    short = kl[:3].copy()                                                # segments shorter than 10 px cannot collect 10 samples
    short["endPointX"] = short["startPointX"] + 5
    short["endPointY"] = short["startPointY"]
    depth = d16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR)
    r = host_lines3d(host_lib, short, d16, synth.TUM3_K)
    o = oracle_lib.lines3d_frame(short, depth, synth.TUM3_K)
    assert not r["valid"].any() and (r["n_points"] <= 6).all() and np.array_equal(r["n_points"], o["n_points"])
    # integer sample positions take the "boundary" branch (col - 1, row - 1): a horizontal segment on integer coordinates
    hz = kl[:1].copy()
    hz["startPointX"], hz["startPointY"], hz["endPointX"], hz["endPointY"] = 100.0, 200.0, 150.0, 200.0
    r = host_lines3d(host_lib, hz, d16, synth.TUM3_K)
    o = oracle_lib.lines3d_frame(hz, depth, synth.TUM3_K)
    for k in ("valid", "n_points", "inliers", "lines3d", "depth_line"):
        assert np.array_equal(r[k], o[k]), k


def test_body_matches_oracle_with_icl_intrinsics(host_lib):
    """Examples/RGB-D/ICL.yaml:6-10: fy = -480 (y flips sign in every back-projection), depth factor 5000."""
    icl = (481.2, -480.0, 319.5, 239.5)
    for s in (1, 5):
        gray, d16, _, _ = synth.render_frame(seed=s, frame=3 * s)
        kl, _ = oracle_lib.extract_line_segments(gray, 40)
        depth = d16.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR)
        o = oracle_lib.lines3d_frame(kl, depth, icl, seed=3)
        p = host_lines3d(host_lib, kl, d16, icl, seed=3)
        for k in ("valid", "n_points", "inliers", "lines3d", "depth_line"):
            assert np.array_equal(p[k], o[k]), k
        assert p["n_drawn"] == o["n_drawn"] and p["valid"].sum() > 20
        assert (p["lines3d"][p["valid"].astype(bool)][:, [1, 4]] != 0).all()


def test_body_matches_oracle_on_corrupted_depth(host_lib):
    """Depth with outliers and noise: the RANSAC iterates (about 10 draws per line), verify3dLine rejects hypotheses, the SVD refit loop runs."""
    draws = n_valid = 0
    for s in range(8):
        gray, d16, _, _ = synth.render_frame(seed=s, frame=3 * s)
        kl, _ = oracle_lib.extract_line_segments(gray, 40)
        dn = synth.noisy_depth(d16, s, 0.1 + 0.04 * s, 0.004 * (s + 1))
        depth = dn.astype(np.float32) * np.float32(1.0 / synth.DEPTH_FACTOR)
        o = oracle_lib.lines3d_frame(kl, depth, synth.TUM3_K, seed=11 + s, skip=s)
        p = host_lines3d(host_lib, kl, dn, synth.TUM3_K, seed=11 + s, skip=s)
        assert p["n_drawn"] == o["n_drawn"], s
        for k in ("valid", "n_points", "inliers", "n_inliers", "lines3d", "depth_line"):
            assert np.array_equal(p[k], o[k]), (s, k)
        assert np.array_equal(p["director"], o["director"], equal_nan=True), s
        draws += o["n_drawn"]
        n_valid += int(p["valid"].sum())
    assert draws > 3000 and 100 < n_valid < 320
