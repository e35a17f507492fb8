"""CPU: Frame::isInFrustum(MapLine*) - oracle (oracle/linesearch.cc lines_in_frustum) vs an independent float64 numpy statement of the
geometry, and planarslam_b200/csrc/linefrustum_body.h (the per-line code of k_lines_in_frustum) compiled for the host vs the oracle
(bit-exact).  GPU run: tests/test_linefrustum_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from planarslam_b200.synth_lines import make_line_frustum

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("linefrustum") / "liblinefrustum_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
                    "-I", os.path.join(ROOT, "planarslam_b200", "csrc"), "-o", str(out), os.path.join(ROOT, "tests", "host_harness", "linefrustum_host.cc")], check=True)
    L = C.CDLL(str(out))
    L.host_lines_in_frustum.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float] + [C.c_void_p] * 4
    return L


def _numpy_frustum(frame, pos, normal, max_d, min_d, cos_limit):
    T = np.asarray(frame["Tcw"], np.float64)
    R, t = T[:3, :3], T[:3, 3]
    Ow = -R.T @ t
    sp, ep = pos[:, :3], pos[:, 3:]
    spc, epc = sp @ R.T + t, ep @ R.T + t
    with np.errstate(divide="ignore", invalid="ignore"):
        u1, v1 = frame["fx"] * spc[:, 0] / spc[:, 2] + frame["cx"], frame["fy"] * spc[:, 1] / spc[:, 2] + frame["cy"]
        u2, v2 = frame["fx"] * epc[:, 0] / epc[:, 2] + frame["cx"], frame["fy"] * epc[:, 1] / epc[:, 2] + frame["cy"]
    om = 0.5 * (sp + ep) - Ow
    dist = np.linalg.norm(om, axis=1)
    vc = (om * normal).sum(1) / dist
    ok = (spc[:, 2] >= 0) & (epc[:, 2] >= 0)
    for a, lo, hi in ((u1, "min_x", "max_x"), (v1, "min_y", "max_y"), (u2, "min_x", "max_x"), (v2, "min_y", "max_y")):
        ok &= (a >= frame[lo]) & (a <= frame[hi])
    ok &= (dist >= 0.8 * min_d) & (dist <= 1.2 * max_d) & (vc >= cos_limit)
    level = np.ceil(np.log(max_d / dist) / frame["log_scale_factor"]).astype(int)
    return ok, np.stack([u1, v1, u2, v2], 1), level, vc, dist


def test_oracle_matches_numpy_geometry():
    for seed in range(4):
        frame, pos, nrm, max_d, min_d = make_line_frustum(seed)
        o = oracle_lib.lines_in_frustum(frame, pos, nrm, max_d, min_d, 0.6)
        ok, proj, level, vc, dist = _numpy_frustum(frame, pos, nrm, max_d.astype(np.float64), min_d.astype(np.float64), 0.6)
        # float32 vs float64 only matters within a hair of a threshold
        margin = np.minimum.reduce([np.abs(vc - 0.6), np.abs(dist - 0.8 * min_d) / dist, np.abs(dist - 1.2 * max_d) / dist]) > 1e-4
        edge = ((np.abs(proj[:, [0, 2]] - 0) > 0.05) & (np.abs(proj[:, [0, 2]] - 640) > 0.05)).all(1) & ((np.abs(proj[:, [1, 3]]) > 0.05) & (np.abs(proj[:, [1, 3]] - 480) > 0.05)).all(1)
        sure = margin & edge
        assert np.array_equal(o["in_view"].astype(bool)[sure], ok[sure])
        iv = o["in_view"].astype(bool)
        assert 0.2 < iv.mean() < 0.8
        assert np.allclose(o["proj"][iv], proj[iv], rtol=0, atol=2e-3) and np.allclose(o["view_cos"][iv], vc[iv], atol=1e-5)
        frac = np.log(max_d / dist) / frame["log_scale_factor"]
        clear = iv & (np.abs(frac - np.rint(frac)) > 1e-3)
        assert np.array_equal(o["level"][clear], level[clear])
        assert (o["proj"][~iv] == 0).all() and (o["level"][~iv] == 0).all()


def test_body_matches_oracle(host_lib):
    for seed in range(6):
        frame, pos, nrm, max_d, min_d = make_line_frustum(seed, n=1000)
        o = oracle_lib.lines_in_frustum(frame, pos, nrm, max_d, min_d, 0.6)
        fv = np.concatenate([np.asarray(frame["Tcw"], np.float32).ravel(), np.array([frame[k] for k in ("fx", "fy", "cx", "cy", "min_x", "max_x", "min_y", "max_y",
                                                                                                          "log_scale_factor")], np.float32)])
        P, Nn = np.ascontiguousarray(pos, np.float64), np.ascontiguousarray(nrm, np.float64)
        iv, proj, level, vc = np.zeros(1000, np.uint8), np.zeros((1000, 4), np.float32), np.zeros(1000, np.int32), np.zeros(1000, np.float32)
        cnt = host_lib.host_lines_in_frustum(fv.ctypes.data, 1000, P.ctypes.data, Nn.ctypes.data, max_d.ctypes.data, min_d.ctypes.data, 0.6, iv.ctypes.data,
                                             proj.ctypes.data, level.ctypes.data, vc.ctypes.data)
        assert cnt == int(o["in_view"].sum())
        assert np.array_equal(iv, o["in_view"]) and np.array_equal(proj, o["proj"]) and np.array_equal(level, o["level"]) and np.array_equal(vc, o["view_cos"])
