"""CPU: the plane-association oracle (oracle/planematch.cc) against an independent plain-Python statement of
PlaneMatcher::SearchMapByCoefficients / PointDistanceFromPlane (src/PlaneMatcher.cpp:10-82) and Frame::ComputePlaneWorldCoeff (src/Frame.cc:815-820),
written from the reference with float32 scalars (the reference ships no vectors for it)."""
import ctypes as C

import numpy as np

import oracle_lib
from planarslam_b200 import synth_map


def _scenario(trial, rng):
    T = synth_map.true_pose(5 * trial).astype(np.float32)
    R, t = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
    world = [np.array([0, 1, 0, -1.2]), np.array([1, 0, 0, 1.6]), np.array([0, 0, 1, -3.2]), np.array([0, 1, 0, -0.4]), np.array([0.6, 0.8, 0, -1.0])]
    n_map = len(world) + trial
    mc = np.array([world[k % len(world)] + (0.3 * (k // len(world)) * np.array([0, 0, 0, 1])) for k in range(n_map)], np.float32)
    bad = (rng.random(n_map) < 0.15).astype(np.uint8)
    fc = []
    for k in range(3):
        n_c = R @ world[k][:3]
        d_c = world[k][3] - t @ n_c
        fc.append(np.concatenate([n_c + rng.normal(0, 0.01, 3), [d_c + rng.normal(0, 0.005)]]))
    fc = np.array(fc, np.float32)
    cnt = rng.integers(0, 200, n_map)
    off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    pts = np.zeros((off[-1], 3), np.float32)
    for j in range(n_map):
        nrm, d = mc[j, :3].astype(np.float64), float(mc[j, 3])
        q = rng.normal(0, 1.5, (cnt[j], 3))
        q -= np.outer(q @ nrm + d, nrm) / (nrm @ nrm)
        pts[off[j]:off[j + 1]] = q + rng.normal(0, 0.01 + 0.02 * (j % 3), q.shape)
    return T, fc, mc, bad, off, pts


def _py_plane_match(T, fc, mc, bad, off, pts, dTh, aTh, verTh, parTh):
    f = np.float32
    n_frame, n_map = len(fc), len(mc)
    match, ver, par = (np.full(n_frame, -1, np.int32) for _ in range(3))
    nmatches = 0
    for i in range(n_frame):
        pM = np.array([f(sum(float(T[k, r]) * float(fc[i][k]) for k in range(4))) for r in range(4)], np.float32)      # mTcw^T * coefficients
        ld, lver, lpar = f(dTh), f(verTh), f(parTh)
        found = False
        for j in range(n_map):
            if bad[j]:
                continue
            pW = mc[j]
            angle = f(f(f(pM[0] * pW[0]) + f(pM[1] * pW[1])) + f(pM[2] * pW[2]))
            if angle > f(aTh) or angle < -f(aTh):
                res = 100.0
                for q in pts[off[j]:off[j + 1]]:
                    dis = float(abs(f(f(f(f(pM[0] * q[0]) + f(pM[1] * q[1])) + f(pM[2] * q[2])) + pM[3])))
                    if dis < res:
                        res = dis
                if res < float(ld):
                    ld = f(res)
                    match[i] = j
                    found = True
                    continue
            if angle < lver and angle > -lver:
                lver = f(abs(angle))
                ver[i] = j
                continue
            if angle > lpar or angle < -lpar:
                lpar = f(abs(angle))
                par[i] = j
        nmatches += int(found)
    return nmatches, match, ver, par


def test_plane_match_oracle_matches_independent_python():
    L = oracle_lib.lib()
    L.orc_plane_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p] * 3
    rng = np.random.default_rng(3)
    th = (0.05, 0.985, 0.08716, 0.9962)                       # TUM3.yaml association thresholds
    total = 0
    for trial in range(10):
        T, fc, mc, bad, off, pts = _scenario(trial, rng)
        om, ov, op = [np.zeros(len(fc), np.int32) for _ in range(3)]
        on = L.orc_plane_match(T.ctypes.data, len(fc), fc.ctypes.data, len(mc), mc.ctypes.data, bad.ctypes.data, off.ctypes.data, pts.ctypes.data, *th,
                               om.ctypes.data, ov.ctypes.data, op.ctypes.data)
        pn, pm, pv, pp = _py_plane_match(T, fc, mc, bad, off, pts, *th)
        assert on == pn and np.array_equal(om, pm) and np.array_equal(ov, pv) and np.array_equal(op, pp), (trial, om, pm, ov, pv, op, pp)
        total += on
    assert total >= 10
