"""GPU: projection-guided searches through the C ABI vs the CPU oracle — identical match assignments."""
import numpy as np
import pytest

import oracle_lib
from test_oracle_search import scenario

pytestmark = pytest.mark.gpu


def test_search_by_projection_map_matches_oracle():
    from planarslam_b200.matcher import ORBmatcher
    m8 = ORBmatcher(0.8)
    for (f0, f1, th, noise) in [(10, 11, 3.0, 0.002), (20, 22, 5.0, 0.01), (30, 31, 1.0, 0.0), (40, 44, 3.0, 0.02)]:
        fv, m, lf = scenario(f0, f1, noise, seed=f0)
        rng = np.random.default_rng(f0)
        pre = np.full(fv["n"], -1, np.int32)
        pre[rng.choice(fv["n"], 50, replace=False)] = rng.integers(0, m["n"], 50)      # matches from an earlier stage
        m["has_obs"][rng.choice(m["n"], m["n"] // 10, replace=False)] = 0              # a few fresh points without observations
        m["skip"][rng.choice(m["n"], m["n"] // 20, replace=False)] = 1
        n, matches, in_view = m8.SearchByProjection(fv, m, th, pre)
        on, omatches, oin = oracle_lib.search_by_projection_map(fv, m, th, 0.8, pre)
        assert np.array_equal(in_view, oin), (f0, f1)
        assert n == on and np.array_equal(matches, omatches), (f0, f1, n, on, (matches != omatches).sum())
        assert n > 50


def test_search_by_projection_last_matches_oracle():
    from planarslam_b200.matcher import ORBmatcher
    for check_ori in (True, False):
        mm = ORBmatcher(0.9, check_ori)
        for (f0, f1, th, mono) in [(10, 11, 15.0, False), (20, 23, 15.0, False), (30, 31, 7.0, True)]:
            fv, m, lf = scenario(f0, f1, 0.005, seed=f1)
            lf["outlier"][::17] = 1
            n, matches = mm.SearchByProjectionLast(fv, lf, m, th, mono)
            on, omatches = oracle_lib.search_by_projection_last(fv, lf, m, th, mono, check_ori, np.full(fv["n"], -1, np.int32))
            assert n == on and np.array_equal(matches, omatches), (f0, f1, check_ori, n, on)
            assert n > 50


def test_plane_matcher_matches_oracle():
    import ctypes as C
    from planarslam_b200.matcher import PlaneMatcher
    from planarslam_b200 import synth_map
    L = oracle_lib.lib()
    L.orc_plane_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p] + [C.c_float] * 4 + [C.c_void_p] * 3
    pm = PlaneMatcher(0.05, 0.985, 0.08716, 0.9962)          # TUM3.yaml association thresholds
    rng = np.random.default_rng(3)
    for trial in range(6):
        T = synth_map.true_pose(5 * trial).astype(np.float32)
        R, t = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
        world = [np.array([0, 1, 0, -1.2]), np.array([1, 0, 0, 1.6]), np.array([0, 0, 1, -3.2]), np.array([0, 1, 0, -0.4]), np.array([0.6, 0.8, 0, -1.0])]
        n_map = len(world) + trial
        mc = np.array([world[k % len(world)] + (0.3 * (k // len(world)) * np.array([0, 0, 0, 1])) for k in range(n_map)], np.float32)
        bad = (rng.random(n_map) < 0.15).astype(np.uint8)
        # frame planes: the first three world planes seen from the camera, slightly perturbed
        fc = []
        for k in range(3):
            n_c = R @ world[k][:3]
            d_c = world[k][3] - t @ n_c
            fc.append(np.concatenate([n_c + rng.normal(0, 0.01, 3), [d_c + rng.normal(0, 0.005)]]))
        fc = np.array(fc, np.float32)
        cnt = rng.integers(0, 200, n_map)
        off = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
        pts = np.zeros((off[-1], 3), np.float32)
        for j in range(n_map):                          # points scattered on each map plane (plus noise)
            nrm, d = mc[j, :3].astype(np.float64), float(mc[j, 3])
            q = rng.normal(0, 1.5, (cnt[j], 3))
            q -= np.outer(q @ nrm + d, nrm) / (nrm @ nrm)
            pts[off[j]:off[j + 1]] = q + rng.normal(0, 0.01, q.shape)
        n, m, v, p = pm.SearchMapByCoefficients(T, fc, mc, bad, off, pts)
        om, ov, op = [np.zeros(3, np.int32) for _ in range(3)]
        on = L.orc_plane_match(T.ctypes.data, 3, fc.ctypes.data, n_map, mc.ctypes.data, bad.ctypes.data, off.ctypes.data, pts.ctypes.data,
                               0.05, 0.985, 0.08716, 0.9962, om.ctypes.data, ov.ctypes.data, op.ctypes.data)
        assert n == on and np.array_equal(m, om) and np.array_equal(v, ov) and np.array_equal(p, op), (trial, m, om, v, ov, p, op)
