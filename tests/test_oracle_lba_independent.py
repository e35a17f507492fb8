"""CPU: the LBA restatement (oracle/lba.cc) checked against an independently coded cost function (numpy): on a small, outlier-free
stereo problem the oracle's result must (a) lower the weighted reprojection cost below the cost at the ground truth (the
observations are noisy, so the truth is not the minimiser), (b) report a final chi2 that the independent cost reproduces, and
(c) have a much smaller numerical gradient (free poses as left-multiplicative se(3) increments, points) than the starting
point - not zero: g2o's modified LM stops after three iterations that each gain less than 0.1 % (optimization_algorithm_
levenberg.cpp:155-161).  This pins the residual model, the information weighting and the descent of the Schur-complement LM."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth_lba


def _exp_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def _cost(p, T, X):
    """sum over stereo observations of invSigma2 * |obs - proj|^2 (the second, kernel-free pass of LocalBundleAdjustment)."""
    fx, fy, cx, cy, bf = (float(v) for v in p["kf_K"][0])
    Xc = np.einsum("nij,nj->ni", T[p["pt_obs_kf"], :3, :3], X[p["pt_obs_pt"]]) + T[p["pt_obs_kf"], :3, 3]
    u = fx * Xc[:, 0] / Xc[:, 2] + cx
    v = fy * Xc[:, 1] / Xc[:, 2] + cy
    ur = u - bf / Xc[:, 2]
    r = p["pt_obs_uvr"].astype(np.float64) - np.stack([u, v, ur], 1)
    return float((p["pt_obs_inv_sigma2"].astype(np.float64)[:, None] * r * r).sum())


def test_lba_result_is_a_stationary_point_of_an_independent_cost():
    p = synth_lba.make_lba_problem(21, n_kf=4, n_fixed=1, n_points=60, n_pt_obs=220, n_lines=0, n_line_obs=0, n_plane_obs=(0, 0, 0),
                                   outlier_frac=0.0, mono_frac=0.0, kf_stride=6)
    r = oracle_lib.local_bundle_adjustment(p)
    assert r["erase_pt"].sum() <= 0.1 * len(r["erase_pt"])
    keep = r["erase_pt"] == 0                                   # the second pass optimises the edges that survived the chi-square gate
    q = dict(p)
    for k in ("pt_obs_kf", "pt_obs_pt", "pt_obs_uvr", "pt_obs_inv_sigma2"):
        q[k] = p[k][keep]
    T, X = r["kf_Tcw_d"].copy(), r["pt_Xw_d"].copy()
    c_opt = _cost(q, T, X)
    c_true = _cost(q, p["kf_Tcw_true"], p["pt_Xw_true"])
    c_init = _cost(q, p["kf_Tcw"].astype(np.float64), p["pt_Xw"].astype(np.float64))
    assert c_opt < c_true < c_init, (c_opt, c_true, c_init)
    assert abs(c_opt - r["chi2"][1]) < 0.02 * c_opt, (c_opt, r["chi2"])      # same cost up to the edges re-classified after the pass

    def grad(T, X):
        h = 1e-6
        g = []
        for k in range(1, 4):                                    # free key frames
            for d in range(6):
                e = np.zeros(6)
                e[d] = h
                vals = []
                for s in (+1, -1):
                    Tk = T.copy()
                    dT = np.eye(4)
                    dT[:3, :3] = _exp_so3(s * e[:3])
                    dT[:3, 3] = s * e[3:]
                    Tk[k] = dT @ T[k]
                    vals.append(_cost(q, Tk, X))
                g.append((vals[0] - vals[1]) / (2 * h))
        for i in np.unique(q["pt_obs_pt"])[:25]:
            for d in range(3):
                vals = []
                for s in (+1, -1):
                    Xk = X.copy()
                    Xk[i, d] += s * h
                    vals.append(_cost(q, T, Xk))
                g.append((vals[0] - vals[1]) / (2 * h))
        return np.array(g)

    g_init = grad(p["kf_Tcw"].astype(np.float64), p["pt_Xw"].astype(np.float64))
    g_opt = grad(T, X)
    assert np.linalg.norm(g_opt) < 0.05 * np.linalg.norm(g_init), (np.linalg.norm(g_opt), np.linalg.norm(g_init))


def _plane_norm(c):
    c = np.asarray(c, np.float64) / np.linalg.norm(c[:3])
    return -c if c[3] < 0 else c


def _plane_rot(v):
    az, el = np.arctan2(v[1], v[0]), np.arctan2(v[2], np.hypot(v[0], v[1]))
    cz, sz, cy, sy = np.cos(az), np.sin(az), np.cos(-el), np.sin(-el)
    return np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])


def _cost_all(p, T, X, Lw, Pw, keep_pt, keep_line, keep_plane):
    """Points (as _cost) + line observations (two squared point-line distances each, information = identity, include/EdgeLine.h:53-153) +
    plane observations of type 0 (EdgePlane, g2oAddition/EdgePlane.h:25-126: (T * plane).ominus(measurement), information
    diag(a, a, d) with a = 3282.8 / AngleInfo^2, d = DistanceInfo^2, src/Optimizer.cc:2236-2239)."""
    q = dict(p)
    for k in ("pt_obs_kf", "pt_obs_pt", "pt_obs_uvr", "pt_obs_inv_sigma2"):
        q[k] = p[k][keep_pt]
    c = _cost(q, T, X)
    fx, fy, cx, cy, _ = (float(v) for v in p["kf_K"][0])
    for j in np.nonzero(keep_line)[0]:
        k, li, l = p["line_obs_kf"][j], p["line_obs_line"][j], p["line_obs_l"][j]
        for a in (0, 3):
            Pc = T[k, :3, :3] @ Lw[li, a:a + 3] + T[k, :3, 3]
            c += float(l[0] * (fx * Pc[0] / Pc[2] + cx) + l[1] * (fy * Pc[1] / Pc[2] + cy) + l[2]) ** 2
    w_ang, w_dis = 3282.8 / p["angle_info"] ** 2, p["dist_info"] ** 2
    for j in np.nonzero(keep_plane)[0]:
        k, pl = p["plane_obs_kf"][0][j], p["plane_obs_plane"][0][j]
        pw = _plane_norm(Pw[pl])
        n = T[k, :3, :3] @ pw[:3]
        local = _plane_norm(np.array([n[0], n[1], n[2], pw[3] - T[k, :3, 3] @ n]))
        meas = _plane_norm(p["plane_obs_meas"][0][j])
        m = _plane_rot(local[:3]).T @ meas[:3]
        e = np.array([np.arctan2(m[1], m[0]), np.arctan2(m[2], np.hypot(m[0], m[1])), (-local[3]) - (-meas[3])])
        c += w_ang * (e[0] ** 2 + e[1] ** 2) + w_dis * e[2] ** 2
    return c


def test_lba_with_lines_and_planes_is_near_stationary_in_the_poses():
    """Points + lines + plane edges: at the oracle's result the gradient of the independent cost with respect to the free key-frame poses
    (landmarks held at their optimised values) is far smaller than at the start."""
    p = synth_lba.make_lba_problem(33, n_kf=4, n_fixed=1, n_points=60, n_pt_obs=200, n_lines=12, n_line_obs=30, n_plane_obs=(8, 0, 0),
                                   outlier_frac=0.0, mono_frac=0.0, kf_stride=6, line_norm3=False)
    r = oracle_lib.local_bundle_adjustment(p)
    kp, kl, kpl = r["erase_pt"] == 0, r["erase_line"] == 0, r["erase_plane"][0] == 0
    assert kp.mean() > 0.9 and kl.mean() > 0.8 and kpl.all()

    def total(T, X, Lw, Pw):
        return _cost_all(p, T, X, Lw, Pw, kp, kl, kpl)

    def grad(T, X, Lw, Pw, h=1e-6):
        g = []
        for k in range(1, 4):
            for d in range(6):
                vals = []
                for s in (+1, -1):
                    e = np.zeros(6)
                    e[d] = s * h
                    dT = np.eye(4)
                    dT[:3, :3] = _exp_so3(e[:3])
                    dT[:3, 3] = e[3:]
                    Tk = T.copy()
                    Tk[k] = dT @ T[k]
                    vals.append(total(Tk, X, Lw, Pw))
                g.append((vals[0] - vals[1]) / (2 * h))
        return np.array(g)
    init = (p["kf_Tcw"].astype(np.float64), p["pt_Xw"].astype(np.float64), p["line_Xw"].astype(np.float64), p["plane_Xw"].astype(np.float64))
    opt = (r["kf_Tcw_d"], r["pt_Xw_d"], r["line_Xw_d"], r["plane_Xw_d"])
    assert total(*opt) < 0.05 * total(*init)
    g0, g1 = grad(*init), grad(*opt)
    assert np.linalg.norm(g1) < 0.02 * np.linalg.norm(g0), (np.linalg.norm(g0), np.linalg.norm(g1))
