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
