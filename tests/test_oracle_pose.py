"""CPU: the PoseOptimization / TranslationOptimization oracle (oracle/poseopt.cc; src/Optimizer.cc:550-1275, :2995-3737 with the vendored g2o LM).
The reference ships no vectors for this path and its g2o / Eigen sources cannot be compiled here, so the restatement is
checked against the problem's ground truth and against an independently coded numpy cost:
  * the optimised pose is close to the truth and much closer than the initial guess; the injected outliers are flagged;
  * at the result, the gradient of the (kernel-free, inlier-only) reprojection + line cost - what the last of the four rounds minimises - is
    far smaller than at the start;
  * TranslationOptimization leaves the rotation untouched; fewer than three correspondences return 0 and keep the pose."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth_pose


def _exp_so3(w):
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        return np.eye(3) + K
    return np.eye(3) + np.sin(th) / th * K + (1 - np.cos(th)) / th ** 2 * K @ K


def _cost(p, T, in_pt, in_line):
    """sum over inlier point observations of invSigma2 |obs - proj|^2 (2 or 3 components) + sum over inlier lines of the two squared
    point-line distances (EdgeLineProjectXYZOnlyPose, information = identity)."""
    fx, fy, cx, cy, bf = p["fx"], p["fy"], p["cx"], p["cy"], p["bf"]
    R, t = T[:3, :3], T[:3, 3]
    Xc = p["Xw"].astype(np.float64) @ R.T + t
    u, v = fx * Xc[:, 0] / Xc[:, 2] + cx, fy * Xc[:, 1] / Xc[:, 2] + cy
    obs = p["obs"].astype(np.float64)
    r2 = (obs[:, 0] - u) ** 2 + (obs[:, 1] - v) ** 2
    stereo = obs[:, 2] >= 0
    r2 = r2 + np.where(stereo, (obs[:, 2] - (u - bf / Xc[:, 2])) ** 2, 0.0)
    c = float((p["inv_sigma2"].astype(np.float64) * r2)[in_pt].sum())
    if len(p["line_obs"]):
        for k in (0, 3):
            Pc = p["line_Xw"][:, k:k + 3] @ R.T + t
            d = p["line_obs"][:, 0] * (fx * Pc[:, 0] / Pc[:, 2] + cx) + p["line_obs"][:, 1] * (fy * Pc[:, 1] / Pc[:, 2] + cy) + p["line_obs"][:, 2]
            c += float((d * d)[in_line].sum())
    return c


def _grad(p, T, in_pt, in_line, h=1e-6):
    g = np.zeros(6)
    for d in range(6):
        vals = []
        for s in (+1, -1):
            e = np.zeros(6)
            e[d] = s * h
            dT = np.eye(4)
            dT[:3, :3] = _exp_so3(e[:3])
            dT[:3, 3] = e[3:]
            vals.append(_cost(p, dT @ T, in_pt, in_line))
        g[d] = (vals[0] - vals[1]) / (2 * h)
    return g


def test_pose_oracle_converges_to_the_truth_and_flags_outliers():
    for seed in range(4):
        p = synth_pose.make_pose_problem(seed, frame=3 * seed)
        r = oracle_lib.pose_optimization(p)
        a0, t0 = synth_pose.pose_error(p["Tcw0"], p["Tcw_true"])
        a1, t1 = synth_pose.pose_error(r["Tcw_d"], p["Tcw_true"])
        assert a1 < 0.004 and t1 < 0.01 and a1 < 0.2 * a0 and t1 < 0.4 * t0, (seed, a0, t0, a1, t1)
        # the injected gross outliers (50 random re-positioned observations) are flagged, few good ones are
        T = p["Tcw_true"]
        Xc = p["Xw"].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
        err = np.hypot(p["fx"] * Xc[:, 0] / Xc[:, 2] + p["cx"] - p["obs"][:, 0], p["fy"] * Xc[:, 1] / Xc[:, 2] + p["cy"] - p["obs"][:, 1])
        gross = err > 30
        assert gross.sum() >= 40 and r["outlier_pt"][gross].mean() > 0.95 and r["outlier_pt"][~gross].mean() < 0.05
        assert r["n_inliers"] > 900
        assert np.array_equal(r["Tcw"], r["Tcw_d"].astype(np.float32))


def test_pose_result_is_near_stationary_for_an_independent_cost():
    for seed in (1, 5):
        p = synth_pose.make_pose_problem(seed, frame=7, n_planes=0, n_par=0, n_ver=0)
        r = oracle_lib.pose_optimization(p)
        in_pt, in_line = r["outlier_pt"] == 0, r["outlier_line"] == 0
        T0, T1 = p["Tcw0"].astype(np.float64), r["Tcw_d"]
        c0, c1, ct = _cost(p, T0, in_pt, in_line), _cost(p, T1, in_pt, in_line), _cost(p, p["Tcw_true"], in_pt, in_line)
        assert c1 < ct < c0, (c0, c1, ct)                         # noisy observations: the truth is not the minimiser
        g0, g1 = _grad(p, T0, in_pt, in_line), _grad(p, T1, in_pt, in_line)
        assert np.linalg.norm(g1) < 1e-3 * np.linalg.norm(g0), (np.linalg.norm(g0), np.linalg.norm(g1))


def test_translation_optimization_keeps_the_rotation():
    p = synth_pose.make_pose_problem(2, frame=5, rot_pert=0.0)
    r = oracle_lib.pose_optimization(p, translation_only=True)
    assert np.array_equal(r["Tcw"][:3, :3], p["Tcw0"][:3, :3])
    _, t0 = synth_pose.pose_error(p["Tcw0"], p["Tcw_true"])
    _, t1 = synth_pose.pose_error(r["Tcw_d"], p["Tcw_true"])
    assert t1 < 0.3 * t0 and t1 < 0.01


def test_fewer_than_three_correspondences_return_zero():
    p = synth_pose.make_pose_problem(3, frame=2, n_points=2, n_lines=0, n_planes=0, n_par=0, n_ver=0, outlier_frac=0.0)
    r = oracle_lib.pose_optimization(p)
    assert r["n_inliers"] == 0 and np.array_equal(r["Tcw"], p["Tcw0"])


def _plane_norm(c):
    c = np.asarray(c, np.float64) / np.linalg.norm(c[:3])
    return -c if c[3] < 0 else c                                 # Plane3D::normalize (g2oAddition/Plane3D.h:175-180): unit normal, last coefficient >= 0


def _plane_to_camera(T, pw):
    """operator*(Isometry3D, Plane3D) (g2oAddition/Plane3D.h:186-199)."""
    pw = _plane_norm(pw)
    n = T[:3, :3] @ pw[:3]
    return _plane_norm(np.array([n[0], n[1], n[2], pw[3] - T[:3, 3] @ n]))


def _plane_rot(v):
    """Plane3D::rotation (:76-82): R = Rz(azimuth) * Ry(-elevation)."""
    az, el = np.arctan2(v[1], v[0]), np.arctan2(v[2], np.hypot(v[0], v[1]))
    cz, sz, cy, sy = np.cos(az), np.sin(az), np.cos(-el), np.sin(-el)
    return np.array([[cz, -sz, 0], [sz, cz, 0], [0, 0, 1]]) @ np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])


def _plane_cost(p, T, inl):
    """EdgePlaneOnlyPose (g2oAddition/EdgePlane.h:128-224): e = (T * map).ominus(measurement), information diag(angleInfo, angleInfo, disInfo)
    with the settings converted as src/Optimizer.cc:771-776 does: angleInfo = 3282.8 / Plane.AngleInfo^2 (degrees -> rad), disInfo = Plane.DistanceInfo^2."""
    w_ang, w_dis = 3282.8 / p["angle_info"] ** 2, p["dist_info"] ** 2
    c = 0.0
    for i in np.nonzero(inl)[0]:
        local, meas = _plane_to_camera(T, p["plane_map"][i]), _plane_norm(p["plane_meas"][i])
        n = _plane_rot(local[:3]).T @ meas[:3]
        e = np.array([np.arctan2(n[1], n[0]), np.arctan2(n[2], np.hypot(n[0], n[1])), (-local[3]) - (-meas[3])])
        c += w_ang * (e[0] ** 2 + e[1] ** 2) + w_dis * e[2] ** 2
    return c


def test_pose_result_with_plane_edges_is_near_stationary_for_an_independent_cost():
    p = synth_pose.make_pose_problem(6, frame=9, n_points=60, n_lines=0, n_planes=3, n_par=0, n_ver=0, outlier_frac=0.0)
    r = oracle_lib.pose_optimization(p)
    in_pt, in_pl = r["outlier_pt"] == 0, r["outlier_plane"] == 0
    assert in_pl.all()

    def total(T):
        return _cost(p, T, in_pt, np.zeros(0, bool)) + _plane_cost(p, T, in_pl)

    def grad(T, h=1e-6):
        g = np.zeros(6)
        for d in range(6):
            v = []
            for sgn in (+1, -1):
                e = np.zeros(6)
                e[d] = sgn * h
                dT = np.eye(4)
                dT[:3, :3] = _exp_so3(e[:3])
                dT[:3, 3] = e[3:]
                v.append(total(dT @ T))
            g[d] = (v[0] - v[1]) / (2 * h)
        return g
    T0, T1 = p["Tcw0"].astype(np.float64), r["Tcw_d"]
    assert total(T1) < total(p["Tcw_true"]) < total(T0)
    assert _plane_cost(p, T1, in_pl) < 0.2 * _plane_cost(p, T0, in_pl) and _plane_cost(p, T0, in_pl) > 0.5 * _cost(p, p["Tcw_true"], in_pt, np.zeros(0, bool))
    assert np.linalg.norm(grad(T1)) < 2e-3 * np.linalg.norm(grad(T0)), (np.linalg.norm(grad(T0)), np.linalg.norm(grad(T1)))
