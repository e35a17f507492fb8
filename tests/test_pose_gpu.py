"""GPU: Optimizer::PoseOptimization through the C ABI vs the CPU oracle.

Bar (BASELINE.json north_star): SE3 pose within 1e-4 rad / 1e-3 m of the reference path after the same LM iteration
count.  Tolerances used here are much tighter (the two implementations differ only in summation order and libm ulps):
rotation 1e-7 rad, translation 1e-7 m, identical outlier flags, identical per-round LM iteration / trial counts."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth_pose

pytestmark = pytest.mark.gpu
ROT_TOL, TRANS_TOL = 1e-7, 1e-7            # rad, m   (requirement: 1e-4 rad, 1e-3 m)


def _compare(r, o, p, tag="", tol_scale=1.0):
    er, et = synth_pose.pose_error(r["Tcw_d"], o["Tcw_d"])
    assert er < ROT_TOL * tol_scale and et < TRANS_TOL * tol_scale, (tag, er, et, r["trace_i"].tolist(), o["trace_i"].tolist(), r["trace_d"].tolist(), o["trace_d"].tolist())
    assert np.abs(r["Tcw"] - o["Tcw"]).max() < 1e-6 * tol_scale
    assert r["n_inliers"] == o["n_inliers"]
    for k in ("outlier_pt", "outlier_line", "outlier_plane", "outlier_par", "outlier_ver"):
        assert np.array_equal(r[k], o[k]), k
    # Same number of outliers in every round and the same converged robust chi2.  The LM iteration / trial counts are
    # compared with a slack of one iteration: at convergence g2o's accept test is the sign of a chi2 difference of the
    # order of 1e-12 * chi2, which depends on the summation order (fixed tree on the GPU, sequential in the oracle); an
    # extra or missing final iteration moves the pose by < 1e-9 (checked above with ROT_TOL / TRANS_TOL).
    assert np.array_equal(r["trace_i"][:, 2], o["trace_i"][:, 2]), (r["trace_i"], o["trace_i"])
    assert np.abs(r["trace_i"][:, 0] - o["trace_i"][:, 0]).max() <= 1, (r["trace_i"], o["trace_i"])
    assert np.allclose(r["trace_d"][:, 0], o["trace_d"][:, 0], rtol=1e-6 * tol_scale, atol=1e-9), (tag, r["trace_d"], o["trace_d"])

def test_pose_optimization_matches_oracle():
    from planarslam_b200.optimizer import Optimizer
    opt = Optimizer()
    probs = [synth_pose.make_pose_problem(s, frame=5 * s) for s in range(8)]
    res = opt.PoseOptimizationBatch(probs)
    for p, r in zip(probs, res):
        o = oracle_lib.pose_optimization(p)
        _compare(r, o, p)
        e_true = synth_pose.pose_error(r["Tcw_d"], p["Tcw_true"])
        assert e_true[0] < 3e-3 and e_true[1] < 5e-3           # and it actually converged to the ground truth


def test_single_call_and_edge_mixes():
    from planarslam_b200.optimizer import Optimizer
    opt = Optimizer()
    cases = [dict(n_points=1000, n_lines=40, n_planes=3, n_par=1, n_ver=2),
             dict(n_points=300, n_lines=0, n_planes=0, n_par=0, n_ver=0),          # points only
             dict(n_points=60, n_lines=40, n_planes=3, n_par=0, n_ver=0),
             dict(n_points=5000, n_lines=100, n_planes=6, n_par=2, n_ver=3, outlier_frac=0.15),
             dict(n_points=2, n_lines=0, n_planes=0, n_par=0, n_ver=0),            # < 3 correspondences -> returns 0, pose untouched
             dict(n_points=4, n_lines=1, n_planes=1, n_par=0, n_ver=0)]            # < 10 edges -> one round only
    for i, kw in enumerate(cases):
        p = synth_pose.make_pose_problem(100 + i, frame=i, **kw)
        n, r = opt.PoseOptimization(p)
        o = oracle_lib.pose_optimization(p)
        assert n == o["n_inliers"], (i, n, o["n_inliers"])
        # case 5 has 7 edges for 6 unknowns: the normal equations are so poorly conditioned that the summation order of H
        # shows up at the 1e-7 level (measured 2.8e-7 rad / 6.9e-7 m); still 100x inside the required 1e-4 rad / 1e-3 m
        _compare(r, o, p, tag=f"case {i}", tol_scale=100.0 if i == 5 else 1.0)
    p = synth_pose.make_pose_problem(7, n_points=2, n_lines=0, n_planes=0, n_par=0, n_ver=0)
    n, r = opt.PoseOptimization(p)
    assert n == 0 and np.array_equal(r["Tcw"], p["Tcw0"])


def test_translation_optimization_matches_oracle():
    """Optimizer::TranslationOptimization (rotation frozen, mapTrans edges): same bar as PoseOptimization."""
    from planarslam_b200.optimizer import Optimizer
    opt = Optimizer()
    probs = [synth_pose.make_pose_problem(40 + s, frame=3 * s, rot_pert=0.0, trans_pert=0.05) for s in range(6)]
    probs.append(synth_pose.make_pose_problem(50, n_points=2, n_lines=5, n_planes=3))                 # < 3 points: returns 0 before planes are added
    probs.append(synth_pose.make_pose_problem(51, n_points=200, n_lines=0, n_planes=0, rot_pert=0.0))
    res = opt.TranslationOptimizationBatch(probs)
    for p, r in zip(probs, res):
        o = oracle_lib.translation_optimization(p)
        _compare(r, o, p)
        # the rotation block is untouched (all rotation Jacobian columns are zero)
        assert np.abs(r["Tcw_d"][:3, :3] - o["Tcw_d"][:3, :3]).max() < 1e-12
    n, r = opt.TranslationOptimization(probs[0])
    assert n == oracle_lib.translation_optimization(probs[0])["n_inliers"]
    assert synth_pose.pose_error(r["Tcw_d"], probs[0]["Tcw_true"])[1] < 5e-3
