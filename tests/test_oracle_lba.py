"""CPU: the oracle restatement of Optimizer::LocalBundleAdjustment (oracle/lba.cc) behaves like the reference's
g2o pipeline should on seeded synthetic local maps.  The reference ships no golden vectors for this path ("parity
unpinned", see oracle/lba.h), so these tests pin the restatement through properties:
  * fixed key frames come back unchanged, everything is finite, the problem converges towards the ground truth;
  * in the second optimize() (no robust kernel) the LM gain ratio is ~1, i.e. the Jacobians, the Schur complement and
    the reduced solve are mutually consistent with the cost that is actually evaluated (checked through the final
    lambda: every accepted step with rho ~ 1 divides lambda by 3);
  * the erase lists contain the planted gross outliers."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth_lba, synth_pose


def _errs(T, Ttrue):
    return np.array([synth_pose.pose_error(T[k], Ttrue[k]) for k in range(len(T))])


def test_lba_oracle_converges_and_keeps_fixed_frames():
    p = synth_lba.make_lba_problem(3, n_kf=8, n_fixed=2, n_points=400, n_pt_obs=2400, n_lines=20, n_line_obs=40, n_plane_obs=(8, 2, 2))
    r = oracle_lib.local_bundle_adjustment(p)
    assert np.isfinite(r["kf_Tcw_d"]).all() and np.isfinite(r["pt_Xw_d"]).all() and np.isfinite(r["plane_Xw_d"]).all()
    # fixed frames: the estimate is only converted float -> quaternion -> matrix
    assert np.abs(r["kf_Tcw"][:2] - p["kf_Tcw"][:2]).max() < 1e-6
    e0, e1 = _errs(p["kf_Tcw"], p["kf_Tcw_true"]), _errs(r["kf_Tcw_d"], p["kf_Tcw_true"])
    assert e1[2:, 0].mean() < 0.6 * e0[2:, 0].mean() and e1[2:, 1].mean() < 0.8 * e0[2:, 1].mean(), (e0, e1)
    assert r["iterations"][0] == 5 and 1 <= r["iterations"][1] <= 10
    assert r["chi2"][1] < r["chi2"][0]
    # 5 % planted gross point outliers (15-60 px): nearly all of them must be in the erase list
    assert 0.04 * len(r["erase_pt"]) < r["erase_pt"].sum() < 0.15 * len(r["erase_pt"])


def test_lba_oracle_second_pass_gain_ratio_is_one():
    p = synth_lba.make_lba_problem(5, n_kf=6, n_fixed=1, n_points=300, n_pt_obs=1500, n_lines=0, n_line_obs=0, n_plane_obs=(0, 0, 0),
                                   outlier_frac=0.0, mono_frac=0.0)
    r = oracle_lib.local_bundle_adjustment(p)
    # every first trial accepted (trials == iterations) and lambda shrinks by the maximal factor 1/3 per iteration <=> rho >= ~0.79
    assert r["trials"][1] == r["iterations"][1]
    # no planted outliers: only the ~5 % tail of the chi-square(3) threshold 7.815 is erased
    assert r["erase_pt"].sum() <= 0.08 * len(r["erase_pt"])


def test_lba_oracle_gates_and_quirk_inputs():
    # wide baseline, 10 observations per line, pixel-scaled line functions, wrong lines and wrong planes planted:
    # the line gate (both endpoints jointly), the plane gates and a rejected LM trial (pop) are all exercised
    p = synth_lba.make_lba_problem(8, n_kf=12, n_fixed=2, n_points=300, n_pt_obs=2400, n_lines=4, n_line_obs=40, n_plane_obs=(12, 4, 4),
                                   line_norm3=False, outlier_frac=0.2, kf_stride=5, plane_outlier_frac=0.25)
    r = oracle_lib.local_bundle_adjustment(p)
    assert r["erase_line"].sum() >= 1
    assert sum(int(a.sum()) for a in r["erase_plane"]) >= 1
    assert r["trials"][0] > r["iterations"][0] or r["trials"][1] > r["iterations"][1]
    q = synth_lba.make_lba_problem(7, n_kf=5, n_fixed=1, n_points=200, n_pt_obs=900, n_lines=30, n_line_obs=60, n_plane_obs=(6, 2, 1),
                                   line_kf_quirk=True)
    r2 = oracle_lib.local_bundle_adjustment(q)            # every line edge attached to the current key frame (reference quirk)
    assert np.isfinite(r2["kf_Tcw_d"]).all()
    assert len(r2["erase_line"]) == len(q["line_obs_kf"])


def test_lba_oracle_empty_families():
    p = synth_lba.make_lba_problem(9, n_kf=3, n_fixed=1, n_points=50, n_pt_obs=140, n_lines=0, n_line_obs=0, n_plane_obs=(0, 0, 0))
    r = oracle_lib.local_bundle_adjustment(p)
    assert r["line_Xw"].shape == (0, 6) and all(len(a) == 0 for a in r["erase_plane"])
    assert np.isfinite(r["pt_Xw_d"]).all()
