"""CPU: the oracle of Tracking::TrackManhattanFrame (oracle/manhattan.cc, src/Tracking.cc:763-1157) on synthetic Manhattan worlds:
recovers the true rotation, agrees with an independent vectorised numpy statement of the algorithm, and reproduces the reference's
corner cases (fewer than two directions -> partially updated matrix returned without the SVD; exactly two -> cross product)."""
import numpy as np

import oracle_lib
from planarslam_b200.synth_manhattan import make_manhattan


def _numpy_track(R_last, normals, dirs):
    """Independent float64 statement (no float32 staging, vectorised)."""
    R = R_last.astype(np.float64).copy()
    n = len(normals)
    N, D = normals.astype(np.float64), dirs.astype(np.float64)
    def T(a):
        c = [(a + 3) % 3, (a + 4) % 3, (a + 5) % 3]
        return R[:, c].T
    sets = []
    for a in (1, 2, 3):
        qn, qd = N @ T(a).T, D @ T(a).T
        sets.append((np.nonzero(np.hypot(qn[:, 0], qn[:, 1]) < np.sin(0.2018))[0], np.nonzero(np.hypot(qd[:, 0], qd[:, 1]) < np.sin(0.1018))[0]))
    cnt = sorted(len(s[0]) for s in sets)
    min_num = n // 20
    if cnt[1] < min_num:
        min_num = (cnt[0] + cnt[1]) // 2
    found = [0, 0, 0]
    for a in (1, 2, 3):
        Ta = T(a)
        q = np.concatenate([N[sets[a - 1][0]] @ Ta.T, D[sets[a - 1][1]] @ Ta.T])
        lam = np.hypot(q[:, 0], q[:, 1])
        q, lam = q[lam < np.sin(0.2518)], lam[lam < np.sin(0.2518)]
        with np.errstate(invalid="ignore", divide="ignore"):
            s = np.arcsin(lam) / (lam / np.abs(q[:, 2]))
            mj = np.stack([s * q[:, 0] / q[:, 2], s * q[:, 1] / q[:, 2]], 1)
        mj = mj[~np.isnan(mj).any(1)]
        if len(mj) > min_num:
            k = np.exp(-20 * (mj ** 2).sum(1))
            sj = (k[:, None] * mj).sum(0) / k.sum()
            al = np.linalg.norm(sj)
            v = Ta.T @ np.array([np.tan(al) / al * sj[0], np.tan(al) / al * sj[1], 1.0])
            R[:, a - 1] = v / np.linalg.norm(v)
            found[a - 1] = 1
    if sum(found) < 2:
        return R, found
    if sum(found) == 2:
        if found[0] and found[1]:
            t, x = 2, np.cross(R[:, 0], R[:, 1])
        elif found[1] and found[2]:
            t, x = 0, np.cross(R[:, 2], R[:, 1])
        else:
            t, x = 1, np.cross(R[:, 0], R[:, 2])
        R[:, t] = x
        if abs(np.linalg.det(R) + 1) < 0.5:
            R[:, t] = -x
    U, _, Vt = np.linalg.svd(R)
    return U @ Vt, found


def _angle(Ra, Rb):
    return np.degrees(np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1)))


def test_manhattan_recovers_rotation_and_matches_numpy():
    for seed in range(8):
        R_last, normals, dirs, R_true = make_manhattan(seed)
        r = oracle_lib.track_manhattan_frame(R_last, normals, dirs)
        assert r["found"].sum() == 3 and r["svd_applied"] == 1
        R = r["R"].astype(np.float64)
        assert np.allclose(R.T @ R, np.eye(3), atol=1e-5) and np.linalg.det(R) > 0.99
        # One call roughly halves the error, it does not remove it: R_cm aliases R_cm_update in the reference (src/Tracking.cc:970), so
        # the tangent basis of the later axes already contains the corrected earlier columns and their own correction shrinks;
        # the final SVD averages the corrected and uncorrected columns.
        assert _angle(R, R_true) < 0.6 * _angle(R_last.astype(np.float64), R_true)
        Rn, fn = _numpy_track(R_last, normals, dirs)
        assert fn == list(r["found"]) and np.allclose(R, Rn, atol=2e-5)
        assert (r["n_selected"] >= r["n_cone"]).all() and (r["n_selected"] <= r["n_cone"] + len(dirs)).all()
        assert (r["density"] > 0).all() and (r["density"] <= 1).all()
        for a in range(3):
            assert int(((r["normal_mask"] >> a) & 1).sum()) <= r["n_cone"][a]


def test_manhattan_two_and_one_direction_cases():
    # only two populated axes: the third comes from the cross product
    R_last, normals, dirs, R_true = make_manhattan(3, weights=(0.5, 0.5, 0.0), clutter=0.02, n_lines=0)
    r = oracle_lib.track_manhattan_frame(R_last, normals, dirs)
    assert list(r["found"]) == [1, 1, 0] and r["svd_applied"] == 1
    R = r["R"].astype(np.float64)
    assert np.linalg.det(R) > 0.99 and _angle(R, R_true) < 0.6 * _angle(R_last.astype(np.float64), R_true)
    Rn, fn = _numpy_track(R_last, normals, dirs)
    assert fn == [1, 1, 0] and np.allclose(R, Rn, atol=2e-5)
    # a single populated axis: the reference returns the partially updated matrix (column 0 replaced, no SVD)
    R_last, normals, dirs, R_true = make_manhattan(4, weights=(1.0, 0.0, 0.0), clutter=0.0, n_lines=0)
    r = oracle_lib.track_manhattan_frame(R_last, normals, dirs)
    assert list(r["found"]) == [1, 0, 0] and r["svd_applied"] == 0
    assert np.array_equal(r["R"][:, 1:], R_last[:, 1:]) and not np.array_equal(r["R"][:, 0], R_last[:, 0])
    assert np.degrees(np.arccos(np.clip(abs(r["R"][:, 0].astype(np.float64) @ R_true[:, 0]), 0, 1))) < 0.5
    # no normals at all
    r = oracle_lib.track_manhattan_frame(R_last, np.zeros((0, 3), np.float32), np.zeros((0, 3)))
    assert r["found"].sum() == 0 and np.array_equal(r["R"], R_last)
