"""CPU: sanity of the projection-search oracle on the synthetic sequence (map built from frame 10, searched from frame 11)."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth, synth_map


def scenario(f0=10, f1=11, pose_noise=0.002, seed=0):
    rng = np.random.default_rng(seed)
    g0, d0, _, _ = synth.render_frame(2, f0)
    g1, d1, _, _ = synth.render_frame(2, f1)
    k0, de0 = oracle_lib.orb_extract(g0)
    k1, de1 = oracle_lib.orb_extract(g1)
    fa0, fa1 = synth_map.frame_arrays(k0, de0, d0), synth_map.frame_arrays(k1, de1, d1)
    T0, T1 = synth_map.true_pose(f0), synth_map.true_pose(f1)
    m = synth_map.map_from_frame(fa0, T0)
    T1n = T1.copy()
    T1n[:3, 3] += rng.normal(0, pose_noise, 3)
    fv = synth_map.frame_view(fa1, T1n)
    lf = dict(n=fa0["n"], keys=fa0["keys_un"], map_point=np.full(fa0["n"], -1, np.int32), outlier=np.zeros(fa0["n"], np.uint8),
              Tcw=np.ascontiguousarray(T0, np.float32))
    lf["map_point"][m["src_index"]] = np.arange(m["n"], dtype=np.int32)
    return fv, m, lf


def test_search_by_projection_map_finds_consistent_matches():
    fv, m, lf = scenario()
    n, matches, in_view = oracle_lib.search_by_projection_map(fv, m, 3.0, 0.8, np.full(fv["n"], -1, np.int32))
    assert in_view.sum() > 0.7 * m["n"]
    assert n == (matches >= 0).sum() and n > 150
    # a matched keypoint lies where its map point projects (within the search radius)
    idx = np.nonzero(matches >= 0)[0]
    T = np.asarray(fv["Tcw"], np.float64)
    Xc = m["pos"][matches[idx]].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    u = fv["fx"] * Xc[:, 0] / Xc[:, 2] + fv["cx"]
    v = fv["fy"] * Xc[:, 1] / Xc[:, 2] + fv["cy"]
    err = np.hypot(u - fv["keys_un"]["x"][idx], v - fv["keys_un"]["y"][idx])
    assert np.median(err) < 3.0 and err.max() < 60.0
    # already-matched keypoints holding a point with observations are never stolen
    pre = np.full(fv["n"], -1, np.int32)
    pre[idx[:20]] = 0
    n2, matches2, _ = oracle_lib.search_by_projection_map(fv, m, 3.0, 0.8, pre)
    assert (matches2[idx[:20]] == 0).all()


def test_search_by_projection_last_frame():
    fv, m, lf = scenario()
    n, matches = oracle_lib.search_by_projection_last(fv, lf, m, 15.0, False, True, np.full(fv["n"], -1, np.int32))
    assert n == (matches >= 0).sum() and n > 150
    n_no_ori, _ = oracle_lib.search_by_projection_last(fv, lf, m, 15.0, False, False, np.full(fv["n"], -1, np.int32))
    assert n_no_ori >= n
