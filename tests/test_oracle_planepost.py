"""CPU: the plane post-processing oracle (oracle/planepost.cc: VoxelGrid + distance check + RANSAC refit of Frame::ComputePlanes / MaxPointDistanceFromPlane,
and the integral-image surface normals).  PARITY UNPINNED against PCL (not in this image); checked here against independent numpy statements of the same
geometry: voxel centroids, a least-squares plane, normals of the rendered planes, and the mt19937 stream against the C++ standard library's."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth

K = synth.TUM3_K


def test_voxel_grid_and_refit_against_numpy():
    _, d16, z, (R, t) = synth.render_frame(seed=2, frame=7)
    planes = oracle_lib.planes_post(d16)
    po = oracle_lib.PeacOracle(d16)
    assert 2 <= len(planes) <= len(po.planes)
    for P in planes:
        idx = po.membership[P["src"]]
        v, u = np.divmod(idx, 640)
        zz = d16.ravel()[idx].astype(np.float64) * float(np.float32(1.0 / 5000.0))
        X = np.stack([(u - K[2]) * zz / K[0], (v - K[3]) * zz / K[1], zz], 1).astype(np.float32)
        key = np.floor(X * np.float32(10.0)).astype(np.int64)
        key -= key.min(0)
        dims = key.max(0) + 1
        lin = key[:, 0] + key[:, 1] * dims[0] + key[:, 2] * dims[0] * dims[1]
        order = np.argsort(lin, kind="stable")
        uniq, start = np.unique(lin[order], return_index=True)
        assert len(uniq) == len(P["points"])
        cent = np.stack([X[order][a:b].astype(np.float64).mean(0) for a, b in zip(start, list(start[1:]) + [len(order)])])
        assert np.abs(cent - P["points"]).max() < 2e-5                           # float vs double accumulation
        # the refit is the least-squares plane of the (all-inlier) centroids, signed like the PEAC plane
        c = P["points"].astype(np.float64)
        n = np.linalg.svd(c - c.mean(0))[2][2]
        d = -n @ c.mean(0)
        coef = P["coef"].astype(np.float64)
        s = np.sign(coef[:3] @ n)
        # pcl::eigen33 is a closed-form cubic solver run in FLOAT: for a nearly flat point set (smallest eigenvalue ~1e-5 of the largest) its root carries an
        # error of ~1e-4 that the cross-product eigenvector turns into ~5e-3 of normal direction - PCL's own behaviour, kept as it is
        assert np.abs(coef[:3] - s * n).max() < 2e-2 and abs(coef[3] - s * d) < 5e-2
        assert 0.97 * len(c) <= P["n_inliers"] <= len(c) and 1 <= P["n_iterations"] <= 51
        assert np.percentile(np.abs(c @ coef[:3] + coef[3]), 97) < 0.05


def test_surface_normals_follow_the_rendered_planes():
    _, d16, z, (R, t) = synth.render_frame(seed=3, frame=20, hole_frac=0.0)
    sn = oracle_lib.surface_normals(d16)
    assert sn.shape == (80 * 106, 8) or sn.shape[0] == (160 // 2) * (214 // 2)
    ok = np.isfinite(sn[:, 0])
    assert 0.5 < ok.mean() < 0.9                                                 # the 10-pixel border and the depth edges are NaN
    nrm = sn[ok, :3]
    assert np.abs(np.linalg.norm(nrm, axis=1) - 1).max() < 1e-5
    assert ((sn[ok, 3:6] * nrm).sum(1) < 0).all()                                # flipped towards the view point
    # the scene has three plane normals (floor, left wall, back wall); in camera coordinates n_c = R_wc^T n_w
    world = np.array([[0, 1, 0], [1, 0, 0], [0, 0, 1]], np.float64)
    cam = world @ R                                                              # rows: R^T n
    cosmax = np.abs(nrm @ cam.T).max(1)
    assert np.mean(cosmax > 0.995) > 0.9, np.mean(cosmax > 0.995)
    assert (sn[:, 6] % 6 == 3).all() and (sn[:, 7] % 6 == 3).all()              # odd columns / rows of the 3x sub-sampled grid


def test_pcl_rng_is_mt19937_seed_12345():
    import ctypes as C
    L = oracle_lib.lib()
    L.orc_pcl_rng.restype = C.c_uint32
    assert L.orc_pcl_rng(1) == 3992670690 and L.orc_pcl_rng(1000) == 47030557  # std::mt19937(12345u): 1st and 1000th output (g++ 13 libstdc++)
