"""GPU: Frame::ComputePlanes' post-processing (pslam_compute_planes_batch: PEAC -> VoxelGrid -> MaxPointDistanceFromPlane -> RANSAC refit; integral-image surface
normals) vs the CPU oracle (oracle/planepost.cc, parity unpinned against PCL - see its header).  Bar: the same planes kept, voxel clouds bit-exact (order-free
fixed-point centroids), plane coefficients to 2e-5 (the closed-form eigen-solver runs float atan2 / cos / sin of two libms), surface normals bit-exact incl. the
NaN pattern.  The products then feed the kernels that consume them in the reference: PlaneMatcher::SearchMapByCoefficients and Tracking::TrackManhattanFrame."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth, synth_map

pytestmark = pytest.mark.gpu
K = synth.TUM3_K


def _ctx(n):
    from planarslam_b200._lib import Context
    return Context(640, 480, max_batch=n, fx=K[0], fy=K[1], cx=K[2], cy=K[3], depth_scale=float(np.float32(1.0 / 5000.0)))


def test_compute_planes_matches_oracle():
    from planarslam_b200.planes import ComputePlanes
    depth = np.stack([synth.render_frame(2, f)[1] for f in (0, 17, 40)] + [synth.piecewise_planar_depth(5, n_rect=11, curved=False)])
    res = ComputePlanes(_ctx(len(depth)), depth, 0.05)
    n_planes = 0
    for f in range(len(depth)):
        o = oracle_lib.planes_post(depth[f])
        r = res[f]
        assert [p["src"] for p in o] == r["src"].tolist(), f
        for k, p in enumerate(o):
            assert np.array_equal(r["points"][k], p["points"]), (f, k)
            assert np.abs(r["coef"][k] - p["coef"]).max() < 2e-5, (f, k, r["coef"][k], p["coef"])
        n_planes += len(o)
        osn = oracle_lib.surface_normals(depth[f])
        assert r["normals"].shape == osn.shape
        assert np.array_equal(np.isnan(r["normals"]), np.isnan(osn)), f
        assert np.array_equal(np.nan_to_num(r["normals"]), np.nan_to_num(osn)), f
    assert n_planes >= 8


def test_planes_and_normals_feed_the_matcher_and_manhattan():
    from planarslam_b200.manhattan import TrackManhattanFrame
    from planarslam_b200.matcher import PlaneMatcher
    from planarslam_b200.planes import ComputePlanes
    depth = np.stack([synth.render_frame(2, f, hole_frac=0.0)[1] for f in (10, 12)])
    ctx = _ctx(2)
    a, b = ComputePlanes(ctx, depth, 0.05)
    assert len(a["coef"]) >= 2 and len(b["coef"]) >= 2
    # frame 10's planes as the "map" (world = camera 10), frame 12 observed from its true relative pose: every plane must associate
    T10, T12 = synth_map.true_pose(10), synth_map.true_pose(12)
    Trel = (T12 @ np.linalg.inv(T10)).astype(np.float32)
    off = np.concatenate([[0], np.cumsum([len(p) for p in a["points"]])]).astype(np.int32)
    pts = np.concatenate(a["points"]).astype(np.float32)
    n, m, v, p = PlaneMatcher(0.05, 0.985, 0.08716, 0.9962, ctx=ctx).SearchMapByCoefficients(Trel, b["coef"], a["coef"], np.zeros(len(a["coef"]), np.uint8), off, pts)
    assert n == len(b["coef"]) and (m >= 0).all(), (n, m)
    # the surface normals of a Manhattan-like scene (floor / two walls) give TrackManhattanFrame three directions
    sn = a["normals"]
    ok = np.isfinite(sn[:, 0])
    R0 = np.eye(3, dtype=np.float32)
    res, _, _ = TrackManhattanFrame(ctx, R0[None], [np.ascontiguousarray(sn[ok, :3])], [np.zeros((0, 3))])
    assert res[0]["n_cone"].sum() > 0
