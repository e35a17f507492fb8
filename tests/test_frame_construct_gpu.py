"""GPU: pslam_frame_construct_batch (the compute of the RGB-D Frame constructor, src/Frame.cc:55-140, one upload) returns exactly what the per-function entry points
return on the same frames - which are themselves held to the oracle / the compiled reference by their own tests."""
import numpy as np
import pytest

from planarslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_frame_construct_equals_the_separate_calls():
    from planarslam_b200._lib import Context
    from planarslam_b200.frame import ComputeStereoFromRGBD, ConstructFrames, FrameOutputs
    from planarslam_b200.lines import LineSegment, isLineGood
    from planarslam_b200.planes import ComputePlanes
    B = 5
    fr = [synth.render_frame(seed=2, frame=7 * k) for k in range(B)]
    gray, depth = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])
    ctx = Context(640, 480, B)
    out = ConstructFrames(ctx, gray, depth, FrameOutputs(ctx, B, cap_plane_pts=16384), bf=40.0, plane_dist_th=0.05, line_seed=1)
    # a second context runs the per-function host-pointer calls
    ref = Context(640, 480, B)
    L = ref.L
    cap = int(L.pslam_orb_max_keypoints(ref.h))
    from planarslam_b200._lib import KEYPOINT_DTYPE
    kps, desc, n = np.zeros((B, cap), KEYPOINT_DTYPE), np.zeros((B, cap, 32), np.uint8), np.zeros(B, np.int32)
    ref.check(L.pslam_orb_extract_batch(ref.h, gray.ctypes.data, B, kps.ctypes.data, desc.ctypes.data, cap, n.ctypes.data))
    assert np.array_equal(out.n_keys, n) and n.min() > 500
    for f in range(B):
        assert np.array_equal(out.keys[f, :n[f]], kps[f, :n[f]]) and np.array_equal(out.desc[f, :n[f]], desc[f, :n[f]])
    ur, dz = ComputeStereoFromRGBD(ref, kps, n, depth, float(ref.cfg.depth_scale), 40.0)
    for f in range(B):
        assert np.array_equal(out.u_right[f, :n[f]], ur[f, :n[f]]) and np.array_equal(out.depth_kp[f, :n[f]], dz[f, :n[f]])
    ls = LineSegment(ctx=ref)
    res = ls.ExtractLineSegmentWithDescriptors(gray, 40)
    K = (ref.cfg.fx, ref.cfg.fy, ref.cfg.cx, ref.cfg.cy)
    for f in range(B):
        kl, lf, ld = res[f][0], res[f][1], res[f][2]
        m = len(kl)
        assert out.n_lines[f] == m > 10
        assert out.keylines[f, :m].tobytes() == kl.tobytes() and np.array_equal(out.line_functions[f, :m], lf) and np.array_equal(out.line_desc[f, :m], ld)
        pad = np.zeros(40, kl.dtype)
        pad[:m] = kl
        l3, drawn = isLineGood(ref, pad[None], [m], depth[f:f + 1], K, float(ref.cfg.depth_scale), seed=1)
        assert out.lines3d[f, :m].tobytes() == l3[0, :m].tobytes() and out.n_rand_drawn[f] == drawn[0]
    planes = ComputePlanes(ref, depth, 0.05, normals=True)
    for f in range(B):
        p = planes[f]
        k = len(p["src"])
        assert out.n_planes[f] == k >= 1
        assert np.array_equal(out.plane_src[f, :k], p["src"]) and np.array_equal(out.plane_coef[f, :k], p["coef"])
        for q in range(k):
            assert np.array_equal(out.plane_pts[f, out.plane_pt_off[f, q]:out.plane_pt_off[f, q + 1]], p["points"][q])
        assert np.array_equal(out.surface_normals8[f], p["normals"], equal_nan=True)
