"""GPU: the reference-TYPED boundary (include/pslam_reference_adapter.hpp: pslam_adapter::ref::Optimizer / ORBmatcher / PlaneMatcher with the reference's own
signatures - Frame*, Frame&, std::vector<MapPoint*>, std::vector<MapPlane*>) against the reference's functions ON THE SAME OBJECTS.

oracle/ref/match_driver.cc builds Frame / MapPoint / MapPlane / MapLine objects (the reference's classes, compiled unmodified) from plain arrays; it is compiled
twice: libmatch_ref.so calls the reference's ORBmatcher::SearchByProjection x2, PlaneMatcher::SearchMapByCoefficients, Optimizer::PoseOptimization /
TranslationOptimization on them, libadapter_ref.so (oracle/ref/adapter_driver.cc) calls the adapter classes of the same names - which gather from the objects
under the reference's mutexes, run the CUDA path through the C ABI and write back mvpMapPoints / mvpMapPlanes / mvb*Outlier / mTcw.  The read-back is shared.
Bar: identical match lists and outlier flags; poses as in tests/test_cuda_vs_reference_functions_gpu.py."""
import numpy as np
import pytest

import ref_lib
from planarslam_b200 import synth_pose
from test_oracle_match_ref import PLANE_TH, last_case, map_case
from test_oracle_planematch import _scenario as plane_scenario

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(ref_lib.adapter_lib() is None, reason="oracle/_ref/libadapter_ref.so not present")]
FLAGS = ("outlier_pt", "outlier_line", "outlier_plane", "outlier_par", "outlier_ver")


def test_orbmatcher_search_by_projection_adapters():
    tot = 0
    for seed, th, nnr in ((0, 3.0, 0.8), (1, 1.0, 0.8), (2, 5.0, 0.9)):
        fv, m, th, nnr, pre = map_case(seed, th, nnr)
        rn, rmatches, rview = ref_lib.ref_search_by_projection_map(fv, m, th, nnr, pre)
        an, amatches, aview = ref_lib.ref_search_by_projection_map(fv, m, th, nnr, pre, impl="adp")
        assert an == rn and np.array_equal(amatches, rmatches) and np.array_equal(aview, rview), seed
        tot += an
    for seed, th, mono, ori in ((0, 15.0, False, True), (2, 15.0, True, False), (3, 30.0, False, True)):
        a = last_case(seed, th, mono, ori)
        rn, rmatches = ref_lib.ref_search_by_projection_last(*a)
        an, amatches = ref_lib.ref_search_by_projection_last(*a, impl="adp")
        assert an == rn and np.array_equal(amatches, rmatches), seed
        tot += an
    assert tot > 1500


def test_plane_matcher_adapter():
    rng = np.random.default_rng(3)
    tot = 0
    for trial in range(8):
        T, fc, mc, bad, off, pts = plane_scenario(trial, rng)
        r = ref_lib.ref_plane_match(T, fc, mc, bad, off, pts, *PLANE_TH)
        a = ref_lib.ref_plane_match(T, fc, mc, bad, off, pts, *PLANE_TH, impl="adp")
        assert a[0] == r[0] and all(np.array_equal(x, y) for x, y in zip(a[1:], r[1:])), trial
        tot += r[0]
    assert tot >= 8


def test_optimizer_pose_and_translation_adapters():
    cases = [dict(seed=s, frame=3 * s) for s in range(4)] + [dict(seed=21, frame=1, outlier_frac=0.25, rot_pert=0.05, trans_pert=0.08),
                                                             dict(seed=30, frame=1, n_points=0, n_lines=0), dict(seed=7, frame=21, n_planes=0, n_par=0, n_ver=0),
                                                             dict(seed=32, frame=3, n_points=2, n_lines=0, n_planes=0, n_par=0, n_ver=0, outlier_frac=0.0)]
    for translation_only in (False, True):
        for kw in cases:
            p = synth_pose.make_pose_problem(**kw)
            r = ref_lib.ref_full_pose_optimization(p, translation_only)
            a = ref_lib.ref_full_pose_optimization(p, translation_only, impl="adp")
            assert a["n_inliers"] == r["n_inliers"], (translation_only, kw)
            for k in FLAGS:
                assert np.array_equal(a[k], r[k]), (translation_only, kw, k)
            da, dt = synth_pose.pose_error(a["Tcw"], r["Tcw"])
            assert da < 5e-6 and dt < 2e-5, (translation_only, kw, da, dt)
