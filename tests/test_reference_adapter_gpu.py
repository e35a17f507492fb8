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
from planarslam_b200 import synth_lba, synth_lines, synth_pose
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


def test_orbmatcher_search_by_bow_adapters():
    tot = 0
    for seed in range(3):
        kf, f = synth_lines.make_bow_pair(seed, n_kf=400, n_f=380, n_nodes=90)
        for ratio, ori in ((0.7, True), (0.9, False)):
            rn, rm = ref_lib.ref_search_by_bow(kf, f, ratio, ori)
            an, am = ref_lib.ref_search_by_bow(kf, f, ratio, ori, impl="adp")
            assert an == rn and np.array_equal(am, rm), (seed, ratio, ori)
            tot += an
        kf1, kf2 = synth_lines.make_bow_kf_pair(seed, n_kf=400, n_f=380, n_nodes=90)
        for ratio, ori in ((0.75, True), (0.9, False)):
            rn, rm = ref_lib.ref_search_by_bow_kf(kf1, kf2, ratio, ori)
            an, am = ref_lib.ref_search_by_bow_kf(kf1, kf2, ratio, ori, impl="adp")
            assert an == rn and np.array_equal(am, rm), (seed, ratio, ori)
            tot += an
    assert tot > 800


def test_lsdmatcher_search_by_projection_adapter():
    tot = 0
    for seed in range(6):
        frame, mp = synth_lines.make_line_search(seed, n_frame=40 + 4 * seed, n_map=120 + 30 * seed)
        for th, nnr in ((3.0, 0.8), (1.0, 0.8), (5.0, 0.6)):
            rn, ra = ref_lib.ref_line_search_by_projection(frame, mp, th, nnr)
            an, aa = ref_lib.ref_line_search_by_projection(frame, mp, th, nnr, impl="adp")
            assert an == rn and np.array_equal(aa, ra), (seed, th, nnr)
            tot += an
    assert tot > 50


def test_keyframe_database_adapter():
    from test_oracle_loopclose_ref import CASES
    found = 0
    for case in CASES:
        db = synth_lines.make_bow_database(**case)
        n_kf = len(db["off"]) - 1
        for min_score in (0.0, 0.03):
            rc, rw, rs = ref_lib.ref_detect_loop_candidates(db, min_score)
            ac, aw, as_ = ref_lib.ref_detect_loop_candidates(db, min_score, impl="adp")       # reads mnLoopWords / mLoopScore back from the KeyFrame objects
            assert np.array_equal(ac, rc) and np.array_equal(aw, rw) and np.array_equal(as_, rs), (case, min_score)
            found += len(rc)
        stale = np.random.default_rng(case["seed"]).uniform(0, 0.05, n_kf).astype(np.float32)
        rc, rw, rs = ref_lib.ref_detect_relocalization_candidates(db, stale)
        ac, aw, as_ = ref_lib.ref_detect_relocalization_candidates(db, stale, impl="adp")
        assert np.array_equal(ac, rc) and np.array_equal(aw, rw) and np.array_equal(as_, rs), case
    assert found > 5


def test_optimizer_local_bundle_adjustment_adapter():
    """Optimizer::LocalBundleAdjustment(KeyFrame*, bool*, Map*) of the reference and of the adapter on the same KeyFrame / MapPoint / MapLine / MapPlane graph: the
    slots each clears (erased observations and the bad-landmark cascade they trigger through the reference's own EraseObservation) must be identical, poses and
    landmark positions agree like the C ABI does with the reference (tests/test_cuda_vs_reference_functions_gpu.py)."""
    from test_oracle_lba_ref import SMALL
    small = dict(SMALL, line_kf_quirk=True)
    cases = [dict(seed=s, **small) for s in (1, 10)] + [dict(seed=2, n_kf=8, n_points=300, n_pt_obs=900, n_lines=0, n_line_obs=0, n_plane_obs=(0, 0, 0))]
    hard = [dict(seed=20, **small, line_norm3=False, outlier_frac=0.2, plane_outlier_frac=0.25)]
    n_cleared = 0
    for kw in cases + hard:
        p = synth_lba.restrict_to_local_planes(synth_lba.make_lba_problem(**kw))
        r = ref_lib.ref_full_local_bundle_adjustment(p)
        a = ref_lib.ref_full_local_bundle_adjustment(p, impl="adp")
        for k in ("pt_bad", "line_bad", "plane_bad", "erase_pt"):
            assert np.array_equal(a[k], r[k]), (kw, k)
        if len(p["line_obs_line"]):
            assert np.array_equal(a["erase_line"], r["erase_line"]), kw
        for t in range(3):
            if len(p["plane_obs_plane"][t]):
                assert np.array_equal(a["erase_plane"][t], r["erase_plane"][t]), (kw, t)
        tol = (2e-5, 5e-5) if kw in hard else (1e-6, 1e-6)
        for k in range(len(a["kf_Tcw_d"])):
            da, dt = synth_pose.pose_error(a["kf_Tcw_d"][k], r["kf_Tcw_d"][k])
            assert da < tol[0] and dt < tol[1], (kw, k, da, dt)
        assert np.median(np.abs(a["pt_Xw_d"] - r["pt_Xw_d"]).max(1)) < (2e-5 if kw in hard else 5e-6), kw
        n_cleared += int(r["erase_pt"].sum())
    assert n_cleared > 50


def test_lsdmatcher_search_by_descriptor_adapter():
    tot = 0
    for seed in range(6):
        rng = np.random.default_rng(100 + seed)
        n_kf, n_f = 40 - seed, 40 - 2 * seed
        kf_desc = rng.integers(0, 256, (n_kf, 32), dtype=np.uint8)
        f_desc = rng.integers(0, 256, (n_f, 32), dtype=np.uint8)
        src = rng.permutation(n_kf)[:n_f * 2 // 3]                          # two thirds of the frame lines re-observe a key-frame line (a few flipped bits)
        flips = np.packbits(rng.random((len(src), 256)) < 0.05, axis=1)
        f_desc[:len(src)] = kf_desc[src] ^ flips
        f_desc[-1] = f_desc[0]                                              # a duplicate: equal best and second-best distances for its source
        has_ml = (rng.random(n_kf) < 0.8).astype(np.uint8)
        rn, rm = ref_lib.ref_line_search_by_descriptor(kf_desc, has_ml, f_desc)
        an, am = ref_lib.ref_line_search_by_descriptor(kf_desc, has_ml, f_desc, impl="adp")
        assert an == rn and np.array_equal(am, rm), seed
        tot += rn
    assert tot > 60
