"""Golden vectors from the REFERENCE's own code compiled in the build container (oracle/_ref, `make -C oracle ref`: the reference
sources under /root/reference, unmodified, against the stand-in headers of oracle/ref/shims/).  /root/reference does not exist on
the GPU box and may not exist in a later container, so the outputs are committed as small fixtures:
  tests/golden/peac_reference.npz   PlaneDetection::readDepthImage + runPlaneDetection on seeded synthetic depth frames: per frame the
                                    label image (membershipImg, int16-packed), plane parameters, supports and plane_vertices_ digests
  tests/golden/orb_reference.npz    Planar_SLAM::ORBextractor::operator() (monotonic allocator, see oracle/ref/orb_driver.cc) on seeded synthetic
                                    frames: the key-point records and descriptors of two frames in full, SHA-1 digests of more
  tests/golden/bow_reference.npz    ORBVocabulary::transform (DBoW2 compiled from the reference; vocabulary read by its loadFromTextFile) on synthetic
                                    vocabularies / features: BowVector and FeatureVector in full for each case
  tests/golden/line3d_reference.npz the 3-D line fit of src/LineExtractor.cpp (+ libc rand) on clean and corrupted depth: every output field
  tests/golden/pose_reference.npz   PoseOptimization by the reference's g2o (libpose_ref.so): optimised pose, inlier count and outlier flags
  tests/golden/manhattan_reference.npz  Tracking::TrackManhattanFrame (libtrack_ref.so): the returned rotation
  tests/golden/match_reference.npz  the matchers (libmatch_ref.so): SearchByProjection x2, SearchByBoW, LSDmatcher::SearchByProjection, PlaneMatcher
  tests/golden/lba_reference.npz    LocalBundleAdjustment by the reference's g2o (libpose_ref.so): key-frame poses, points, lines, planes, erase flags
Run: python tools/make_golden_ref.py"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_lib  # noqa: E402
from planarslam_b200 import synth  # noqa: E402

PEAC_SCENES = [("room", s) for s in (0, 3)] + [("patches", s) for s in (0, 2, 5, 11)]


def peac_scene(kind, seed):
    if kind == "room":
        return synth.render_frame(seed=seed, frame=3 * seed)[1]
    return synth.piecewise_planar_depth(seed, n_rect=6 + seed, curved=(seed % 2 == 0))


def digest(a):
    return np.frombuffer(hashlib.sha1(np.ascontiguousarray(a).tobytes()).digest(), np.uint8).copy()


ORB_CASES = [("tum%d" % s, dict(seed=s), {}) for s in (0, 7)] + [("nf500", dict(seed=1), dict(nfeatures=500)), ("nf2000", dict(seed=1), dict(nfeatures=2000)),
                                                                    ("l4s15", dict(seed=1), dict(nlevels=4, scale=1.5)), ("th40", dict(seed=1), dict(ini_th=40, min_th=10)),
                                                                    ("big", dict(seed=2, width=1280, height=960), dict(nfeatures=2000)),
                                                                    ("small", dict(seed=3, width=320, height=240), {})]


def orb_image(kw):
    return synth.render_frame(frame=3 * kw["seed"], **kw)[0]


BOW_CASES = [(0, 10, 3, 4), (1, 10, 4, 4), (2, 6, 5, 2), (3, 9, 3, 0)]          # (seed, k, L, levelsup)


def bow_case(seed, k, L):
    from planarslam_b200 import synth_lines as sl
    voc = sl.make_vocabulary(seed, k=k, L=L)
    return voc, sl.make_features_for_vocabulary(seed + 10, voc, 1000)


if __name__ == "__main__":
    out = {}
    for kind, seed in PEAC_SCENES:
        labels, planes, members = ref_lib.ref_peac_run(peac_scene(kind, seed))
        key = f"{kind}{seed}"
        assert labels.min() >= -32768 and labels.max() <= 32767
        out[key + "_labels"] = labels.astype(np.int16)
        out[key + "_planes"] = np.stack([p[0] for p in planes])
        out[key + "_N"] = np.array([p[1] for p in planes], np.int32)
        out[key + "_members_sha1"] = np.stack([digest(m) for m in members])
        print(key, len(planes), "planes")
    path = os.path.join(ROOT, "tests", "golden", "peac_reference.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
    out = {}
    for name, ikw, okw in ORB_CASES:
        k, d = ref_lib.ref_orb_extract(orb_image(ikw), **okw)
        out[name + "_n"] = np.array([len(k)], np.int32)
        out[name + "_kps_sha1"], out[name + "_desc_sha1"] = digest(k), digest(d)
        if name.startswith("tum"):
            out[name + "_kps"], out[name + "_desc"] = k.view(np.uint8).reshape(len(k), 28), d
        print(name, len(k), "key points")
    import tempfile
    bow = {}
    for seed, k, L, lu in BOW_CASES:
        voc, feats = bow_case(seed, k, L)
        with tempfile.TemporaryDirectory() as td:
            ref_lib.write_vocabulary_txt(voc, os.path.join(td, "voc.txt"))
            r = ref_lib.RefVocabulary(os.path.join(td, "voc.txt")).transform(feats, lu)
        for key, v in r.items():
            bow[f"s{seed}_{key}"] = v
        print("bow", seed, len(r["word_id"]), "words", len(r["node_id"]), "nodes")
    bpath = os.path.join(ROOT, "tests", "golden", "bow_reference.npz")
    np.savez_compressed(bpath, **bow)
    print(bpath, os.path.getsize(bpath), "bytes")
    from test_oracle_line3d_ref import CASES as L3_CASES, KEYS as L3_KEYS, case_inputs
    l3 = {}
    for s_, frac, sigma, seed in L3_CASES:
        kl, depth = case_inputs(s_, frac, sigma)
        r = ref_lib.ref_lines3d_frame(kl, depth, synth.TUM3_K, seed=seed)
        for k in L3_KEYS:
            l3[f"c{s_}_{k}"] = r[k]
        print("line3d", s_, int(r["valid"].sum()), "valid lines")
    lpath = os.path.join(ROOT, "tests", "golden", "line3d_reference.npz")
    np.savez_compressed(lpath, **l3)
    print(lpath, os.path.getsize(lpath), "bytes")
    from test_oracle_pose_ref import FLAGS as POSE_FLAGS, GOLD_CASES as POSE_CASES
    from planarslam_b200 import synth_pose
    pg = {}
    for i, kw in enumerate(POSE_CASES):
        r = ref_lib.ref_pose_optimization(synth_pose.make_pose_problem(**kw))
        pg[f"c{i}_Tcw_d"], pg[f"c{i}_n"] = r["Tcw_d"], np.array([r["n_inliers"]], np.int32)
        for k in POSE_FLAGS:
            pg[f"c{i}_{k}"] = r[k]
        print("pose", i, r["n_inliers"], "inliers")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "pose_reference.npz"), **pg)

    from test_oracle_lba_ref import GOLD_CASES as LBA_CASES
    from planarslam_b200 import synth_lba
    lg = {}
    for i, kw in enumerate(LBA_CASES):
        r = ref_lib.ref_local_bundle_adjustment(synth_lba.make_lba_problem(**kw))
        for k in ("kf_Tcw_d", "pt_Xw_d", "line_Xw_d", "plane_Xw_d", "erase_pt", "erase_line"):
            lg[f"c{i}_{k}"] = r[k]
        for t in range(3):
            lg[f"c{i}_erase_plane{t}"] = r["erase_plane"][t]
        print("lba", i, r["iterations"], int(r["erase_pt"].sum()), "point observations erased")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "lba_reference.npz"), **lg)

    from test_oracle_match_ref import gold_cases
    mg = {}
    for name, _, ref_call in gold_cases():
        r = ref_call()
        mg[f"{name}_n"] = np.array([r[0]], np.int32)
        for k, arr in enumerate(r[1:]):
            mg[f"{name}_{k}"] = arr
        print("match", name, int(r[0]), "matches")
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "match_reference.npz"), **mg)

    from test_oracle_manhattan_ref import CASES as MH_CASES
    from planarslam_b200.synth_manhattan import make_manhattan
    hg = {}
    for i, kw in enumerate(MH_CASES):
        R_last, normals, dirs, _ = make_manhattan(**kw)
        hg[f"R{i}"] = ref_lib.ref_track_manhattan_frame(R_last, normals, dirs)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "manhattan_reference.npz"), **hg)
    print("manhattan", len(MH_CASES), "cases")
    path = os.path.join(ROOT, "tests", "golden", "orb_reference.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes")
