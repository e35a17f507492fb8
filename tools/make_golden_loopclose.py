#!/usr/bin/env python3
"""Golden vectors of the loop-closure / relocalisation candidate search, produced by THE REFERENCE'S OWN CODE compiled here (oracle/_ref/libmatch_ref.so:
src/KeyFrameDatabase.cc, Thirdparty/DBoW2 scoring, src/ORBmatcher.cc called unmodified through oracle/ref/match_driver.cc) on the seeded inputs of
planarslam_b200.synth_lines.  Run in the build container (needs /root/reference to have built the library):
    python tools/make_golden_loopclose.py        -> tests/golden/loopclose_reference.npz
tests/test_oracle_loopclose_ref.py::test_oracle_matches_loopclose_golden checks the oracle against the file wherever the library is absent (the GPU box)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_lib                                               # noqa: E402
from planarslam_b200 import synth_lines                      # noqa: E402
from test_oracle_loopclose_ref import GOLD_CASES, GOLD_MIN_SCORES, GOLD_KF_SEEDS   # noqa: E402

out = {}
for i, case in enumerate(GOLD_CASES):
    db = synth_lines.make_bow_database(**case)
    for j, ms in enumerate(GOLD_MIN_SCORES):
        c, w, s = ref_lib.ref_detect_loop_candidates(db, ms)
        out[f"loop{i}_{j}_cand"], out[f"loop{i}_{j}_words"], out[f"loop{i}_{j}_score"] = c, w, s
    stale = np.random.default_rng(case["seed"]).uniform(0, 0.05, len(db["off"]) - 1).astype(np.float32)
    c, w, s = ref_lib.ref_detect_relocalization_candidates(db, stale)
    out[f"reloc{i}_cand"], out[f"reloc{i}_words"], out[f"reloc{i}_score"] = c, w, s
    print("database", case, "loop candidates", [len(out[f"loop{i}_{j}_cand"]) for j in range(len(GOLD_MIN_SCORES))], "reloc", len(c))
for seed in GOLD_KF_SEEDS:
    kf1, kf2 = synth_lines.make_bow_kf_pair(seed, n_kf=400, n_f=380, n_nodes=90)
    n, m = ref_lib.ref_search_by_bow_kf(kf1, kf2, 0.75, True)
    out[f"bowkf{seed}_n"], out[f"bowkf{seed}_match"] = np.array([n], np.int32), m
    print("SearchByBoW(KF, KF) seed", seed, n, "matches")
path = os.path.join(ROOT, "tests", "golden", "loopclose_reference.npz")
np.savez_compressed(path, **out)
print(path, os.path.getsize(path), "bytes")
