"""CPU: planarslam_b200/csrc/bowdb_select.h - the host-side list logic of pslam_detect_loop_candidates / pslam_detect_relocalization_candidates - compiled
with g++ and fed with per-key-frame triples from a scalar loop in the kernel's arithmetic, compared with the oracle (oracle/loopclose.cc, itself identical to
the compiled src/KeyFrameDatabase.cc).  The warp kernel that produces the triples on the device is covered by tests/test_loopclose_gpu.py."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth_lines
from test_oracle_loopclose_ref import CASES

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    out = tmp_path_factory.mktemp("bowdb") / "libbowdb_host.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-I", os.path.join(ROOT, "planarslam_b200", "csrc"),
                    "-o", str(out), os.path.join(ROOT, "tests", "host_harness", "bowdb_select_host.cc")], check=True)
    L = C.CDLL(str(out))
    L.host_detect_loop_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_float] + [C.c_void_p] * 3
    L.host_detect_relocalization_candidates.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 4 + [C.c_int] + [C.c_void_p] * 3
    return L


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"seed{c['seed']}")
def test_selection_logic_matches_oracle(host_lib, case):
    db = synth_lines.make_bow_database(**case)
    d, n_kf, covis, stride = oracle_lib._db_args(db)
    for min_score in (0.0, 0.01, 0.03, 0.08):
        cand, words, score = np.zeros(max(n_kf, 1), np.int32), np.zeros(max(n_kf, 1), np.int32), np.full(max(n_kf, 1), -1.0, np.float32)
        n = host_lib.host_detect_loop_candidates(d["q_word"].ctypes.data, d["q_val"].ctypes.data, len(d["q_word"]), n_kf, d["off"].ctypes.data, d["word"].ctypes.data,
                                                 d["val"].ctypes.data, covis, stride, d["connected"].ctypes.data, min_score, cand.ctypes.data, words.ctypes.data,
                                                 score.ctypes.data)
        oc, ow, os_ = oracle_lib.detect_loop_candidates(db, min_score)
        assert n == len(oc) and np.array_equal(cand[:n], oc) and np.array_equal(words[:n_kf], ow) and np.array_equal(score[:n_kf], os_)
    rng = np.random.default_rng(case["seed"])
    for stale in (np.zeros(n_kf, np.float32), rng.uniform(0, 0.05, n_kf).astype(np.float32)):
        cand, words, score = np.zeros(max(n_kf, 1), np.int32), np.zeros(max(n_kf, 1), np.int32), stale.copy()
        n = host_lib.host_detect_relocalization_candidates(d["q_word"].ctypes.data, d["q_val"].ctypes.data, len(d["q_word"]), n_kf, d["off"].ctypes.data,
                                                           d["word"].ctypes.data, d["val"].ctypes.data, covis, stride, score.ctypes.data, cand.ctypes.data,
                                                           words.ctypes.data)
        oc, ow, os_ = oracle_lib.detect_relocalization_candidates(db, stale)
        assert n == len(oc) and np.array_equal(cand[:n], oc) and np.array_equal(words[:n_kf], ow) and np.array_equal(score, os_)


def test_selection_logic_edge_cases(host_lib):
    from test_oracle_loopclose_ref import _edge_databases
    for name, db in _edge_databases():
        d, n_kf, covis, stride = oracle_lib._db_args(db)
        for min_score in (0.0, 0.02, 0.9):
            cand, words, score = np.zeros(max(n_kf, 1), np.int32), np.zeros(max(n_kf, 1), np.int32), np.full(max(n_kf, 1), -1.0, np.float32)
            n = host_lib.host_detect_loop_candidates(d["q_word"].ctypes.data, d["q_val"].ctypes.data, len(d["q_word"]), n_kf, d["off"].ctypes.data, d["word"].ctypes.data,
                                                     d["val"].ctypes.data, covis, stride, d["connected"].ctypes.data, min_score, cand.ctypes.data, words.ctypes.data,
                                                     score.ctypes.data)
            oc, ow, os_ = oracle_lib.detect_loop_candidates(db, min_score)
            assert n == len(oc) and np.array_equal(cand[:n], oc) and np.array_equal(words[:n_kf], ow) and np.array_equal(score[:n_kf], os_), (name, min_score)
