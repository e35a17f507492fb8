"""GPU: Frame::isLineGood through the C ABI (pslam_lines3d_batch) vs the CPU oracle (oracle/line3d.cc).

Bar: identical accept flags, inlier sets, rand() draw counts, end points (they are back-projected depth samples: exact) and
mvDepthLine.  The kernel body is the host-checked line3d_body.h (tests/test_line3d_host.py: identical to the oracle on 16 runs)."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_lines3d_match_oracle():
    from planarslam_b200._lib import Context
    from planarslam_b200.lines import KEYLINE_DTYPE, isLineGood
    nf = 6
    frames = [synth.render_frame(seed=s, frame=3 * s) for s in range(nf)]
    kls = [oracle_lib.extract_line_segments(f[0], 40)[0] for f in frames]
    kl = np.zeros((nf, 40), KEYLINE_DTYPE)
    for f in range(nf):
        kl[f, :len(kls[f])] = kls[f]
    kl[nf - 1, 30:] = 0                                                          # a frame with fewer lines than max_lines
    n_lines = np.array([len(k) for k in kls], np.int32)
    n_lines[nf - 1] = 30
    d16 = np.stack([f[1] if k % 2 == 0 else synth.noisy_depth(f[1], k, 0.1 + 0.05 * k, 0.004 * k) for k, f in enumerate(frames)])     # odd frames: RANSAC has to work
    seeds = np.array([1, 1, 5, 99, 2 ** 31 + 7, 0], np.uint32)
    skips = np.array([0, 13, 0, 250, 0, 1], np.int32)
    factor = np.float32(1.0 / synth.DEPTH_FACTOR)
    ctx = Context(640, 480, max_batch=nf)
    out, drawn = isLineGood(ctx, kl, n_lines, d16, synth.TUM3_K, factor, seeds, skips)
    n_valid = 0
    for f in range(nf):
        o = oracle_lib.lines3d_frame(kl[f, :n_lines[f]], d16[f].astype(np.float32) * factor, synth.TUM3_K, seed=int(seeds[f]), skip=int(skips[f]))
        g = out[f, :n_lines[f]]
        assert drawn[f] == o["n_drawn"], f
        assert np.array_equal(g["valid"], o["valid"]) and np.array_equal(g["n_points"], o["n_points"]), f
        assert np.array_equal(g["inliers"], o["inliers"]) and np.array_equal(g["n_inliers"], o["n_inliers"]), f
        assert np.array_equal(np.concatenate([g["A"], g["B"]], 1), o["lines3d"]) and np.array_equal(g["depth"], o["depth_line"]), f
        assert np.array_equal(g["director"], o["director"], equal_nan=True), f
        assert not out[f, n_lines[f]:]["valid"].any()
        n_valid += int(g["valid"].sum())
    assert n_valid > 100 and drawn.sum() > 800
