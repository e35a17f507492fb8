"""GPU: the LBD line descriptors (pslam_lines_extract_describe_batch = the whole LineSegment::ExtractLineSegment, src/LSDextractor.cpp:13-39) vs the CPU
oracle (oracle/lbd.cc; descriptor logic parity-unpinned upstream, primitives pinned to cv2 - tests/test_oracle_lbd.py).  Bar: the 72 floats and the 32
descriptor bytes bit-exact; and the descriptors feed LSDmatcher::SearchByDescriptor (knn-2 with the 1/1.5 ratio) to geometrically consistent matches."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth

pytestmark = pytest.mark.gpu


def test_lbd_bit_exact_vs_oracle_and_feeds_the_line_matcher():
    from planarslam_b200.lines import LineSegment
    from planarslam_b200.matcher import LSDmatcher
    g = np.stack([synth.render_frame(2, f)[0] for f in (10, 11, 30)] + [synth.polygon_image(11)])
    ls = LineSegment(max_batch=len(g))
    res = ls.ExtractLineSegmentWithDescriptors(g, 40)
    for f in range(len(g)):
        kl, lf, desc, lbd = res[f]
        okl, olf = oracle_lib.extract_line_segments(g[f], 40)
        assert len(kl) == len(okl) == 40 and np.array_equal(kl["startPointX"], okl["startPointX"]) and np.array_equal(lf, olf)
        olbd, odesc = oracle_lib.lbd_compute(g[f], kl)
        assert np.array_equal(lbd, olbd), (f, np.abs(lbd - olbd).max())
        assert np.array_equal(desc, odesc), f
    # frames 10 and 11 of the sequence: LSDmatcher::SearchByDescriptor on the device descriptors
    (k0, _, d0, _), (k1, _, d1, _) = res[0], res[1]
    n, keep = LSDmatcher(0.6, ctx=ls.ctx).SearchByDescriptor(d0, d1)
    assert n == len(keep) >= 8
    geo = np.array([np.hypot(k0["pt"][i, 0] - k1["pt"][j, 0], k0["pt"][i, 1] - k1["pt"][j, 1]) for i, j in keep])
    assert np.median(geo) < 12.0, (np.median(geo), geo)
