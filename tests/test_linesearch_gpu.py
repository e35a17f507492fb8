"""GPU: LSDmatcher::SearchByProjection through the C ABI vs the CPU oracle: identical assignments and match counts."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth_lines

pytestmark = pytest.mark.gpu


def test_line_search_matches_oracle():
    from planarslam_b200.matcher import LSDmatcher
    total = 0
    for seed in range(8):
        f, m = synth_lines.make_line_search(seed, n_frame=40 if seed % 2 else 64, n_map=150)
        for th, ratio in ((3.0, 0.6), (1.0, 0.9), (5.0, 0.7)):
            n, a = LSDmatcher(ratio).SearchByProjection(f, m, th)
            on, oa = oracle_lib.line_search_by_projection(f, m, th, ratio)
            assert n == on and np.array_equal(a, oa), (seed, th)
            total += n
    assert total > 100
