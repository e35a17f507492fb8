"""GPU: ORBmatcher::SearchByBoW through the C ABI vs the CPU oracle: identical matches and counts."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth_lines

pytestmark = pytest.mark.gpu


def test_search_by_bow_matches_oracle():
    from planarslam_b200._lib import Context
    from planarslam_b200.matcher import search_by_bow
    ctx = Context(640, 480, 1)
    tot = 0
    for seed in range(6):
        kf, f = synth_lines.make_bow_pair(seed, n_kf=1000 if seed % 2 else 2000, n_f=1000, n_nodes=300 if seed < 4 else 40)
        for ratio, ori in ((0.7, True), (0.9, False), (0.6, True)):
            n, m = search_by_bow(ctx, kf, f, ratio, ori)
            on, om = oracle_lib.search_by_bow(kf, f, ratio, ori)
            assert n == on and np.array_equal(m, om), (seed, ratio, ori)
            tot += n
    assert tot > 3000
