"""GPU: Frame::isInFrustum(MapLine*) through the C ABI (pslam_lines_in_frustum) vs the CPU oracle; bar: bit-exact fields."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200.synth_lines import make_line_frustum

pytestmark = pytest.mark.gpu


def test_lines_in_frustum_match_oracle():
    from planarslam_b200._lib import Context
    from planarslam_b200.matcher import lines_in_frustum
    ctx = Context(640, 480, max_batch=1)
    for seed, n in ((0, 400), (1, 1), (2, 3000)):
        frame, pos, nrm, max_d, min_d = make_line_frustum(seed, n=n)
        o = oracle_lib.lines_in_frustum(frame, pos, nrm, max_d, min_d, 0.6)
        cnt, g = lines_in_frustum(ctx, frame, pos, nrm, max_d, min_d, 0.6)
        assert cnt == int(o["in_view"].sum())
        for k in ("in_view", "proj", "level", "view_cos"):
            assert np.array_equal(g[k], o[k]), (seed, k)
