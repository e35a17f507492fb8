"""GPU: Tracking::TrackManhattanFrame through the C ABI (pslam_track_manhattan_batch) vs the CPU oracle (oracle/manhattan.cc).
Bar: identical counts / found flags / membership masks, rotation to 2e-6 (float results; asin / exp / tan come from two libms)."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200.synth_manhattan import make_manhattan

pytestmark = pytest.mark.gpu


def test_track_manhattan_matches_oracle():
    from planarslam_b200._lib import Context
    from planarslam_b200.manhattan import TrackManhattanFrame
    cases = [dict(seed=s) for s in range(4)] + [dict(seed=3, weights=(0.5, 0.5, 0.0), clutter=0.02, n_lines=0), dict(seed=4, weights=(1.0, 0.0, 0.0), clutter=0.0, n_lines=0),
                                                 dict(seed=7, n_normals=300, n_lines=40, perturb_deg=8.0)]
    data = [make_manhattan(**kw) for kw in cases]
    ctx = Context(640, 480, max_batch=1)
    res, nmasks, dmasks = TrackManhattanFrame(ctx, np.stack([d[0] for d in data]), [d[1] for d in data], [d[2] for d in data])
    for f, d in enumerate(data):
        o = oracle_lib.track_manhattan_frame(d[0], d[1], d[2])
        for k in ("found", "n_cone", "n_selected"):
            assert np.array_equal(res[f][k], o[k]), (f, k)
        assert res[f]["min_num"] == o["min_num"] and res[f]["svd_applied"] == o["svd_applied"], f
        assert np.allclose(res[f]["R"], o["R"], rtol=0, atol=2e-6) and np.allclose(res[f]["density"], o["density"], rtol=1e-6), f
        assert np.array_equal(nmasks[f] & 7, o["normal_mask"]) and np.array_equal(dmasks[f] & 7, o["dir_mask"]), f
