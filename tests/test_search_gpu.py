"""GPU: projection-guided searches through the C ABI vs the CPU oracle — identical match assignments."""
import numpy as np
import pytest

import oracle_lib
from test_oracle_search import scenario

pytestmark = pytest.mark.gpu


def test_search_by_projection_map_matches_oracle():
    from planarslam_b200.matcher import ORBmatcher
    m8 = ORBmatcher(0.8)
    for (f0, f1, th, noise) in [(10, 11, 3.0, 0.002), (20, 22, 5.0, 0.01), (30, 31, 1.0, 0.0), (40, 44, 3.0, 0.02)]:
        fv, m, lf = scenario(f0, f1, noise, seed=f0)
        rng = np.random.default_rng(f0)
        pre = np.full(fv["n"], -1, np.int32)
        pre[rng.choice(fv["n"], 50, replace=False)] = rng.integers(0, m["n"], 50)      # matches from an earlier stage
        m["has_obs"][rng.choice(m["n"], m["n"] // 10, replace=False)] = 0              # a few fresh points without observations
        m["skip"][rng.choice(m["n"], m["n"] // 20, replace=False)] = 1
        n, matches, in_view = m8.SearchByProjection(fv, m, th, pre)
        on, omatches, oin = oracle_lib.search_by_projection_map(fv, m, th, 0.8, pre)
        assert np.array_equal(in_view, oin), (f0, f1)
        assert n == on and np.array_equal(matches, omatches), (f0, f1, n, on, (matches != omatches).sum())
        assert n > 50


def test_search_by_projection_last_matches_oracle():
    from planarslam_b200.matcher import ORBmatcher
    for check_ori in (True, False):
        mm = ORBmatcher(0.9, check_ori)
        for (f0, f1, th, mono) in [(10, 11, 15.0, False), (20, 23, 15.0, False), (30, 31, 7.0, True)]:
            fv, m, lf = scenario(f0, f1, 0.005, seed=f1)
            lf["outlier"][::17] = 1
            n, matches = mm.SearchByProjectionLast(fv, lf, m, th, mono)
            on, omatches = oracle_lib.search_by_projection_last(fv, lf, m, th, mono, check_ori, np.full(fv["n"], -1, np.int32))
            assert n == on and np.array_equal(matches, omatches), (f0, f1, check_ori, n, on)
            assert n > 50
