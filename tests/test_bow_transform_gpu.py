"""GPU: the DBoW2 vocabulary transform through the C ABI vs the CPU oracle: identical BowVector (word ids and double values, bit for
bit: the weights are accumulated by the same repeated additions and normalised by the same in-order L1 sum) and FeatureVector."""
import numpy as np
import pytest

import oracle_lib
from planarslam_b200 import synth_lines

pytestmark = pytest.mark.gpu


def test_bow_transform_matches_oracle():
    from planarslam_b200._lib import Context
    from planarslam_b200.matcher import bow_transform
    ctx = Context(640, 480, 1)
    for seed, (k, L, lup, n) in enumerate([(10, 4, 2, 1000), (10, 3, 4, 2000), (4, 5, 4, 500), (32, 2, 1, 300), (10, 6, 4, 1000)]):
        voc = synth_lines.make_vocabulary(seed, k=k, L=L) if L < 6 else synth_lines.make_vocabulary(seed, k=4, L=6)
        feats = synth_lines.make_features_for_vocabulary(seed, voc, n=n)
        r, o = bow_transform(ctx, voc, feats, lup), oracle_lib.bow_transform(voc, feats, lup)
        for key in ("word_id", "word_val", "node_id", "node_off", "node_feat"):
            assert np.array_equal(r[key], o[key]), (seed, key)
        assert len(r["word_id"]) > 10
