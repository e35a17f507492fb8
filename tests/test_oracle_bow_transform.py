"""CPU: oracle DBoW2 transform (oracle/bow_transform.cc, restating the vendored Thirdparty/DBoW2) against a plain-Python re-derivation
on a synthetic vocabulary tree, plus properties of the L1 score."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth_lines


def _py_transform(voc, feats, levelsup):
    v, fv = {}, {}
    nid_level = voc["L"] - levelsup
    for i, f in enumerate(feats):
        nid, cur, lvl = 0, 0, 0
        while True:
            lvl += 1
            ch = voc["child_id"][voc["child_off"][cur]:voc["child_off"][cur + 1]]
            d = [int(np.unpackbits(f ^ voc["desc"][c]).sum()) for c in ch]
            cur = int(ch[int(np.argmin(d))])                  # first minimum
            if lvl == nid_level:
                nid = cur
            if voc["child_off"][cur + 1] == voc["child_off"][cur]:
                break
        w = float(voc["weight"][cur])
        if w > 0:
            wid = int(voc["word_id"][cur])
            v[wid] = v[wid] + w if wid in v else w
            fv.setdefault(nid, []).append(i)
    norm = 0.0
    for k in sorted(v):
        norm += abs(v[k])
    ids = sorted(v)
    return ids, [v[k] / norm for k in ids] if norm > 0 else [v[k] for k in ids], {k: fv[k] for k in sorted(fv)}


def test_bow_transform_matches_python():
    for seed, (k, L, lup) in enumerate([(4, 3, 1), (10, 3, 2), (3, 5, 4), (5, 2, 4)]):
        voc = synth_lines.make_vocabulary(seed, k=k, L=L)
        feats = synth_lines.make_features_for_vocabulary(seed, voc, n=150)
        o = oracle_lib.bow_transform(voc, feats, lup)
        ids, vals, fv = _py_transform(voc, feats, lup)
        assert o["word_id"].tolist() == ids
        assert np.array_equal(o["word_val"], np.array(vals))
        assert o["node_id"].tolist() == list(fv.keys())
        for j, key in enumerate(fv):
            assert o["node_feat"][o["node_off"][j]:o["node_off"][j + 1]].tolist() == fv[key]
        assert abs(o["word_val"].sum() - 1.0) < 1e-12
        if L - lup <= 0:
            assert o["node_id"].tolist() == [0]                # levelsup >= L: every feature lands in the root


def test_bow_score_properties():
    voc = synth_lines.make_vocabulary(3, k=6, L=3)
    a = oracle_lib.bow_transform(voc, synth_lines.make_features_for_vocabulary(1, voc, 300), 2)
    b = oracle_lib.bow_transform(voc, synth_lines.make_features_for_vocabulary(2, voc, 300), 2)
    assert abs(oracle_lib.bow_score_l1(a, a) - 1.0) < 1e-12
    s = oracle_lib.bow_score_l1(a, b)
    assert 0.0 <= s < 1.0 and abs(s - oracle_lib.bow_score_l1(b, a)) < 1e-15
