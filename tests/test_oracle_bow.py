"""CPU: oracle ORBmatcher::SearchByBoW (oracle/bow.cc) against a plain-Python re-derivation on seeded inputs."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth_lines


def _py_ref(kf, f, ratio, ori):
    nf = len(f["angle"])
    match = np.full(nf, -1, np.int32)
    hist = [[] for _ in range(30)]
    kn = {int(i): k for k, i in enumerate(kf["node_id"])}
    n = 0
    for b, nid in enumerate(f["node_id"]):
        a = kn.get(int(nid))
        if a is None:
            continue
        for ik in kf["node_feat"][kf["node_off"][a]:kf["node_off"][a + 1]]:
            if not kf["has_mp"][ik]:
                continue
            b1, bi, b2 = 256, -1, 256
            for jf in f["node_feat"][f["node_off"][b]:f["node_off"][b + 1]]:
                if match[jf] >= 0:
                    continue
                d = int(np.unpackbits(kf["desc"][ik] ^ f["desc"][jf]).sum())
                if d < b1:
                    b2, b1, bi = b1, d, jf
                elif d < b2:
                    b2 = d
            if b1 <= 50 and np.float32(b1) < np.float32(ratio) * np.float32(b2):
                match[bi] = ik
                if ori:
                    rot = np.float32(kf["angle"][ik] - f["angle"][bi])
                    if rot < 0:
                        rot = np.float32(rot + np.float32(360.0))
                    v = float(np.float32(rot * np.float32(1.0 / 30)))
                    bn = int(np.floor(v + 0.5))                     # round half away from zero for v >= 0
                    hist[0 if bn == 30 else bn].append(bi)
                n += 1
    if ori:
        m1 = m2 = m3 = 0
        i1 = i2 = i3 = -1
        for i in range(30):
            s = len(hist[i])
            if s > m1:
                m3, m2, m1, i3, i2, i1 = m2, m1, s, i2, i1, i
            elif s > m2:
                m3, m2, i3, i2 = m2, s, i2, i
            elif s > m3:
                m3, i3 = s, i
        if m2 < np.float32(0.1) * np.float32(m1):
            i2 = i3 = -1
        elif m3 < np.float32(0.1) * np.float32(m1):
            i3 = -1
        for i in range(30):
            if i in (i1, i2, i3):
                continue
            for j in hist[i]:
                match[j] = -1
                n -= 1
    return n, match


def test_search_by_bow_oracle_matches_python():
    tot = 0
    for seed in range(3):
        kf, f = synth_lines.make_bow_pair(seed, n_kf=400, n_f=380, n_nodes=90)
        for ratio, ori in ((0.7, True), (0.9, False)):
            n, m = oracle_lib.search_by_bow(kf, f, ratio, ori)
            n2, m2 = _py_ref(kf, f, ratio, ori)
            assert n == n2 and np.array_equal(m, m2), (seed, ratio, ori)
            tot += n
    assert tot > 300
