"""CPU: sanity of the projection-search oracle on the synthetic sequence (map built from frame 10, searched from frame 11)."""
import numpy as np

import oracle_lib
from planarslam_b200 import synth, synth_map


def scenario(f0=10, f1=11, pose_noise=0.002, seed=0):
    rng = np.random.default_rng(seed)
    g0, d0, _, _ = synth.render_frame(2, f0)
    g1, d1, _, _ = synth.render_frame(2, f1)
    k0, de0 = oracle_lib.orb_extract(g0)
    k1, de1 = oracle_lib.orb_extract(g1)
    fa0, fa1 = synth_map.frame_arrays(k0, de0, d0), synth_map.frame_arrays(k1, de1, d1)
    T0, T1 = synth_map.true_pose(f0), synth_map.true_pose(f1)
    m = synth_map.map_from_frame(fa0, T0)
    T1n = T1.copy()
    T1n[:3, 3] += rng.normal(0, pose_noise, 3)
    fv = synth_map.frame_view(fa1, T1n)
    lf = dict(n=fa0["n"], keys=fa0["keys_un"], map_point=np.full(fa0["n"], -1, np.int32), outlier=np.zeros(fa0["n"], np.uint8),
              Tcw=np.ascontiguousarray(T0, np.float32))
    lf["map_point"][m["src_index"]] = np.arange(m["n"], dtype=np.int32)
    return fv, m, lf


def test_search_by_projection_map_finds_consistent_matches():
    fv, m, lf = scenario()
    n, matches, in_view = oracle_lib.search_by_projection_map(fv, m, 3.0, 0.8, np.full(fv["n"], -1, np.int32))
    assert in_view.sum() > 0.7 * m["n"]
    assert n == (matches >= 0).sum() and n > 150
    # a matched keypoint lies where its map point projects (within the search radius)
    idx = np.nonzero(matches >= 0)[0]
    T = np.asarray(fv["Tcw"], np.float64)
    Xc = m["pos"][matches[idx]].astype(np.float64) @ T[:3, :3].T + T[:3, 3]
    u = fv["fx"] * Xc[:, 0] / Xc[:, 2] + fv["cx"]
    v = fv["fy"] * Xc[:, 1] / Xc[:, 2] + fv["cy"]
    err = np.hypot(u - fv["keys_un"]["x"][idx], v - fv["keys_un"]["y"][idx])
    assert np.median(err) < 3.0 and err.max() < 60.0
    # already-matched keypoints holding a point with observations are never stolen
    pre = np.full(fv["n"], -1, np.int32)
    pre[idx[:20]] = 0
    n2, matches2, _ = oracle_lib.search_by_projection_map(fv, m, 3.0, 0.8, pre)
    assert (matches2[idx[:20]] == 0).all()


def test_search_by_projection_last_frame():
    fv, m, lf = scenario()
    n, matches = oracle_lib.search_by_projection_last(fv, lf, m, 15.0, False, True, np.full(fv["n"], -1, np.int32))
    assert n == (matches >= 0).sum() and n > 150
    n_no_ori, _ = oracle_lib.search_by_projection_last(fv, lf, m, 15.0, False, False, np.full(fv["n"], -1, np.int32))
    assert n_no_ori >= n


def _py_search_by_projection_map(fv, m, th, nnratio, matches0, cos_limit=0.5):
    """Independent plain-Python statement of Tracking::SearchLocalPoints' visibility pass + ORBmatcher::SearchByProjection(Frame&, vector<MapPoint*>&, th)
    (src/Frame.cc:312-367, :439-490, :526-535; src/MapPoint.cc:419-434; src/ORBmatcher.cc:46-140), written from the reference with numpy float32
    scalars so that every operation rounds as the C++ float code does."""
    f = np.float32
    T = np.asarray(fv["Tcw"], np.float32)
    R, t = T[:3, :3], T[:3, 3]
    fx, fy, cx, cy, bf = (f(fv[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    mnx, mxx, mny, mxy = (f(fv[k]) for k in ("min_x", "max_x", "min_y", "max_y"))
    sf, nlev, logsf = np.asarray(fv["scale_factors"], np.float32), int(fv["n_levels"]), f(fv["log_scale_factor"])
    Ow = np.array([f(sum(float(-R[k, r]) * float(t[k]) for k in range(3))) for r in range(3)], np.float32)        # -R^T t, cv::Mat product
    invw, invh = f(64) / f(mxx - mnx), f(48) / f(mxy - mny)
    kp = fv["keys_un"]
    grid = {}
    for i in range(fv["n"]):                                          # Frame::AssignFeaturesToGrid / PosInGrid: round(), not floor()
        gx, gy = float(f(f(f(kp["x"][i]) - mnx) * invw)), float(f(f(f(kp["y"][i]) - mny) * invh))
        px = int(np.floor(gx + 0.5)) if gx >= 0 else -int(np.floor(-gx + 0.5))          # C round(): halves away from zero (numpy rounds to even)
        py = int(np.floor(gy + 0.5)) if gy >= 0 else -int(np.floor(-gy + 0.5))
        if 0 <= px < 64 and 0 <= py < 48:
            grid.setdefault((px, py), []).append(i)
    matches = np.asarray(matches0, np.int32).copy()
    in_view = np.zeros(m["n"], np.uint8)
    n_matches = 0
    for k in range(m["n"]):
        if m["skip"][k]:
            continue
        P = m["pos"][k].astype(np.float32)
        Pc = np.array([f(f(sum(float(R[r, c]) * float(P[c]) for c in range(3))) + t[r]) for r in range(3)], np.float32)
        if Pc[2] < f(0):
            continue
        invz = f(1) / Pc[2]
        u, v = f(f(f(fx * Pc[0]) * invz) + cx), f(f(f(fy * Pc[1]) * invz) + cy)
        if u < mnx or u > mxx or v < mny or v > mxy:
            continue
        maxd, mind = f(f(1.2) * m["max_distance"][k]), f(f(0.8) * m["min_distance"][k])
        PO = (P - Ow).astype(np.float32)
        dist = f(np.sqrt(sum(float(x) * float(x) for x in PO)))
        if dist < mind or dist > maxd:
            continue
        view_cos = f(sum(float(PO[c]) * float(m["normal"][k][c]) for c in range(3)) / float(dist))
        if view_cos < f(cos_limit):
            continue
        ratio = f(m["max_distance"][k] / dist)
        lvl = int(np.ceil(f(f(np.log(float(ratio))) / logsf)))
        lvl = 0 if lvl < 0 else (nlev - 1 if lvl >= nlev else lvl)
        in_view[k] = 1
        ur_proj = f(u - f(bf * invz))
        r = f(2.5) if view_cos > f(0.998) else f(4.0)
        if th != 1.0:
            r = f(r * f(th))
        rr = f(r * sf[lvl])
        min_level, max_level = lvl - 1, lvl
        x0, x1 = max(0, int(np.floor(float(f(f(f(u - mnx) - rr) * invw))))), min(63, int(np.ceil(float(f(f(f(u - mnx) + rr) * invw)))))
        y0, y1 = max(0, int(np.floor(float(f(f(f(v - mny) - rr) * invh))))), min(47, int(np.ceil(float(f(f(f(v - mny) + rr) * invh)))))
        if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
            continue
        best, best2, blev, blev2, bidx = 256, 256, -1, -1, -1
        for ix in range(x0, x1 + 1):
            for iy in range(y0, y1 + 1):
                for i in grid.get((ix, iy), ()):
                    octv = int(kp["octave"][i])
                    if octv < min_level or octv > max_level:
                        continue
                    if not (abs(f(f(kp["x"][i]) - u)) < rr and abs(f(f(kp["y"][i]) - v)) < rr):
                        continue
                    if matches[i] >= 0 and m["has_obs"][matches[i]]:
                        continue
                    if fv["u_right"][i] > 0 and abs(f(ur_proj - f(fv["u_right"][i]))) > rr:
                        continue
                    d = int(np.unpackbits(m["desc"][k] ^ fv["desc"][i]).sum())
                    if d < best:
                        best2, best, blev2, blev, bidx = best, d, blev, octv, i
                    elif d < best2:
                        blev2, best2 = octv, d
        if best <= 100:
            if blev == blev2 and best > f(nnratio) * best2:
                continue
            matches[bidx] = k
            n_matches += 1
    return n_matches, matches, in_view


def test_search_by_projection_map_oracle_matches_independent_python():
    for seed, th, nnr in ((0, 3.0, 0.8), (1, 1.0, 0.8), (2, 5.0, 0.9)):
        fv, m, lf = scenario(f0=10 + seed, f1=11 + seed, seed=seed)
        m["skip"][::17] = 1                                            # some points already seen in this frame / bad
        pre = np.full(fv["n"], -1, np.int32)
        pre[::23] = 3                                                  # key points that already hold a map point with observations
        n, matches, in_view = oracle_lib.search_by_projection_map(fv, m, th, nnr, pre)
        pn, pmatches, pin_view = _py_search_by_projection_map(fv, m, th, nnr, pre)
        assert np.array_equal(in_view, pin_view), seed
        assert n == pn and np.array_equal(matches, pmatches), seed
        assert n > 100


class _PyFrame:
    """Plain-Python Frame pieces shared by the independent statements: the feature grid (PosInGrid rounds, src/Frame.cc:526-535) and
    GetFeaturesInArea (:439-490)."""

    def __init__(self, fv):
        f = np.float32
        self.fv, self.kp = fv, fv["keys_un"]
        self.mnx, self.mxx, self.mny, self.mxy = (f(fv[k]) for k in ("min_x", "max_x", "min_y", "max_y"))
        self.invw, self.invh = f(64) / f(self.mxx - self.mnx), f(48) / f(self.mxy - self.mny)
        self.grid = {}
        for i in range(fv["n"]):
            gx, gy = float(f(f(f(self.kp["x"][i]) - self.mnx) * self.invw)), float(f(f(f(self.kp["y"][i]) - self.mny) * self.invh))
            px = int(np.floor(gx + 0.5)) if gx >= 0 else -int(np.floor(-gx + 0.5))
            py = int(np.floor(gy + 0.5)) if gy >= 0 else -int(np.floor(-gy + 0.5))
            if 0 <= px < 64 and 0 <= py < 48:
                self.grid.setdefault((px, py), []).append(i)

    def features_in_area(self, u, v, r, min_level, max_level=-1):
        f = np.float32
        x0, x1 = max(0, int(np.floor(float(f(f(f(u - self.mnx) - r) * self.invw))))), min(63, int(np.ceil(float(f(f(f(u - self.mnx) + r) * self.invw)))))
        y0, y1 = max(0, int(np.floor(float(f(f(f(v - self.mny) - r) * self.invh))))), min(47, int(np.ceil(float(f(f(f(v - self.mny) + r) * self.invh)))))
        if x0 >= 64 or x1 < 0 or y0 >= 48 or y1 < 0:
            return []
        check = min_level > 0 or max_level >= 0
        out = []
        for ix in range(x0, x1 + 1):
            for iy in range(y0, y1 + 1):
                for i in self.grid.get((ix, iy), ()):
                    o = int(self.kp["octave"][i])
                    if check and (o < min_level or (max_level >= 0 and o > max_level)):
                        continue
                    if abs(f(f(self.kp["x"][i]) - u)) < r and abs(f(f(self.kp["y"][i]) - v)) < r:
                        out.append(i)
        return out


def _py_search_by_projection_last(fv, lf, m, th, mono, check_ori, matches0):
    """Independent statement of ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono) (src/ORBmatcher.cc:1396-1535) with
    ComputeThreeMaxima (:1666-1707)."""
    f = np.float32
    F = _PyFrame(fv)
    T, Tl = np.asarray(fv["Tcw"], np.float32), np.asarray(lf["Tcw"], np.float32)
    R, t = T[:3, :3], T[:3, 3]
    fx, fy, cx, cy, bf = (f(fv[k]) for k in ("fx", "fy", "cx", "cy", "bf"))
    sf = np.asarray(fv["scale_factors"], np.float32)
    twc = np.array([f(sum(float(-R[k, r]) * float(t[k]) for k in range(3))) for r in range(3)], np.float32)
    tlc = np.array([f(f(sum(float(Tl[r, c]) * float(twc[c]) for c in range(3))) + Tl[r, 3]) for r in range(3)], np.float32)
    mb = f(bf / fx)
    forward, backward = (tlc[2] > mb) and not mono, (-tlc[2] > mb) and not mono
    matches = np.asarray(matches0, np.int32).copy()
    hist = [[] for _ in range(30)]
    n = 0
    for i in range(lf["n"]):
        k = int(lf["map_point"][i])
        if k < 0 or lf["outlier"][i]:
            continue
        P = m["pos"][k].astype(np.float32)
        xc = np.array([f(f(sum(float(R[r, c]) * float(P[c]) for c in range(3))) + t[r]) for r in range(3)], np.float32)
        invzc = f(1.0 / float(xc[2]))
        if invzc < 0:
            continue
        u, v = f(f(f(fx * xc[0]) * invzc) + cx), f(f(f(fy * xc[1]) * invzc) + cy)
        if u < F.mnx or u > F.mxx or v < F.mny or v > F.mxy:
            continue
        octv = int(lf["keys"]["octave"][i])
        radius = f(f(th) * sf[octv])
        if forward:
            cand = F.features_in_area(u, v, radius, octv)
        elif backward:
            cand = F.features_in_area(u, v, radius, 0, octv)
        else:
            cand = F.features_in_area(u, v, radius, octv - 1, octv + 1)
        best, bidx = 256, -1
        for i2 in cand:
            if matches[i2] >= 0 and m["has_obs"][matches[i2]]:
                continue
            if fv["u_right"][i2] > 0:
                ur = f(u - f(bf * invzc))
                if abs(f(ur - f(fv["u_right"][i2]))) > radius:
                    continue
            d = int(np.unpackbits(m["desc"][k] ^ fv["desc"][i2]).sum())
            if d < best:
                best, bidx = d, i2
        if best <= 100:
            matches[bidx] = k
            n += 1
            if check_ori:
                rot = f(f(lf["keys"]["angle"][i]) - f(F.kp["angle"][bidx]))
                if rot < 0:
                    rot = f(rot + f(360))
                g = float(f(rot * f(f(1.0) / f(30))))
                b = int(np.floor(g + 0.5))
                if b == 30:
                    b = 0
                hist[b].append(bidx)
    if check_ori:
        mx, ind = [0, 0, 0], [-1, -1, -1]
        for i in range(30):
            sz = len(hist[i])
            if sz > mx[0]:
                mx, ind = [sz, mx[0], mx[1]], [i, ind[0], ind[1]]
            elif sz > mx[1]:
                mx, ind = [mx[0], sz, mx[1]], [ind[0], i, ind[1]]
            elif sz > mx[2]:
                mx[2], ind[2] = sz, i
        if mx[1] < f(0.1) * f(mx[0]):
            ind[1] = ind[2] = -1
        elif mx[2] < f(0.1) * f(mx[0]):
            ind[2] = -1
        for i in range(30):
            if i not in ind:
                for j in hist[i]:
                    matches[j] = -1
                    n -= 1
    return n, matches


def test_search_by_projection_last_oracle_matches_independent_python():
    for seed, th, mono, ori in ((0, 15.0, False, True), (1, 7.0, False, True), (2, 15.0, True, False), (3, 30.0, False, True)):
        fv, m, lf = scenario(f0=10 + 2 * seed, f1=11 + 2 * seed + (seed == 3), seed=seed)
        lf["outlier"][::13] = 1
        pre = np.full(fv["n"], -1, np.int32)
        pre[::29] = 5
        n, matches = oracle_lib.search_by_projection_last(fv, lf, m, th, mono, ori, pre)
        pn, pmatches = _py_search_by_projection_last(fv, lf, m, th, mono, ori, pre)
        assert n == pn and np.array_equal(matches, pmatches), seed
        assert n > 100
