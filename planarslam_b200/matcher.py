"""Host-side mirror of the reference's brute-force descriptor searches (include/ORBmatcher.h:37-64,
include/LSDmatcher.h:21-24) on top of the C ABI."""
from __future__ import annotations

import ctypes as C

import numpy as np

from ._lib import Context, FrameView, LastFrame, MapPoints


def _frame_view(fv: dict) -> FrameView:
    s = FrameView()
    s._keep = fv
    s.n = fv["n"]
    for k in ("keys_un", "u_right", "desc", "scale_factors"):
        setattr(s, k, fv[k].ctypes.data)
    s.Tcw = (C.c_float * 16)(*np.asarray(fv["Tcw"], np.float32).ravel().tolist())
    for k in ("fx", "fy", "cx", "cy", "bf", "min_x", "max_x", "min_y", "max_y", "log_scale_factor"):
        setattr(s, k, fv[k])
    s.n_levels = fv["n_levels"]
    return s


def _map_points(m: dict) -> MapPoints:
    s = MapPoints()
    s._keep = m
    s.n = m["n"]
    for k in ("pos", "normal", "max_distance", "min_distance", "desc", "skip", "has_obs"):
        setattr(s, k, m[k].ctypes.data)
    return s


def _last_frame(lf: dict) -> LastFrame:
    s = LastFrame()
    s._keep = lf
    s.n = lf["n"]
    for k in ("keys", "map_point", "outlier"):
        setattr(s, k, lf[k].ctypes.data)
    s.Tcw = (C.c_float * 16)(*np.asarray(lf["Tcw"], np.float32).ravel().tolist())
    return s


class ORBmatcher:
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30          # src/ORBmatcher.cc:38-40

    def __init__(self, nnratio: float = 0.6, checkOri: bool = True, ctx: Context | None = None):
        self.mfNNratio, self.mbCheckOrientation = nnratio, checkOri
        self.ctx = ctx or Context(640, 480, 1)

    def knn2(self, q: np.ndarray, t: np.ndarray, gate: bool = False):
        q, t = np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)
        nq, nt = len(q), len(t)
        idx, dist = np.full((max(nq, 1), 2), -1, np.int32), np.full((max(nq, 1), 2), 256, np.int32)
        good, ng = np.zeros(max(nq, 1), np.int32), C.c_int32(0)
        self.ctx.check(self.ctx.L.pslam_hamming_knn2(self.ctx.h, q.ctypes.data if nq else None, nq, t.ctypes.data if nt else None, nt,
                                                     idx.ctypes.data, dist.ctypes.data, good.ctypes.data if gate else None, C.byref(ng)))
        return idx[:nq], dist[:nq], good[:ng.value]

    # int SearchByProjection(Frame& F, const vector<MapPoint*>& vpMapPoints, const float th)
    def SearchByProjection(self, frame_view: dict, map_points: dict, th: float = 3.0, matches=None):
        m = np.full(frame_view["n"], -1, np.int32) if matches is None else np.ascontiguousarray(matches, np.int32).copy()
        in_view = np.zeros(max(map_points["n"], 1), np.uint8)
        n = self.ctx.L.pslam_search_by_projection_map(self.ctx.h, C.byref(_frame_view(frame_view)), C.byref(_map_points(map_points)), th,
                                                      self.mfNNratio, m.ctypes.data, in_view.ctypes.data)
        if n < 0:
            self.ctx.check(n)
        return n, m, in_view[:map_points["n"]]

    # int SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, const float th, const bool bMono)
    def SearchByProjectionLast(self, cur_view: dict, last: dict, map_points: dict, th: float, bMono: bool = False, matches=None):
        m = np.full(cur_view["n"], -1, np.int32) if matches is None else np.ascontiguousarray(matches, np.int32).copy()
        n = self.ctx.L.pslam_search_by_projection_last(self.ctx.h, C.byref(_frame_view(cur_view)), C.byref(_last_frame(last)),
                                                       C.byref(_map_points(map_points)), th, int(bMono), int(self.mbCheckOrientation), m.ctypes.data)
        if n < 0:
            self.ctx.check(n)
        return n, m

    @staticmethod
    def DescriptorDistance(a: np.ndarray, b: np.ndarray) -> int:
        return int(np.unpackbits(np.bitwise_xor(a, b)).sum())

    def MatchORBPoints(self, cur_desc: np.ndarray, last_desc: np.ndarray):
        """Returns (NPair, [(queryIdx, trainIdx, distance)] of the good matches)."""
        idx, dist, good = self.knn2(cur_desc, last_desc, gate=True)
        return len(good), [(int(i), int(idx[i, 0]), int(dist[i, 0])) for i in good]


class PlaneMatcher:
    """PlaneMatcher(dTh, aTh, verTh, parTh), include/PlaneMatcher.h:16-30."""

    def __init__(self, dTh: float = 0.1, aTh: float = 0.86, verTh: float = 0.08716, parTh: float = 0.9962, ctx: Context | None = None):
        self.dTh, self.aTh, self.verTh, self.parTh = dTh, aTh, verTh, parTh
        self.ctx = ctx or Context(640, 480, 1)

    def SearchMapByCoefficients(self, Tcw, frame_coef, map_coef, map_bad, pts_off, pts):
        Tcw = np.ascontiguousarray(Tcw, np.float32)
        fc, mc = np.ascontiguousarray(frame_coef, np.float32).reshape(-1, 4), np.ascontiguousarray(map_coef, np.float32).reshape(-1, 4)
        bad, off, pts = np.ascontiguousarray(map_bad, np.uint8), np.ascontiguousarray(pts_off, np.int32), np.ascontiguousarray(pts, np.float32)
        nf = len(fc)
        out = [np.full(max(nf, 1), -1, np.int32) for _ in range(3)]
        n = self.ctx.L.pslam_plane_match(self.ctx.h, Tcw.ctypes.data, nf, fc.ctypes.data, len(mc), mc.ctypes.data, bad.ctypes.data, off.ctypes.data,
                                         pts.ctypes.data, self.dTh, self.aTh, self.verTh, self.parTh, *[o.ctypes.data for o in out])
        if n < 0:
            self.ctx.check(n)
        return n, out[0][:nf], out[1][:nf], out[2][:nf]


class LSDmatcher:
    """Mirror of Planar_SLAM::LSDmatcher for the projection search (include/LSDmatcher.h:21-24)."""

    def __init__(self, nnratio: float = 0.6, ctx: Context | None = None):
        self.nnratio = nnratio
        self.ctx = ctx or Context(640, 480, 1)
        self.orb = ORBmatcher(ctx=self.ctx)

    # int SearchByDescriptor(KeyFrame* pKF, Frame& F, std::vector<MapLine*>& vpMapLineMatches)
    def SearchByDescriptor(self, kf_desc: np.ndarray, frame_desc: np.ndarray):
        """knn-2 of the key frame's line descriptors in the frame + ratio test 1/1.5 (src/LSDmatcher.cpp:256-276)."""
        idx, dist, _ = self.orb.knn2(kf_desc, frame_desc)
        keep = [(i, int(idx[i, 0])) for i in range(len(idx)) if idx[i, 1] >= 0 and dist[i, 0] / max(dist[i, 1], 1e-9) < 1.0 / 1.5]
        return len(keep), keep

    # int SearchByProjection(Frame& F, const std::vector<MapLine*>& vpMapLines, const float th = 3)
    def SearchByProjection(self, frame: dict, map_lines: dict, th: float = 3.0):
        """frame: pt [n][2] f32, angle [n] f32, octave [n] i32, desc [n][32] u8, has_obs [n] u8, scale_factors [L] f32;
        map_lines: skip u8, level i32, view_cos f32, proj [m][4] f32, desc [m][32] u8, has_obs u8.  Returns (nmatches, assigned)."""
        f = {k: np.ascontiguousarray(v) for k, v in frame.items()}
        m = {k: np.ascontiguousarray(v) for k, v in map_lines.items()}
        nf, nm = len(f["angle"]), len(m["level"])
        assigned = np.full(max(nf, 1), -1, np.int32)
        n = self.ctx.L.pslam_line_search_by_projection(
            self.ctx.h, nf, f["pt"].ctypes.data, f["angle"].ctypes.data, f["octave"].ctypes.data, f["desc"].ctypes.data, f["has_obs"].ctypes.data,
            f["scale_factors"].ctypes.data, len(f["scale_factors"]), nm, m["skip"].ctypes.data, m["level"].ctypes.data, m["view_cos"].ctypes.data,
            m["proj"].ctypes.data, m["desc"].ctypes.data, m["has_obs"].ctypes.data, th, self.nnratio, assigned.ctypes.data)
        if n < 0:
            self.ctx.check(n)
        return n, assigned[:nf]


def search_by_bow(ctx: Context, kf: dict, frame: dict, nnratio: float = 0.7, check_orientation: bool = True):
    """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF, F, vpMapPointMatches).  kf / frame: desc [n][32] u8, angle [n] f32, node_id,
    node_off, node_feat i32 (DBoW2 FeatureVector as CSR); kf additionally has_mp [n] u8.  Returns (nmatches, match [n_frame])."""
    k = {a: np.ascontiguousarray(b) for a, b in kf.items()}
    f = {a: np.ascontiguousarray(b) for a, b in frame.items()}
    nf = len(f["angle"])
    match = np.full(max(nf, 1), -1, np.int32)
    n = ctx.L.pslam_search_by_bow(ctx.h, len(k["angle"]), k["desc"].ctypes.data, k["angle"].ctypes.data, k["has_mp"].ctypes.data, len(k["node_id"]),
                                  k["node_id"].ctypes.data, k["node_off"].ctypes.data, k["node_feat"].ctypes.data, nf, f["desc"].ctypes.data,
                                  f["angle"].ctypes.data, len(f["node_id"]), f["node_id"].ctypes.data, f["node_off"].ctypes.data,
                                  f["node_feat"].ctypes.data, nnratio, 1 if check_orientation else 0, match.ctypes.data)
    if n < 0:
        ctx.check(n)
    return n, match[:nf]


def bow_transform(ctx: Context, voc: dict, features: np.ndarray, levelsup: int = 4) -> dict:
    """ORBVocabulary::transform(vCurrentDesc, mBowVec, mFeatVec, levelsup).  voc: L, desc [nodes][32] u8, child_off, child_id, word_id i32, weight f64
    (planarslam_b200.synth_lines.make_vocabulary builds synthetic ones).  Returns dict(word_id, word_val, node_id, node_off, node_feat)."""
    f = np.ascontiguousarray(features, np.uint8)
    n = len(f)
    o = dict(word_id=np.zeros(max(n, 1), np.int32), word_val=np.zeros(max(n, 1)), node_id=np.zeros(max(n, 1), np.int32),
             node_off=np.zeros(n + 1, np.int32), node_feat=np.zeros(max(n, 1), np.int32))
    cnt = np.zeros(2, np.int32)
    ctx.check(ctx.L.pslam_bow_transform(ctx.h, len(voc["word_id"]), voc["L"], voc["desc"].ctypes.data, voc["child_off"].ctypes.data,
                                        voc["child_id"].ctypes.data, voc["word_id"].ctypes.data, voc["weight"].ctypes.data, f.ctypes.data, n, levelsup,
                                        o["word_id"].ctypes.data, o["word_val"].ctypes.data, o["node_id"].ctypes.data, o["node_off"].ctypes.data,
                                        o["node_feat"].ctypes.data, cnt.ctypes.data))
    nw, nn = int(cnt[0]), int(cnt[1])
    o["word_id"], o["word_val"] = o["word_id"][:nw], o["word_val"][:nw]
    o["node_id"], o["node_off"] = o["node_id"][:nn], o["node_off"][:nn + 1]
    o["node_feat"] = o["node_feat"][:int(o["node_off"][-1])] if nn else o["node_feat"][:0]
    return o


def lines_in_frustum(ctx: Context, frame: dict, pos, normal, max_distance, min_distance, cos_limit: float = 0.6):
    """bool Frame::isInFrustum(MapLine*, float viewingCosLimit) (src/Frame.cc:369-437) for n map lines - the visibility pass of
    Tracking::SearchLocalLines.  frame: dict(Tcw 4x4 float32, fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor).
    Returns (n_in_view, dict(in_view, proj [n][4], level, view_cos)) - the map-line fields LSDmatcher.SearchByProjection takes."""
    fv = np.concatenate([np.asarray(frame["Tcw"], np.float32).ravel(), np.array([frame[k] for k in ("fx", "fy", "cx", "cy", "min_x", "max_x", "min_y", "max_y",
                                                                                                      "log_scale_factor")], np.float32)])
    P, Nn = np.ascontiguousarray(pos, np.float64).reshape(-1, 6), np.ascontiguousarray(normal, np.float64).reshape(-1, 3)
    mx, mn = np.ascontiguousarray(max_distance, np.float32), np.ascontiguousarray(min_distance, np.float32)
    n = len(P)
    o = dict(in_view=np.zeros(n, np.uint8), proj=np.zeros((n, 4), np.float32), level=np.zeros(n, np.int32), view_cos=np.zeros(n, np.float32))
    rc = ctx.L.pslam_lines_in_frustum(ctx.h, fv.ctypes.data, n, P.ctypes.data, Nn.ctypes.data, mx.ctypes.data, mn.ctypes.data, cos_limit, o["in_view"].ctypes.data,
                                      o["proj"].ctypes.data, o["level"].ctypes.data, o["view_cos"].ctypes.data)
    if rc < 0:
        ctx.check(rc)
    return rc, o


def search_by_bow_kf(ctx: Context, kf1: dict, kf2: dict, nnratio: float = 0.75, check_orientation: bool = True):
    """ORBmatcher(nnratio, checkOri).SearchByBoW(pKF1, pKF2, vpMatches12) - the loop-closure matcher (src/ORBmatcher.cc:526-659).  kf1 / kf2: desc [n][32] u8,
    angle [n] f32 (mvKeysUn), has_mp [n] u8, node_id, node_off, node_feat i32.  Returns (nmatches, match12 [n1]: feature of key frame 2 or -1)."""
    a = {k: np.ascontiguousarray(v) for k, v in kf1.items()}
    b = {k: np.ascontiguousarray(v) for k, v in kf2.items()}
    n1 = len(a["angle"])
    match = np.full(max(n1, 1), -1, np.int32)
    fn = ctx.L.pslam_search_by_bow_kf
    fn.argtypes = [C.c_void_p] + [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p] * 2 + [C.c_float, C.c_int, C.c_void_p]
    side = lambda s: (len(s["angle"]), s["desc"].ctypes.data, s["angle"].ctypes.data, s["has_mp"].ctypes.data, len(s["node_id"]), s["node_id"].ctypes.data,
                      s["node_off"].ctypes.data, s["node_feat"].ctypes.data)
    n = fn(ctx.h, *side(a), *side(b), nnratio, 1 if check_orientation else 0, match.ctypes.data)
    if n < 0:
        ctx.check(n)
    return n, match[:n1]


class KeyFrameDatabase:
    """The candidate searches of Planar_SLAM::KeyFrameDatabase (include/KeyFrameDatabase.h) over a database whose BowVectors are resident in HBM.
    db: off [n_kf + 1], word, val (CSR, key frames in KeyFrameDatabase::add order), covis [n_kf][<= 10] (GetBestCovisibilityKeyFrames(10), -1 padded)."""

    def __init__(self, ctx: Context, off, word, val, covis=None):
        self.ctx = ctx
        self.off, self.word, self.val = np.ascontiguousarray(off, np.int32), np.ascontiguousarray(word, np.int32), np.ascontiguousarray(val, np.float64)
        self.n_kf = len(self.off) - 1
        self.covis = np.ascontiguousarray(covis, np.int32) if covis is not None else np.full((max(self.n_kf, 1), 1), -1, np.int32)
        L = ctx.L
        L.pslam_bow_database_set.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.pslam_detect_loop_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_float] + [C.c_void_p] * 3
        L.pslam_detect_relocalization_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int] + [C.c_void_p] * 3
        ctx.check(L.pslam_bow_database_set(ctx.h, self.n_kf, self.off.ctypes.data, self.word.ctypes.data, self.val.ctypes.data))

    def DetectLoopCandidates(self, q_word, q_val, min_score: float, connected=None, sentinel: float = -1.0):
        """Returns (candidates, mnLoopWords [n_kf], mLoopScore [n_kf]; `sentinel` where the reference does not evaluate the score)."""
        qw, qv = np.ascontiguousarray(q_word, np.int32), np.ascontiguousarray(q_val, np.float64)
        cand, words = np.zeros(max(self.n_kf, 1), np.int32), np.zeros(max(self.n_kf, 1), np.int32)
        score = np.full(max(self.n_kf, 1), sentinel, np.float32)
        con = np.ascontiguousarray(connected, np.uint8) if connected is not None else None
        n = self.ctx.L.pslam_detect_loop_candidates(self.ctx.h, len(qw), qw.ctypes.data, qv.ctypes.data, self.covis.ctypes.data, self.covis.shape[1],
                                                    con.ctypes.data if con is not None else None, min_score, cand.ctypes.data, words.ctypes.data, score.ctypes.data)
        if n < 0:
            self.ctx.check(n)
        return cand[:n].copy(), words[:self.n_kf], score[:self.n_kf]

    def DetectRelocalizationCandidates(self, q_word, q_val, reloc_score):
        """reloc_score: KeyFrame::mRelocScore of every database key frame before the call.  Returns (candidates, mnRelocWords, mRelocScore after)."""
        qw, qv = np.ascontiguousarray(q_word, np.int32), np.ascontiguousarray(q_val, np.float64)
        cand, words = np.zeros(max(self.n_kf, 1), np.int32), np.zeros(max(self.n_kf, 1), np.int32)
        score = np.ascontiguousarray(reloc_score, np.float32).copy()
        n = self.ctx.L.pslam_detect_relocalization_candidates(self.ctx.h, len(qw), qw.ctypes.data, qv.ctypes.data, self.covis.ctypes.data, self.covis.shape[1],
                                                              score.ctypes.data, cand.ctypes.data, words.ctypes.data)
        if n < 0:
            self.ctx.check(n)
        return cand[:n].copy(), words[:self.n_kf], score
