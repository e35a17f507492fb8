// TEST INFRASTRUCTURE ONLY.  The reference's own matchers - src/ORBmatcher.cc (SearchByProjection x2, SearchByBoW), src/LSDmatcher.cpp (SearchByProjection),
// src/PlaneMatcher.cpp (SearchMapByCoefficients), with the Frame / KeyFrame /
// MapPoint / Map classes they walk (src/Frame.cc: SetPose, AssignFeaturesToGrid, PosInGrid, GetFeaturesInArea, isInFrustum; src/MapPoint.cc:
// PredictScale, GetDescriptor ...; src/KeyFrame.cc) - compiled unmodified against the stand-ins of oracle/ref/shims/ into oracle/_ref/libmatch_ref.so.
// This driver only BUILDS the object graph from the plain arrays of the C ABI views (pslam_frame_view, pslam_map_points, pslam_last_frame ...) and
// reads the result back; every decision (frustum test, candidate gathering, descriptor gates, rotation histogram) is taken by the reference's code.
// What of Tracking is restated: SearchLocalPoints' loop (src/Tracking.cc: skip the points already seen / bad, isInFrustum(pMP, 0.5) for the rest).
// The reference's members are protected; the driver opens them for itself only (the reference's translation units are compiled as they are).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <set>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include <chrono>
#include <tuple>

#define private public
#define protected public
#include "Frame.h"
#include "KeyFrame.h"
#include "Map.h"
#include "MapPoint.h"
#include "MapLine.h"
#include "MapPlane.h"
#include "ORBmatcher.h"
#include "LSDmatcher.h"
#include "PlaneMatcher.h"
#include "Optimizer.h"
#include "Config.h"
#include "KeyFrameDatabase.h"
#include "ORBVocabulary.h"
#undef private
#undef protected
#include "pslam_abi.h"

// Which implementation the entry points below exercise.  Default: the reference's own classes (libmatch_ref.so, ref_* symbols).  oracle/ref/adapter_driver.cc
// includes this file with the product's reference-typed adapter (include/pslam_reference_adapter.hpp) in their place (libadapter_ref.so, adp_* symbols): the
// same object graphs, the same read-back, only the function under test differs.
#ifdef PSLAM_ADAPTER_BUILD
#include "pslam_reference_adapter.hpp"
#endif
#ifndef DRV
#define DRV(name) ref_##name
#define DRV_ORBMATCHER ORBmatcher
#define DRV_LSDMATCHER LSDmatcher
#define DRV_PLANEMATCHER PlaneMatcher
#define DRV_OPTIMIZER Optimizer
#endif

using namespace Planar_SLAM;

// src/LSDextractor.cpp wraps OpenCV-contrib's LSDDetector / BinaryDescriptor classes, which are not in this image; the only caller is Frame's image
// constructor (Frame.cc:170-176), which no pinned path runs.  Defined here so that Frame.cc links.
#ifndef PSLAM_ADAPTER_BUILD
namespace Planar_SLAM {
void LineSegment::ExtractLineSegment(const cv::Mat&, std::vector<cv::line_descriptor::KeyLine>&, cv::Mat&, std::vector<Eigen::Vector3d>&, float, int) {
    std::fprintf(stderr, "oracle/ref/match_driver.cc: LineSegment::ExtractLineSegment is a link-only stand-in\n");
    std::abort();
}
}  // namespace Planar_SLAM
#endif

namespace {

cv::KeyPoint to_kp(const pslam_keypoint& k) {
    cv::KeyPoint kp;
    kp.pt.x = k.x; kp.pt.y = k.y; kp.size = k.size; kp.angle = k.angle; kp.response = k.response; kp.octave = k.octave; kp.class_id = k.class_id;
    return kp;
}
cv::Mat mat44(const float* T) { cv::Mat m(4, 4, CV_32F); for (int i = 0; i < 16; ++i) m.at<float>(i / 4, i % 4) = T[i]; return m; }

// a Frame carrying what the matchers read (everything else stays default-constructed)
void fill_frame(Frame& F, const pslam_frame_view& v, unsigned long id) {
    F.mnId = id;
    F.N = v.n;
    F.mvKeysUn.resize(v.n);
    for (int i = 0; i < v.n; ++i) F.mvKeysUn[i] = to_kp(v.keys_un[i]);
    F.mvKeys = F.mvKeysUn;
    F.mvuRight.assign(v.u_right, v.u_right + v.n);
    F.mvDepth.assign(v.n, -1.f);
    F.mDescriptors = cv::Mat(v.n, 32, CV_8U);
    for (int i = 0; i < v.n; ++i) std::memcpy(F.mDescriptors.ptr(i), v.desc + 32 * (size_t)i, 32);
    Frame::fx = v.fx; Frame::fy = v.fy; Frame::cx = v.cx; Frame::cy = v.cy; Frame::invfx = 1.0f / v.fx; Frame::invfy = 1.0f / v.fy;
    Frame::mnMinX = v.min_x; Frame::mnMaxX = v.max_x; Frame::mnMinY = v.min_y; Frame::mnMaxY = v.max_y;
    Frame::mfGridElementWidthInv = static_cast<float>(FRAME_GRID_COLS) / static_cast<float>(Frame::mnMaxX - Frame::mnMinX);     // Frame.cc:124-125
    Frame::mfGridElementHeightInv = static_cast<float>(FRAME_GRID_ROWS) / static_cast<float>(Frame::mnMaxY - Frame::mnMinY);
    F.mbf = v.bf; F.mb = F.mbf / Frame::fx;
    F.mnScaleLevels = v.n_levels;
    F.mvScaleFactors.assign(v.scale_factors, v.scale_factors + v.n_levels);
    F.mfScaleFactor = v.n_levels > 1 ? v.scale_factors[1] : 1.2f;
    F.mfLogScaleFactor = v.log_scale_factor;
    F.mvLevelSigma2.resize(v.n_levels); F.mvInvLevelSigma2.resize(v.n_levels); F.mvInvScaleFactors.resize(v.n_levels);
    for (int l = 0; l < v.n_levels; ++l) { F.mvLevelSigma2[l] = v.scale_factors[l] * v.scale_factors[l]; F.mvInvLevelSigma2[l] = 1.0f / F.mvLevelSigma2[l]; F.mvInvScaleFactors[l] = 1.0f / v.scale_factors[l]; }
    F.mvpMapPoints.assign(v.n, static_cast<MapPoint*>(NULL));
    F.mvbOutlier.assign(v.n, false);
    F.mK = cv::Mat::eye(3, 3, CV_32F);
    F.mK.at<float>(0, 0) = v.fx; F.mK.at<float>(1, 1) = v.fy; F.mK.at<float>(0, 2) = v.cx; F.mK.at<float>(1, 2) = v.cy;
    F.SetPose(mat44(v.Tcw));
    F.AssignFeaturesToGrid();
}

struct World {                          // one Map, one anchor key frame (MapPoint's constructor reads its ids), the map points of a pslam_map_points view
    Map map;
    Frame anchor_frame;
    KeyFrame* anchor = nullptr;
    std::vector<MapPoint*> pts;
    std::unordered_map<MapPoint*, int> index;
    World() {
        anchor_frame.N = 0;
        anchor_frame.mTcw = cv::Mat::eye(4, 4, CV_32F);
        anchor_frame.SetPose(cv::Mat::eye(4, 4, CV_32F));
        anchor = new KeyFrame(anchor_frame, &map, static_cast<KeyFrameDatabase*>(NULL));
    }
    MapPoint* make_point(const float* pos, const float* normal, float maxd, float mind, const uint8_t* desc, bool has_obs) {
        cv::Mat X(3, 1, CV_32F);
        for (int c = 0; c < 3; ++c) X.at<float>(c) = pos[c];
        MapPoint* p = new MapPoint(X, anchor, &map);
        p->mNormalVector = cv::Mat(3, 1, CV_32F);
        for (int c = 0; c < 3; ++c) p->mNormalVector.at<float>(c) = normal ? normal[c] : 0.f;
        p->mfMaxDistance = maxd; p->mfMinDistance = mind;
        p->mDescriptor = cv::Mat(1, 32, CV_8U);
        if (desc) std::memcpy(p->mDescriptor.ptr(0), desc, 32); else std::memset(p->mDescriptor.ptr(0), 0, 32);
        p->nObs = has_obs ? 1 : 0;
        p->mbTrackInView = false;
        return p;
    }
    void add(const pslam_map_points& m) {
        for (int i = 0; i < m.n; ++i) {
            MapPoint* p = make_point(m.pos + 3 * i, m.normal + 3 * i, m.max_distance[i], m.min_distance[i], m.desc + 32 * (size_t)i, m.has_obs[i] != 0);
            index[p] = (int)pts.size();
            pts.push_back(p);
        }
    }
    ~World() { for (MapPoint* p : pts) delete p; delete anchor; }
};

void hold(Frame& F, const World& w, const int32_t* matches) {
    for (int i = 0; i < F.N; ++i) F.mvpMapPoints[i] = matches[i] >= 0 ? w.pts[matches[i]] : static_cast<MapPoint*>(NULL);
}
void read_back(const Frame& F, const World& w, int32_t* matches) {
    for (int i = 0; i < F.N; ++i) { MapPoint* p = F.mvpMapPoints[i]; matches[i] = p ? w.index.at(p) : -1; }
}

DBoW2::FeatureVector feat_vec(int n_nodes, const int32_t* node_id, const int32_t* node_off, const int32_t* node_feat) {
    DBoW2::FeatureVector fv;
    for (int k = 0; k < n_nodes; ++k)
        for (int j = node_off[k]; j < node_off[k + 1]; ++j) fv.addFeature((DBoW2::NodeId)node_id[k], (unsigned)node_feat[j]);
    return fv;
}

}  // namespace

// Tracking::SearchLocalPoints + ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)   src/ORBmatcher.cc:46-130
extern "C" int DRV(search_by_projection_map)(const pslam_frame_view* fv, const pslam_map_points* m, float th, float nnratio, int32_t* matches_io, uint8_t* in_view) {
    World w;
    w.add(*m);
    Frame F;
    fill_frame(F, *fv, 7);
    hold(F, w, matches_io);
    for (int i = 0; i < m->n; ++i) {
        MapPoint* pMP = w.pts[i];
        if (m->skip[i]) { pMP->mbTrackInView = false; continue; }
        F.isInFrustum(pMP, 0.5);
    }
    DRV_ORBMATCHER matcher(nnratio);
    const int n = matcher.SearchByProjection(F, w.pts, th);
    read_back(F, w, matches_io);
    if (in_view) for (int i = 0; i < m->n; ++i) in_view[i] = w.pts[i]->mbTrackInView ? 1 : 0;
    return n;
}

// ORBmatcher::SearchByProjection(Frame& CurrentFrame, const Frame& LastFrame, th, bMono)   src/ORBmatcher.cc:1396-1535
extern "C" int DRV(search_by_projection_last)(const pslam_frame_view* cur, const pslam_last_frame* last, const pslam_map_points* m, float th, int mono,
                                             int check_orientation, int32_t* matches_io) {
    World w;
    w.add(*m);
    Frame C;
    fill_frame(C, *cur, 8);
    hold(C, w, matches_io);
    Frame L;
    L.mnId = 7;
    L.N = last->n;
    L.mvKeys.resize(last->n);
    for (int i = 0; i < last->n; ++i) L.mvKeys[i] = to_kp(last->keys[i]);
    L.mvKeysUn = L.mvKeys;
    L.mvpMapPoints.assign(last->n, static_cast<MapPoint*>(NULL));
    L.mvbOutlier.assign(last->n, false);
    for (int i = 0; i < last->n; ++i) { if (last->map_point[i] >= 0) L.mvpMapPoints[i] = w.pts[last->map_point[i]]; L.mvbOutlier[i] = last->outlier[i] != 0; }
    L.SetPose(mat44(last->Tcw));
    DRV_ORBMATCHER matcher(0.9, check_orientation != 0);
    const int n = matcher.SearchByProjection(C, L, th, mono != 0);
    read_back(C, w, matches_io);
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)   src/ORBmatcher.cc:160-292
extern "C" int DRV(search_by_bow)(int n_kf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_mp, int kf_nodes, const int32_t* kf_node_id,
                                 const int32_t* kf_node_off, const int32_t* kf_node_feat, int n_f, const uint8_t* f_desc, const float* f_angle, int f_nodes,
                                 const int32_t* f_node_id, const int32_t* f_node_off, const int32_t* f_node_feat, float nnratio, int check_orientation, int32_t* match) {
    World w;
    const float zero3[3] = {0, 0, 0};
    auto basic = [](Frame& F, int n, const uint8_t* desc, const float* angle) {
        F.N = n;
        F.mvKeysUn.resize(n);
        for (int i = 0; i < n; ++i) { F.mvKeysUn[i].angle = angle[i]; F.mvKeysUn[i].octave = 0; }
        F.mvKeys = F.mvKeysUn;
        F.mDescriptors = cv::Mat(n, 32, CV_8U);
        for (int i = 0; i < n; ++i) std::memcpy(F.mDescriptors.ptr(i), desc + 32 * (size_t)i, 32);
        F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
        F.mvbOutlier.assign(n, false);
        F.SetPose(cv::Mat::eye(4, 4, CV_32F));
    };
    Frame K;
    basic(K, n_kf, kf_desc, kf_angle);
    K.mFeatVec = feat_vec(kf_nodes, kf_node_id, kf_node_off, kf_node_feat);
    for (int i = 0; i < n_kf; ++i)
        if (kf_has_mp[i]) { MapPoint* p = w.make_point(zero3, zero3, 0, 0, nullptr, true); w.index[p] = i; w.pts.push_back(p); K.mvpMapPoints[i] = p; }
    KeyFrame* pKF = new KeyFrame(K, &w.map, static_cast<KeyFrameDatabase*>(NULL));
    Frame F;
    basic(F, n_f, f_desc, f_angle);
    F.mFeatVec = feat_vec(f_nodes, f_node_id, f_node_off, f_node_feat);
    DRV_ORBMATCHER matcher(nnratio, check_orientation != 0);
    std::vector<MapPoint*> vpMapPointMatches;
    const int n = matcher.SearchByBoW(pKF, F, vpMapPointMatches);
    for (int i = 0; i < n_f; ++i) match[i] = vpMapPointMatches[i] ? w.index.at(vpMapPointMatches[i]) : -1;     // index of the key-frame feature whose map point was taken
    delete pKF;
    return n;
}

// ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)   src/ORBmatcher.cc:526-659 (the loop-closure matcher, LoopClosing::ComputeSim3 :265)
extern "C" int DRV(search_by_bow_kf)(int n1, const uint8_t* desc1, const float* angle1, const uint8_t* has_mp1, int nodes1, const int32_t* node_id1, const int32_t* node_off1,
                                    const int32_t* node_feat1, int n2, const uint8_t* desc2, const float* angle2, const uint8_t* has_mp2, int nodes2,
                                    const int32_t* node_id2, const int32_t* node_off2, const int32_t* node_feat2, float nnratio, int check_orientation, int32_t* match12) {
    World w;
    const float zero3[3] = {0, 0, 0};
    std::unordered_map<MapPoint*, int> index2;
    auto make_kf = [&](int n, const uint8_t* desc, const float* angle, const uint8_t* has_mp, int nodes, const int32_t* id, const int32_t* off, const int32_t* feat,
                       std::unordered_map<MapPoint*, int>* index) {
        Frame F;
        F.N = n;
        F.mvKeysUn.resize(n);
        for (int i = 0; i < n; ++i) { F.mvKeysUn[i].angle = angle[i]; F.mvKeysUn[i].octave = 0; }
        F.mvKeys = F.mvKeysUn;
        F.mDescriptors = cv::Mat(n, 32, CV_8U);
        for (int i = 0; i < n; ++i) std::memcpy(F.mDescriptors.ptr(i), desc + 32 * (size_t)i, 32);
        F.mvpMapPoints.assign(n, static_cast<MapPoint*>(NULL));
        F.mvbOutlier.assign(n, false);
        F.SetPose(cv::Mat::eye(4, 4, CV_32F));
        F.mFeatVec = feat_vec(nodes, id, off, feat);
        for (int i = 0; i < n; ++i)
            if (has_mp[i]) { MapPoint* p = w.make_point(zero3, zero3, 0, 0, nullptr, true); w.pts.push_back(p); F.mvpMapPoints[i] = p; if (index) (*index)[p] = i; }
        return new KeyFrame(F, &w.map, static_cast<KeyFrameDatabase*>(NULL));
    };
    KeyFrame* k1 = make_kf(n1, desc1, angle1, has_mp1, nodes1, node_id1, node_off1, node_feat1, nullptr);
    KeyFrame* k2 = make_kf(n2, desc2, angle2, has_mp2, nodes2, node_id2, node_off2, node_feat2, &index2);
    DRV_ORBMATCHER matcher(nnratio, check_orientation != 0);
    std::vector<MapPoint*> vpMatches12;
    const int n = matcher.SearchByBoW(k1, k2, vpMatches12);
    for (int i = 0; i < n1; ++i) match12[i] = vpMatches12[i] ? index2.at(vpMatches12[i]) : -1;
    delete k1; delete k2;
    return n;
}

// KeyFrameDatabase::DetectLoopCandidates / DetectRelocalizationCandidates   src/KeyFrameDatabase.cc:76-305 with DBoW2's L1 score (Thirdparty/DBoW2/DBoW2/ScoringObject.cpp)
namespace {
DBoW2::BowVector bow_vec(const int32_t* word, const double* val, int n) {
    DBoW2::BowVector v;
    for (int i = 0; i < n; ++i) v.insert(std::make_pair((DBoW2::WordId)word[i], (DBoW2::WordValue)val[i]));
    return v;
}
struct LoopWorld {
    World w;
#ifdef PSLAM_ADAPTER_BUILD
    pslam_adapter::ref::KeyFrameDatabase db;
    KeyFrameDatabase* ref_db() { return static_cast<KeyFrameDatabase*>(NULL); }
#else
    ORBVocabulary voc;                   // default: TF_IDF weighting, L1_NORM scoring (the reference's ORBvoc.txt header says the same)
    KeyFrameDatabase db;
    KeyFrameDatabase* ref_db() { return &db; }
#endif
    std::vector<KeyFrame*> kfs;
    std::unordered_map<KeyFrame*, int> index;
    LoopWorld(int n_kf, const int32_t* off, const int32_t* word, const double* val, const int32_t* covis, int covis_stride, int32_t max_query_word)
#ifndef PSLAM_ADAPTER_BUILD
        : db(voc)
#endif
    {
#ifndef PSLAM_ADAPTER_BUILD
        int32_t max_word = max_query_word;
        for (int i = 0; i < off[n_kf]; ++i) max_word = std::max(max_word, word[i]);
        db.mvInvertedFile.clear();
        db.mvInvertedFile.resize((size_t)max_word + 1);
#else
        (void)max_query_word;
#endif
        for (int k = 0; k < n_kf; ++k) {
            Frame F;
            F.N = 0;
            F.SetPose(cv::Mat::eye(4, 4, CV_32F));
            F.mBowVec = bow_vec(word + off[k], val + off[k], off[k + 1] - off[k]);
            KeyFrame* kf = new KeyFrame(F, &w.map, ref_db());
            index[kf] = k;
            kfs.push_back(kf);
            db.add(kf);
        }
        for (int k = 0; k < n_kf && covis; ++k)
            for (int j = 0; j < covis_stride; ++j) {
                const int k2 = covis[(size_t)k * covis_stride + j];
                if (k2 < 0) break;
                kfs[k]->mvpOrderedConnectedKeyFrames.push_back(kfs[k2]);
            }
    }
    ~LoopWorld() { for (KeyFrame* k : kfs) delete k; }
};
}  // namespace

extern "C" int DRV(detect_loop_candidates)(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word, const double* val,
                                          const int32_t* covis, int covis_stride, const uint8_t* connected, float min_score, int32_t* cand, int32_t* common_words,
                                          float* score) {
    LoopWorld L(n_kf, off, word, val, covis, covis_stride, n_q ? q_word[n_q - 1] : 0);
    Frame F;
    F.N = 0;
    F.SetPose(cv::Mat::eye(4, 4, CV_32F));
    F.mBowVec = bow_vec(q_word, q_val, n_q);
    KeyFrame* q = new KeyFrame(F, &L.w.map, L.ref_db());
    for (int k = 0; k < n_kf; ++k) {
        if (connected && connected[k]) q->mConnectedKeyFrameWeights[L.kfs[k]] = 1;
        L.kfs[k]->mLoopScore = score[k];           // uninitialised in the reference's constructor: the caller's sentinel shows where it was evaluated
    }
    const std::vector<KeyFrame*> r = L.db.DetectLoopCandidates(q, min_score);
    for (size_t i = 0; i < r.size(); ++i) cand[i] = L.index.at(r[i]);
    for (int k = 0; k < n_kf; ++k) { common_words[k] = L.kfs[k]->mnLoopWords; score[k] = L.kfs[k]->mLoopScore; }
    delete q;
    return (int)r.size();
}

extern "C" int DRV(detect_relocalization_candidates)(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word,
                                                    const double* val, const int32_t* covis, int covis_stride, float* reloc_score_io, int32_t* cand,
                                                    int32_t* common_words) {
    LoopWorld L(n_kf, off, word, val, covis, covis_stride, n_q ? q_word[n_q - 1] : 0);
    Frame F;
    F.N = 0;
    F.SetPose(cv::Mat::eye(4, 4, CV_32F));
    F.mBowVec = bow_vec(q_word, q_val, n_q);
    F.mnId = 1000000;                              // mnRelocQuery starts at 0: a frame id of 0 would list nothing
    for (int k = 0; k < n_kf; ++k) L.kfs[k]->mRelocScore = reloc_score_io[k];
    const std::vector<KeyFrame*> r = L.db.DetectRelocalizationCandidates(&F);
    for (size_t i = 0; i < r.size(); ++i) cand[i] = L.index.at(r[i]);
    for (int k = 0; k < n_kf; ++k) { common_words[k] = L.kfs[k]->mnRelocWords; reloc_score_io[k] = L.kfs[k]->mRelocScore; }
    return (int)r.size();
}

#ifndef PSLAM_ADAPTER_BUILD
// Frame::isInFrustum(MapLine*, viewingCosLimit)   src/Frame.cc:369-437 (with MapLine::PredictScale / Get{Min,Max}DistanceInvariance, src/MapLine.cpp:364-390)
// fv: Tcw[16], fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor
extern "C" void ref_lines_in_frustum(const float* fv, int n, const double* pos, const double* normal, const float* max_distance, const float* min_distance,
                                     float cos_limit, uint8_t* in_view, float* proj, int32_t* level, float* view_cos) {
    World w;
    Frame F;
    Frame::fx = fv[16]; Frame::fy = fv[17]; Frame::cx = fv[18]; Frame::cy = fv[19];
    Frame::mnMinX = fv[20]; Frame::mnMaxX = fv[21]; Frame::mnMinY = fv[22]; Frame::mnMaxY = fv[23];
    F.mfLogScaleFactor = fv[24];
    F.SetPose(mat44(fv));
    for (int i = 0; i < n; ++i) {
        Vector6d P;
        for (int c = 0; c < 6; ++c) P(c) = pos[6 * i + c];
        MapLine* pML = new MapLine(P, w.anchor, &w.map);
        pML->mNormalVector = Eigen::Vector3d(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        pML->mfMaxDistance = max_distance[i]; pML->mfMinDistance = min_distance[i];
        pML->mTrackProjX1 = pML->mTrackProjY1 = pML->mTrackProjX2 = pML->mTrackProjY2 = 0; pML->mnTrackScaleLevel = 0; pML->mTrackViewCos = 0;
        F.isInFrustum(pML, cos_limit);
        in_view[i] = pML->mbTrackInView ? 1 : 0;
        proj[4 * i] = pML->mTrackProjX1; proj[4 * i + 1] = pML->mTrackProjY1; proj[4 * i + 2] = pML->mTrackProjX2; proj[4 * i + 3] = pML->mTrackProjY2;
        level[i] = pML->mnTrackScaleLevel; view_cos[i] = pML->mTrackViewCos;
        delete pML;
    }
}

// LSDmatcher::SearchByProjection(Frame&, const vector<MapLine*>&, th)   src/LSDmatcher.cpp:141-211 (Frame::GetLinesInArea src/Frame.cc:491-523)
#endif  // !PSLAM_ADAPTER_BUILD

extern "C" int DRV(line_search_by_projection)(int nf, const float* pt, const float* angle, const int32_t* octave, const uint8_t* desc, const uint8_t* has_obs,
                                             const float* scale_factors, int n_levels, int nm, const uint8_t* skip, const int32_t* level, const float* view_cos,
                                             const float* proj, const uint8_t* mdesc, const uint8_t* m_has_obs, float th, float nnratio, int32_t* assigned) {
    World w;
    Vector6d zero6 = Vector6d::Zero();
    Frame F;
    F.NL = nf;
    F.mvKeylinesUn.resize(nf);
    for (int i = 0; i < nf; ++i) { F.mvKeylinesUn[i].pt.x = pt[2 * i]; F.mvKeylinesUn[i].pt.y = pt[2 * i + 1]; F.mvKeylinesUn[i].angle = angle[i]; F.mvKeylinesUn[i].octave = octave[i]; }
    F.mLdesc = cv::Mat(nf, 32, CV_8U);
    for (int i = 0; i < nf; ++i) std::memcpy(F.mLdesc.ptr(i), desc + 32 * (size_t)i, 32);
    F.mvScaleFactors.assign(scale_factors, scale_factors + n_levels);
    std::vector<MapLine*> own;
    F.mvpMapLines.assign(nf, static_cast<MapLine*>(NULL));
    for (int i = 0; i < nf; ++i)
        if (has_obs[i]) { MapLine* p = new MapLine(zero6, w.anchor, &w.map); p->nObs = 1; own.push_back(p); F.mvpMapLines[i] = p; }     // already holds an observed map line
    std::vector<MapLine*> held = F.mvpMapLines;
    std::vector<MapLine*> vpMapLines(nm);
    std::unordered_map<MapLine*, int> index;
    for (int j = 0; j < nm; ++j) {
        MapLine* p = new MapLine(zero6, w.anchor, &w.map);
        p->mbTrackInView = !skip[j];
        p->mnTrackScaleLevel = level[j]; p->mTrackViewCos = view_cos[j];
        p->mTrackProjX1 = proj[4 * j]; p->mTrackProjY1 = proj[4 * j + 1]; p->mTrackProjX2 = proj[4 * j + 2]; p->mTrackProjY2 = proj[4 * j + 3];
        p->mLDescriptor = cv::Mat(1, 32, CV_8U);
        std::memcpy(p->mLDescriptor.ptr(0), mdesc + 32 * (size_t)j, 32);
        p->nObs = m_has_obs[j] ? 1 : 0;
        vpMapLines[j] = p; index[p] = j; own.push_back(p);
    }
    DRV_LSDMATCHER matcher(nnratio);
    const int n = matcher.SearchByProjection(F, vpMapLines, th);
    for (int i = 0; i < nf; ++i) assigned[i] = (F.mvpMapLines[i] && F.mvpMapLines[i] != held[i]) ? index.at(F.mvpMapLines[i]) : -1;
    for (MapLine* p : own) delete p;
    return n;
}

// LSDmatcher::SearchByDescriptor(KeyFrame*, Frame&, vector<MapLine*>&)   src/LSDmatcher.cpp:242-279: kf_desc [n_kf][32] = the key frame's mLineDescriptors, kf_has_ml[i] = it
// holds a map line for key line i, f_desc [n_f][32] = the frame's mLdesc; match[j] = key line of the key frame whose map line is stored into vpMapLineMatches[j] (-1: NULL)
extern "C" int DRV(line_search_by_descriptor)(int n_kf, const uint8_t* kf_desc, const uint8_t* kf_has_ml, int n_f, const uint8_t* f_desc, int32_t* match) {
    World w;
    Vector6d zero6 = Vector6d::Zero();
    Frame K;
    K.N = 0; K.NL = n_kf;
    K.mvKeylinesUn.resize(n_kf); K.mvKeyLineFunctions.assign(n_kf, Eigen::Vector3d(0, 0, 1));
    K.mLdesc = cv::Mat(n_kf, 32, CV_8U);
    for (int i = 0; i < n_kf; ++i) std::memcpy(K.mLdesc.ptr(i), kf_desc + 32 * (size_t)i, 32);
    K.mvpMapLines.assign(n_kf, static_cast<MapLine*>(NULL));
    K.SetPose(cv::Mat::eye(4, 4, CV_32F));
    std::vector<MapLine*> own;
    std::unordered_map<MapLine*, int> index;
    for (int i = 0; i < n_kf; ++i)
        if (kf_has_ml[i]) { MapLine* p = new MapLine(zero6, w.anchor, &w.map); own.push_back(p); index[p] = i; K.mvpMapLines[i] = p; }
    KeyFrame* pKF = new KeyFrame(K, &w.map, static_cast<KeyFrameDatabase*>(NULL));
    Frame F;
    F.NL = n_f;
    F.mLdesc = cv::Mat(n_f, 32, CV_8U);
    for (int i = 0; i < n_f; ++i) std::memcpy(F.mLdesc.ptr(i), f_desc + 32 * (size_t)i, 32);
    DRV_LSDMATCHER matcher;
    std::vector<MapLine*> vpMapLineMatches;
    const int n = matcher.SearchByDescriptor(pKF, F, vpMapLineMatches);
    for (int j = 0; j < n_f; ++j) match[j] = vpMapLineMatches[j] ? index.at(vpMapLineMatches[j]) : -1;
    delete pKF;
    for (MapLine* p : own) delete p;
    return n;
}

// PlaneMatcher::SearchMapByCoefficients(Frame&, const vector<MapPlane*>&)   src/PlaneMatcher.cpp:10-82 (Frame::ComputePlaneWorldCoeff src/Frame.cc:815-820)
extern "C" int DRV(plane_match)(const float* Tcw, int n_frame, const float* frame_coef, int n_map, const float* map_coef, const uint8_t* map_bad, const int32_t* pts_off,
                               const float* pts, float dTh, float aTh, float verTh, float parTh, int32_t* match, int32_t* ver, int32_t* par) {
    World w;
    Frame F;
    F.SetPose(mat44(Tcw));
    F.mnPlaneNum = n_frame;
    for (int i = 0; i < n_frame; ++i) {
        cv::Mat c(4, 1, CV_32F);
        for (int k = 0; k < 4; ++k) c.at<float>(k) = frame_coef[4 * i + k];
        F.mvPlaneCoefficients.push_back(c);
    }
    F.mvpMapPlanes.assign(n_frame, static_cast<MapPlane*>(nullptr));
    F.mvpVerticalPlanes.assign(n_frame, static_cast<MapPlane*>(nullptr));
    F.mvpParallelPlanes.assign(n_frame, static_cast<MapPlane*>(nullptr));
    std::vector<MapPlane*> planes(n_map);
    std::unordered_map<MapPlane*, int> index;
    for (int j = 0; j < n_map; ++j) {
        cv::Mat c(4, 1, CV_32F);
        for (int k = 0; k < 4; ++k) c.at<float>(k) = map_coef[4 * j + k];
        MapPlane* p = new MapPlane(c, w.anchor, &w.map);
        p->mbBad = map_bad[j] != 0;
        for (int q = pts_off[j]; q < pts_off[j + 1]; ++q) { pcl::PointXYZRGB pt; pt.x = pts[3 * q]; pt.y = pts[3 * q + 1]; pt.z = pts[3 * q + 2]; p->mvPlanePoints->points.push_back(pt); }
        planes[j] = p; index[p] = j;
    }
    DRV_PLANEMATCHER matcher(dTh, aTh, verTh, parTh);
    const int n = matcher.SearchMapByCoefficients(F, planes);
    for (int i = 0; i < n_frame; ++i) {
        match[i] = F.mvpMapPlanes[i] ? index.at(F.mvpMapPlanes[i]) : -1;
        ver[i] = F.mvpVerticalPlanes[i] ? index.at(F.mvpVerticalPlanes[i]) : -1;
        par[i] = F.mvpParallelPlanes[i] ? index.at(F.mvpParallelPlanes[i]) : -1;
    }
    for (MapPlane* p : planes) delete p;
    return n;
}

// ---- the optimisers, called as they are ------------------------------------------------------------------------------------------------------------
// Optimizer::PoseOptimization(Frame*) src/Optimizer.cc:550-1275 and Optimizer::TranslationOptimization(Frame*) :2995-3737 run on a Frame built from a
// pslam_pose_problem: graph construction, the four optimise-and-classify rounds and the pose write-back are the reference's (oracle/ref/pose_driver.cc
// restates them; this entry point removes that caveat).  The Plane.* settings come through Config::Get from the settings table of the FileStorage
// stand-in.  A frame has one coefficient vector per plane slot; the problem's plane / parallel / vertical observations take consecutive slots.
static void set_plane_settings(double angle_info, double dist_info, double par_info, double ver_info, double plane_chi, double vp_chi) {
    auto& t = cv::FileStorage::table();
    t["Plane.AngleInfo"] = angle_info; t["Plane.DistanceInfo"] = dist_info; t["Plane.ParallelInfo"] = par_info; t["Plane.VerticalInfo"] = ver_info;
    t["Plane.Chi"] = plane_chi; t["Plane.VPChi"] = vp_chi;
    Config::SetParameterFile("pslam-settings-table");
}

extern "C" int DRV(full_pose_optimization)(const pslam_pose_problem* P, const float* Tcw_in, int translation_only, float* Tcw_out, uint8_t* o_pt, uint8_t* o_line,
                                          uint8_t* o_plane, uint8_t* o_par, uint8_t* o_ver) {
    set_plane_settings(P->angle_info, P->dist_info, P->par_info, P->ver_info, P->plane_chi, P->vp_chi);
    World w;
    Frame F;
    Frame::fx = P->fx; Frame::fy = P->fy; Frame::cx = P->cx; Frame::cy = P->cy; Frame::invfx = 1.0f / P->fx; Frame::invfy = 1.0f / P->fy;
    F.mbf = P->bf; F.mb = F.mbf / Frame::fx;
    F.N = P->n_points;
    F.mvKeysUn.resize(F.N); F.mvuRight.resize(F.N); F.mvInvLevelSigma2.resize(std::max(F.N, 1)); F.mvpMapPoints.resize(F.N); F.mvbOutlier.assign(F.N, false);
    const float zero3[3] = {0, 0, 0};
    for (int i = 0; i < F.N; ++i) {
        F.mvKeysUn[i].pt.x = P->obs[3 * i]; F.mvKeysUn[i].pt.y = P->obs[3 * i + 1]; F.mvKeysUn[i].octave = i;      // one "level" per point: mvInvLevelSigma2[octave]
        F.mvuRight[i] = P->obs[3 * i + 2];
        F.mvInvLevelSigma2[i] = P->inv_sigma2[i];
        F.mvpMapPoints[i] = w.make_point(P->Xw + 3 * i, zero3, 0, 0, nullptr, true);
        w.pts.push_back(F.mvpMapPoints[i]);
    }
    F.mvKeys = F.mvKeysUn;
    F.NL = P->n_lines;
    F.mvpMapLines.resize(F.NL); F.mvbLineOutlier.assign(F.NL, false); F.mvKeyLineFunctions.resize(F.NL); F.mvKeylinesUn.resize(F.NL);
    std::vector<MapLine*> lines;
    for (int i = 0; i < F.NL; ++i) {
        Vector6d X;
        for (int c = 0; c < 6; ++c) X(c) = P->line_Xw[6 * i + c];
        lines.push_back(new MapLine(X, w.anchor, &w.map));
        F.mvpMapLines[i] = lines.back();
        F.mvKeyLineFunctions[i] = Eigen::Vector3d(P->line_obs[3 * i], P->line_obs[3 * i + 1], P->line_obs[3 * i + 2]);
    }
    const int M = P->n_planes + P->n_par + P->n_ver;
    F.mnPlaneNum = M;
    F.mvpMapPlanes.assign(M, static_cast<MapPlane*>(nullptr)); F.mvpParallelPlanes.assign(M, static_cast<MapPlane*>(nullptr)); F.mvpVerticalPlanes.assign(M, static_cast<MapPlane*>(nullptr));
    F.mvbPlaneOutlier.assign(M, false); F.mvbParPlaneOutlier.assign(M, false); F.mvbVerPlaneOutlier.assign(M, false);
    std::vector<MapPlane*> planes;
    auto coef = [](const float* v) { cv::Mat m(4, 1, CV_32F); for (int k = 0; k < 4; ++k) m.at<float>(k) = v[k]; return m; };
    for (int i = 0; i < M; ++i) {
        const bool is_plane = i < P->n_planes, is_par = !is_plane && i < P->n_planes + P->n_par;
        const int j = is_plane ? i : is_par ? i - P->n_planes : i - P->n_planes - P->n_par;
        F.mvPlaneCoefficients.push_back(coef((is_plane ? P->plane_meas : is_par ? P->par_meas : P->ver_meas) + 4 * j));
        planes.push_back(new MapPlane(coef((is_plane ? P->plane_map : is_par ? P->par_map : P->ver_map) + 4 * j), w.anchor, &w.map));
        (is_plane ? F.mvpMapPlanes : is_par ? F.mvpParallelPlanes : F.mvpVerticalPlanes)[i] = planes.back();
    }
    F.SetPose(mat44(Tcw_in));
    const int n = translation_only ? DRV_OPTIMIZER::TranslationOptimization(&F) : DRV_OPTIMIZER::PoseOptimization(&F);
    for (int i = 0; i < 16; ++i) Tcw_out[i] = F.mTcw.at<float>(i / 4, i % 4);
    for (int i = 0; i < F.N; ++i) o_pt[i] = F.mvbOutlier[i] ? 1 : 0;
    for (int i = 0; i < F.NL; ++i) o_line[i] = F.mvbLineOutlier[i] ? 1 : 0;
    for (int i = 0; i < P->n_planes; ++i) o_plane[i] = F.mvbPlaneOutlier[i] ? 1 : 0;
    for (int i = 0; i < P->n_par; ++i) o_par[i] = F.mvbParPlaneOutlier[P->n_planes + i] ? 1 : 0;
    for (int i = 0; i < P->n_ver; ++i) o_ver[i] = F.mvbVerPlaneOutlier[P->n_planes + P->n_par + i] ? 1 : 0;
    for (MapLine* l : lines) delete l;
    for (MapPlane* q : planes) delete q;
    return n;
}

// Optimizer::LocalBundleAdjustment(KeyFrame*, bool*, Map*) src/Optimizer.cc:1853-2678 called as it is on a KeyFrame / MapPoint / MapLine / MapPlane graph
// built from a pslam_lba_problem.  The current key frame is the last one; its covisible list holds every non-fixed key frame plus key frame 0 (which the
// reference fixes through mnId == 0); the other fixed key frames are left for the function to discover through the observations.  Observation j of a
// family sits in feature slot j of its key frame.  Read back: the poses / positions the function wrote (float, Converter::toCvMat), which slots it cleared
// (erase_*: the observation was erased, or its landmark went bad - *_bad tells which) - everything in between is the reference's code.
extern "C" int DRV(full_local_bundle_adjustment)(const pslam_lba_problem* P, pslam_lba_result* R, uint8_t* pt_bad, uint8_t* line_bad, uint8_t* plane_bad) {
    set_plane_settings(P->angle_info, P->dist_info, 0, 0, P->plane_chi, P->vp_chi);
    World w;
    w.anchor->mnId = 1000000;
    const int NP = std::max(P->n_pt_obs, 1), NLn = std::max(P->n_line_obs, 1), off1 = P->n_plane_obs[0], off2 = off1 + P->n_plane_obs[1], NPl = std::max(off2 + P->n_plane_obs[2], 1);
    const int cur = P->n_kf - 1;
    std::vector<KeyFrame*> kfs(P->n_kf);
    for (int k = 0; k < P->n_kf; ++k) {
        Frame F;
        const float* K = P->kf_K + 5 * k;
        Frame::fx = K[0]; Frame::fy = K[1]; Frame::cx = K[2]; Frame::cy = K[3]; Frame::invfx = 1.0f / K[0]; Frame::invfy = 1.0f / K[1];
        F.mbf = K[4]; F.mb = F.mbf / Frame::fx;
        F.N = NP;
        F.mvKeysUn.resize(NP); F.mvuRight.assign(NP, -1.f); F.mvDepth.assign(NP, -1.f); F.mvInvLevelSigma2.assign(std::max(NP, NLn), 1.f);
        F.mvLevelSigma2.assign(std::max(NP, NLn), 1.f); F.mvScaleFactors.assign(std::max(NP, NLn), 1.f); F.mnScaleLevels = std::max(NP, NLn);
        for (int j = 0; j < P->n_pt_obs; ++j) {
            F.mvKeysUn[j].octave = j;
            if (P->pt_obs_kf[j] != k) continue;
            F.mvKeysUn[j].pt.x = P->pt_obs_uvr[3 * j]; F.mvKeysUn[j].pt.y = P->pt_obs_uvr[3 * j + 1]; F.mvuRight[j] = P->pt_obs_uvr[3 * j + 2];
            F.mvInvLevelSigma2[j] = P->pt_obs_inv_sigma2[j];
        }
        F.mvKeys = F.mvKeysUn;
        F.mvpMapPoints.assign(NP, static_cast<MapPoint*>(NULL));
        F.NL = NLn;
        F.mvKeylinesUn.resize(NLn); F.mvKeyLineFunctions.assign(NLn, Eigen::Vector3d(0, 0, 1)); F.mvpMapLines.assign(NLn, static_cast<MapLine*>(NULL));
        for (int j = 0; j < P->n_line_obs; ++j) F.mvKeyLineFunctions[j] = Eigen::Vector3d(P->line_obs_l[3 * j], P->line_obs_l[3 * j + 1], P->line_obs_l[3 * j + 2]);
        F.mnPlaneNum = NPl;
        for (int q = 0; q < NPl; ++q) { cv::Mat c = cv::Mat::zeros(4, 1, CV_32F); c.at<float>(3) = 1.f; F.mvPlaneCoefficients.push_back(c); }
        for (int t = 0; t < 3; ++t)
            for (int j = 0; j < P->n_plane_obs[t]; ++j)
                if (P->plane_obs_kf[t][j] == k) for (int c = 0; c < 4; ++c) F.mvPlaneCoefficients[(t == 0 ? 0 : t == 1 ? off1 : off2) + j].at<float>(c) = P->plane_obs_meas[t][4 * j + c];
        F.mvPlanePoints.resize(NPl);
        F.mvpMapPlanes.assign(NPl, static_cast<MapPlane*>(nullptr)); F.mvpParallelPlanes.assign(NPl, static_cast<MapPlane*>(nullptr)); F.mvpVerticalPlanes.assign(NPl, static_cast<MapPlane*>(nullptr));
        F.mTcw = mat44(P->kf_Tcw + 16 * k);
        F.SetPose(mat44(P->kf_Tcw + 16 * k));
        kfs[k] = new KeyFrame(F, &w.map, static_cast<KeyFrameDatabase*>(NULL));
        kfs[k]->mnId = k;
    }
    for (int k = 0; k < P->n_kf; ++k)
        if (k != cur && (!P->kf_fixed[k] || k == 0)) kfs[cur]->mvpOrderedConnectedKeyFrames.push_back(kfs[k]);
    const float zero3[3] = {0, 0, 0};
    std::vector<MapPoint*> pts(P->n_points);
    for (int i = 0; i < P->n_points; ++i) { pts[i] = w.make_point(P->pt_Xw + 3 * i, zero3, 0, 0, nullptr, false); pts[i]->mnId = i; pts[i]->mpRefKF = nullptr; }
    for (int j = 0; j < P->n_pt_obs; ++j) {
        KeyFrame* kf = kfs[P->pt_obs_kf[j]]; MapPoint* p = pts[P->pt_obs_pt[j]];
        if (p->IsInKeyFrame(kf)) return -2;                                  // the object graph holds one observation per (point, key frame)
        kf->mvpMapPoints[j] = p; p->AddObservation(kf, j);
        if (!p->mpRefKF) p->mpRefKF = kf;
    }
    for (MapPoint* p : pts) if (!p->mpRefKF) p->mpRefKF = kfs[cur];
    std::vector<MapLine*> lines(P->n_lines);
    for (int i = 0; i < P->n_lines; ++i) {
        Vector6d X;
        for (int c = 0; c < 6; ++c) X(c) = P->line_Xw[6 * i + c];
        lines[i] = new MapLine(X, kfs[cur], &w.map);
        lines[i]->mnId = i; lines[i]->mpRefKF = nullptr;
    }
    std::vector<KeyFrame*> line_obs_holder(P->n_line_obs);
    for (int j = 0; j < P->n_line_obs; ++j) {                               // every line edge hangs on the CURRENT key frame (the reference's quirk); the observing
        MapLine* l = lines[P->line_obs_line[j]];                            // key frame only has to be distinct per observation of a line: take them in order
        int k = 0;
        while (k < P->n_kf && (l->mObservations.count(kfs[(cur + P->n_kf - k) % P->n_kf]))) ++k;
        if (k == P->n_kf) return -3;
        KeyFrame* kf = kfs[(cur + P->n_kf - k) % P->n_kf];
        kf->mvpMapLines[j] = l; l->AddObservation(kf, j); line_obs_holder[j] = kf;
        if (!l->mpRefKF) l->mpRefKF = kf;
    }
    for (MapLine* l : lines) if (!l->mpRefKF) l->mpRefKF = kfs[cur];
    std::vector<MapPlane*> planes(P->n_planes);
    for (int i = 0; i < P->n_planes; ++i) {
        cv::Mat c(4, 1, CV_32F);
        for (int q = 0; q < 4; ++q) c.at<float>(q) = P->plane_Xw[4 * i + q];
        planes[i] = new MapPlane(c, kfs[cur], &w.map);
        planes[i]->mnId = i;
    }
    for (int t = 0; t < 3; ++t)
        for (int j = 0; j < P->n_plane_obs[t]; ++j) {
            KeyFrame* kf = kfs[P->plane_obs_kf[t][j]]; MapPlane* q = planes[P->plane_obs_plane[t][j]];
            const int slot = (t == 0 ? 0 : t == 1 ? off1 : off2) + j;
            if (t == 0) { kf->mvpMapPlanes[slot] = q; q->AddObservation(kf, slot); }
            else if (t == 1) { kf->mvpVerticalPlanes[slot] = q; q->AddVerObservation(kf, slot); }
            else { kf->mvpParallelPlanes[slot] = q; q->AddParObservation(kf, slot); }
        }
    bool stop = false;
    DRV_OPTIMIZER::LocalBundleAdjustment(kfs[cur], &stop, &w.map);
    for (int k = 0; k < P->n_kf; ++k) {
        cv::Mat T = kfs[k]->GetPose();
        for (int i = 0; i < 16; ++i) { R->kf_Tcw[16 * k + i] = T.at<float>(i / 4, i % 4); R->kf_Tcw_d[16 * k + i] = T.at<float>(i / 4, i % 4); }
    }
    for (int i = 0; i < P->n_points; ++i) {
        cv::Mat X = pts[i]->GetWorldPos();
        for (int c = 0; c < 3; ++c) { R->pt_Xw[3 * i + c] = X.at<float>(c); R->pt_Xw_d[3 * i + c] = X.at<float>(c); }
        pt_bad[i] = pts[i]->isBad() ? 1 : 0;
    }
    for (int i = 0; i < P->n_lines; ++i) {
        Vector6d X = lines[i]->GetWorldPos();
        for (int c = 0; c < 6; ++c) { R->line_Xw[6 * i + c] = X(c); R->line_Xw_d[6 * i + c] = X(c); }
        line_bad[i] = lines[i]->isBad() ? 1 : 0;
    }
    for (int i = 0; i < P->n_planes; ++i) {
        cv::Mat X = planes[i]->GetWorldPos();
        for (int c = 0; c < 4; ++c) { R->plane_Xw[4 * i + c] = X.at<float>(c); R->plane_Xw_d[4 * i + c] = X.at<float>(c); }
        plane_bad[i] = planes[i]->isBad() ? 1 : 0;
    }
    for (int j = 0; j < P->n_pt_obs; ++j) R->erase_pt[j] = kfs[P->pt_obs_kf[j]]->mvpMapPoints[j] ? 0 : 1;
    for (int j = 0; j < P->n_line_obs; ++j) R->erase_line[j] = line_obs_holder[j]->mvpMapLines[j] ? 0 : 1;
    for (int j = 0; j < P->n_plane_obs[0]; ++j) R->erase_plane[0][j] = kfs[P->plane_obs_kf[0][j]]->mvpMapPlanes[j] ? 0 : 1;
    for (int j = 0; j < P->n_plane_obs[1]; ++j) R->erase_plane[1][j] = kfs[P->plane_obs_kf[1][j]]->mvpVerticalPlanes[off1 + j] ? 0 : 1;
    for (int j = 0; j < P->n_plane_obs[2]; ++j) R->erase_plane[2][j] = kfs[P->plane_obs_kf[2][j]]->mvpParallelPlanes[off2 + j] ? 0 : 1;
    R->iterations[0] = R->iterations[1] = -1;
    for (MapPlane* q : planes) delete q;
    for (MapLine* l : lines) delete l;
    for (MapPoint* q : pts) delete q;
    for (KeyFrame* k : kfs) delete k;
    return 0;
}

#ifndef PSLAM_ADAPTER_BUILD
// Frame::ComputeStereoFromRGBD(imDepth) src/Frame.cc:603-621 and Frame::isLineGood(imGray, imDepth, K) src/Frame.cc:189-267 called as they are (with the
// reference's src/LineExtractor.cpp and libc rand()) on a Frame holding the key points / key lines.
extern "C" void ref_full_compute_stereo_from_rgbd(int n, const float* keys_xy, const float* keys_un_xy, const float* depth, int w, int h, float bf, float* u_right, float* z) {
    Frame F;
    F.N = n; F.mbf = bf;
    F.mvKeys.resize(n); F.mvKeysUn.resize(n);
    for (int i = 0; i < n; ++i) { F.mvKeys[i].pt.x = keys_xy[2 * i]; F.mvKeys[i].pt.y = keys_xy[2 * i + 1]; F.mvKeysUn[i].pt.x = keys_un_xy[2 * i]; F.mvKeysUn[i].pt.y = keys_un_xy[2 * i + 1]; }
    cv::Mat imDepth(h, w, CV_32F, (void*)depth);
    F.ComputeStereoFromRGBD(imDepth);
    for (int i = 0; i < n; ++i) { u_right[i] = F.mvuRight[i]; z[i] = F.mvDepth[i]; }
}

extern "C" void ref_full_lines3d_frame(const void* keylines, int n_lines, const float* depth, int w, int h, const float* cam, uint32_t seed, int skip, float* depth_line,
                                       double* lines3d) {
    static_assert(sizeof(cv::line_descriptor::KeyLine) == 68, "KeyLine layout");
    const cv::line_descriptor::KeyLine* kl = (const cv::line_descriptor::KeyLine*)keylines;
    Frame F;
    Frame::fx = cam[0]; Frame::fy = cam[1]; Frame::cx = cam[2]; Frame::cy = cam[3]; Frame::invfx = 1.0f / cam[0]; Frame::invfy = 1.0f / cam[1];
    F.mvKeylinesUn.assign(kl, kl + n_lines);
    F.NL = n_lines;
    cv::Mat K = (cv::Mat_<double>(3, 3) << cam[0], 0, cam[2], 0, cam[1], cam[3], 0, 0, 1);       // tmpK, src/Frame.cc:84-86
    cv::Mat imDepth(h, w, CV_32F, (void*)depth), imGray;
    srand(seed);
    for (int i = 0; i < skip; ++i) (void)rand();
    F.isLineGood(imGray, imDepth, K);
    for (int i = 0; i < n_lines; ++i) { depth_line[i] = F.mvDepthLine[i]; for (int c = 0; c < 6; ++c) lines3d[6 * i + c] = F.mvLines3D[i](c); }
}
#endif  // !PSLAM_ADAPTER_BUILD
