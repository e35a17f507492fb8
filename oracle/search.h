// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Projection-guided descriptor search on the tracking path, over plain-array "views" of a Frame and of the map points:
//   Frame::AssignFeaturesToGrid / PosInGrid      src/Frame.cc:155-168, 526-535   (64 x 48 grid, round() binning)
//   Frame::GetFeaturesInArea                     src/Frame.cc:440-489
//   Frame::isInFrustum(MapPoint*, cosLimit)      src/Frame.cc:312-367 ; MapPoint::PredictScale src/MapPoint.cc:419-434
//   ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)          src/ORBmatcher.cc:46-130
//   ORBmatcher::SearchByProjection(Frame& cur, const Frame& last, th, bMono)      src/ORBmatcher.cc:1396-1535
//   ORBmatcher::ComputeThreeMaxima               src/ORBmatcher.cc:1666-1707
// Arithmetic conventions (the reference leaves them to OpenCV / libm): cv::Mat float products and cv::norm
// accumulate in double and round once to float; float expressions are evaluated left to right without FMA;
// log(ratio) of PredictScale is evaluated in double and rounded to float.
// Pinned: identical match lists / mbTrackInView flags to the reference's own src/ORBmatcher.cc + src/Frame.cc + src/MapPoint.cc compiled here
// (oracle/ref/match_driver.cc -> oracle/_ref/libmatch_ref.so, tests/test_oracle_match_ref.py, tests/golden/match_reference.npz).
#pragma once
#include <cstdint>
#include <vector>

#include "orb.h"

namespace oracle {

struct FrameView {
    int n = 0;
    const KeyPoint* keys_un = nullptr;     // undistorted keypoints (pt, octave, angle)
    const float* u_right = nullptr;        // mvuRight
    const uint8_t* desc = nullptr;         // [n][32]
    float Tcw[16];                         // row-major float 4x4
    float fx, fy, cx, cy, bf;
    float min_x, max_x, min_y, max_y;      // mnMinX ...
    int n_levels = 8;
    const float* scale_factors = nullptr;  // mvScaleFactors
    float log_scale_factor = 0;            // mfLogScaleFactor
};

struct MapPointsView {
    int n = 0;
    const float* pos = nullptr;            // [n][3] GetWorldPos
    const float* normal = nullptr;         // [n][3] GetNormal
    const float* max_distance = nullptr;   // [n] mfMaxDistance (GetMaxDistanceInvariance = 1.2f * it)
    const float* min_distance = nullptr;   // [n] mfMinDistance (GetMinDistanceInvariance = 0.8f * it)
    const uint8_t* desc = nullptr;         // [n][32] GetDescriptor
    const uint8_t* skip = nullptr;         // [n] mnLastFrameSeen == frame id, or isBad()
    const uint8_t* has_obs = nullptr;      // [n] Observations() > 0
};

// matches[i] = map point index held by keypoint i (mvpMapPoints), -1 = none; updated in place. in_view receives
// mbTrackInView of every map point. Returns nmatches.
int search_by_projection_map(const FrameView& F, const MapPointsView& M, float th, float nnratio, int32_t* matches, uint8_t* in_view);

struct LastFrameView {
    int n = 0;
    const KeyPoint* keys = nullptr;        // mvKeys (octave) — same angle as keys_un
    const int32_t* map_point = nullptr;    // [n] index into the map point arrays or -1
    const uint8_t* outlier = nullptr;      // [n] mvbOutlier
    float Tcw[16];
};
// Motion-model search. matches as above. Returns nmatches after the rotation-consistency filter.
int search_by_projection_last(const FrameView& cur, const LastFrameView& last, const MapPointsView& M, float th, bool mono, bool check_orientation,
                              int32_t* matches);

}  // namespace oracle
