// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of LSDmatcher::SearchByProjection(Frame&, const vector<MapLine*>&, th)  src/LSDmatcher.cpp:141-211 with
// Frame::GetLinesInArea src/Frame.cc:491-523 (mid-point distance gate, the one-sided slope-minus-angle gate, octave gate)
// and LSDmatcher::RadiusByViewingCos :369-375.  The map-line fields are those Frame::isInFrustum(MapLine*) fills
// (mbTrackInView, mTrackProjX1..Y2, mnTrackScaleLevel, mTrackViewCos).  Pinned: identical to the reference's own
// src/LSDmatcher.cpp + src/Frame.cc compiled here (oracle/ref/match_driver.cc -> oracle/_ref/libmatch_ref.so, tests/test_oracle_match_ref.py).
#pragma once
#include <cstdint>

namespace oracle {

struct LineFrameView {            // what the function reads from the Frame
    int n = 0;
    const float* pt = nullptr;     // [n][2] KeyLine::pt
    const float* angle = nullptr;  // [n]
    const int32_t* octave = nullptr;
    const uint8_t* desc = nullptr; // [n][32] mLdesc rows
    const uint8_t* has_obs = nullptr;   // [n] mvpMapLines[i] && mvpMapLines[i]->Observations() > 0 on entry
    const float* scale_factors = nullptr; int n_levels = 0;
};
struct MapLinesView {
    int n = 0;
    const uint8_t* skip = nullptr;      // null / isBad() / !mbTrackInView
    const int32_t* level = nullptr;     // mnTrackScaleLevel
    const float* view_cos = nullptr;    // mTrackViewCos
    const float* proj = nullptr;        // [n][4] mTrackProjX1, Y1, X2, Y2
    const uint8_t* desc = nullptr;      // [n][32]
    const uint8_t* has_obs = nullptr;   // Observations() > 0
};
// assigned[i] = map line written into F.mvpMapLines[i] by this call (-1: untouched); returns nmatches
int line_search_by_projection(const LineFrameView& F, const MapLinesView& M, float th, float nnratio, int32_t* assigned);

}  // namespace oracle

namespace oracle {
// Frame::isInFrustum(MapLine*, viewingCosLimit)  src/Frame.cc:369-437 with MapLine::PredictScale src/MapLine.cpp:381-390 and
// Get{Min,Max}DistanceInvariance :370-379 - the pass of Tracking::SearchLocalLines (src/Tracking.cc:2352-2366) that fills the
// MapLinesView fields above.  Float arithmetic as cv::Mat CV_32F does it (double accumulation inside a product / norm / dot, one
// rounding per Mat operation), same conventions as the map-point version in search.cc.
struct LineFrustumFrame { float Tcw[16]; float fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor; };
// pos [n][6] (GetWorldPos: start, end), normal [n][3] (GetNormal) as double (Vector6d / Vector3d); max_distance / min_distance: mfMaxDistance / mfMinDistance
void lines_in_frustum(const LineFrustumFrame& F, int n, const double* pos, const double* normal, const float* max_distance, const float* min_distance,
                      float cos_limit, uint8_t* in_view, float* proj /* [n][4] */, int32_t* level, float* view_cos);
}  // namespace oracle
