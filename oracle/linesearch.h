// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of LSDmatcher::SearchByProjection(Frame&, const vector<MapLine*>&, th)  src/LSDmatcher.cpp:141-211 with
// Frame::GetLinesInArea src/Frame.cc:491-523 (mid-point distance gate, the one-sided slope-minus-angle gate, octave gate)
// and LSDmatcher::RadiusByViewingCos :369-375.  The map-line fields are those Frame::isInFrustum(MapLine*) fills
// (mbTrackInView, mTrackProjX1..Y2, mnTrackScaleLevel, mTrackViewCos).  "parity unpinned": the reference ships no tests.
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
