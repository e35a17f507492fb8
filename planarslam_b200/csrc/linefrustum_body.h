// Frame::isInFrustum(MapLine*, viewingCosLimit) (src/Frame.cc:369-437) with MapLine::PredictScale (src/MapLine.cpp:381-390) and
// Get{Min,Max}DistanceInvariance (:370-379): one map line per thread.  Host/device shared source (plain float / double arithmetic in
// the order cv::Mat CV_32F operations round: double accumulation inside a product, norm or dot, one float rounding per Mat operation),
// so the CPU suite can check it against the oracle (tests/test_linefrustum_host.py).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define LF_HD __host__ __device__
#else
#define LF_HD
#endif

struct LfFrame { float Tcw[16]; float Ow[3]; float fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor; };

LF_HD inline void lf_camera_center(LfFrame& F) {                  // mOw = -mRcw.t() * mtcw
    for (int r = 0; r < 3; ++r) {
        const float a = -F.Tcw[0 + r], b = -F.Tcw[4 + r], c = -F.Tcw[8 + r];
        F.Ow[r] = (float)((double)a * F.Tcw[3] + (double)b * F.Tcw[7] + (double)c * F.Tcw[11]);
    }
}
LF_HD inline void lf_rt(const float* T, const float* P, float* out) {
    for (int r = 0; r < 3; ++r) {
        const float m = (float)((double)T[4 * r] * P[0] + (double)T[4 * r + 1] * P[1] + (double)T[4 * r + 2] * P[2]);
        out[r] = m + T[4 * r + 3];
    }
}
// pos: 6 doubles (start, end), normal: 3 doubles.  Returns mbTrackInView; proj = mTrackProjX1, Y1, X2, Y2.
LF_HD inline bool lf_line_in_frustum(const LfFrame& F, const double* pos, const double* normal, float max_distance, float min_distance, float cos_limit, float* proj,
                                     int32_t& level, float& view_cos) {
    level = 0; view_cos = 0;
    for (int q = 0; q < 4; ++q) proj[q] = 0;
    const float SP[3] = {(float)pos[0], (float)pos[1], (float)pos[2]}, EP[3] = {(float)pos[3], (float)pos[4], (float)pos[5]};
    float SPc[3], EPc[3];
    lf_rt(F.Tcw, SP, SPc);
    lf_rt(F.Tcw, EP, EPc);
    if (SPc[2] < 0.0f || EPc[2] < 0.0f) return false;
    const float invz1 = 1.0f / SPc[2];
    const float u1 = F.fx * SPc[0] * invz1 + F.cx, v1 = F.fy * SPc[1] * invz1 + F.cy;
    if (u1 < F.min_x || u1 > F.max_x) return false;
    if (v1 < F.min_y || v1 > F.max_y) return false;
    const float invz2 = 1.0f / EPc[2];
    const float u2 = F.fx * EPc[0] * invz2 + F.cx, v2 = F.fy * EPc[1] * invz2 + F.cy;
    if (u2 < F.min_x || u2 > F.max_x) return false;
    if (v2 < F.min_y || v2 > F.max_y) return false;
    const float maxD = 1.2f * max_distance, minD = 0.8f * min_distance;
    float OM[3];
    for (int c = 0; c < 3; ++c) OM[c] = 0.5f * (SP[c] + EP[c]) - F.Ow[c];
    const float dist = (float)sqrt((double)OM[0] * OM[0] + (double)OM[1] * OM[1] + (double)OM[2] * OM[2]);
    if (dist < minD || dist > maxD) return false;
    const float pn[3] = {(float)normal[0], (float)normal[1], (float)normal[2]};
    const float vc = (float)(((double)OM[0] * pn[0] + (double)OM[1] * pn[1] + (double)OM[2] * pn[2]) / dist);
    if (vc < cos_limit) return false;
    const float ratio = max_distance / dist;
    level = (int32_t)ceilf((float)log((double)ratio) / F.log_scale_factor);           // MapLine::PredictScale does not clamp
    view_cos = vc;
    proj[0] = u1; proj[1] = v1; proj[2] = u2; proj[3] = v2;
    return true;
}
