// TEST INFRASTRUCTURE ONLY — see linesearch.h.
#include "linesearch.h"

#include <cmath>
#include <vector>

#include "match.h"

namespace oracle {

int line_search_by_projection(const LineFrameView& F, const MapLinesView& M, float th, float nnratio, int32_t* assigned) {
    int nmatches = 0;
    const int TH_HIGH = 100;
    const bool bFactor = th != 1.0;
    std::vector<uint8_t> occupied(F.has_obs, F.has_obs + F.n);        // F.mvpMapLines[idx] && ->Observations() > 0
    for (int i = 0; i < F.n; ++i) assigned[i] = -1;
    for (int m = 0; m < M.n; ++m) {
        if (M.skip[m]) continue;
        const int nPredictLevel = M.level[m];
        float r = M.view_cos[m] > 0.998 ? 5.0f : 8.0f;               // RadiusByViewingCos
        if (bFactor) r *= th;
        // GetLinesInArea(x1, y1, x2, y2, r * scale[level], level - 1, level)
        const float x1 = M.proj[4 * m], y1 = M.proj[4 * m + 1], x2 = M.proj[4 * m + 2], y2 = M.proj[4 * m + 3];
        const float rr = r * F.scale_factors[std::min(std::max(nPredictLevel, 0), F.n_levels - 1)];   // the reference reads past mvScaleFactors here (level not clamped by MapLine::PredictScale): index clamped, level gate keeps the raw level
        const int minLevel = nPredictLevel - 1, maxLevel = nPredictLevel;
        const bool bCheckLevels = (minLevel > 0) || (maxLevel > 0);
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        bool any = false;
        for (int i = 0; i < F.n; ++i) {
            const float distance = (0.5 * (x1 + x2) - F.pt[2 * i]) * (0.5 * (x1 + x2) - F.pt[2 * i]) +
                                   (0.5 * (y1 + y2) - F.pt[2 * i + 1]) * (0.5 * (y1 + y2) - F.pt[2 * i + 1]);
            if (distance > rr * rr) continue;
            const float slope = (y1 - y2) / (x1 - x2) - F.angle[i];
            if (slope > rr * 0.01) continue;
            if (bCheckLevels) {
                if (F.octave[i] < minLevel) continue;
                if (maxLevel >= 0 && F.octave[i] > maxLevel) continue;
            }
            any = true;
            if (occupied[i]) continue;
            const int dist = descriptor_distance(M.desc + 32 * m, F.desc + 32 * i);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.octave[i]; bestIdx = i; }
            else if (dist < bestDist2) { bestLevel2 = F.octave[i]; bestDist2 = dist; }
        }
        if (!any) continue;
        if (bestDist <= TH_HIGH) {
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            assigned[bestIdx] = m;
            occupied[bestIdx] = M.has_obs[m];
            nmatches++;
        }
    }
    return nmatches;
}

}  // namespace oracle

namespace oracle {
namespace {
inline void rt_f(const float* T, const float* P, float out[3]) {
    for (int r = 0; r < 3; ++r) {
        const float m = (float)((double)T[4 * r] * P[0] + (double)T[4 * r + 1] * P[1] + (double)T[4 * r + 2] * P[2]);
        out[r] = m + T[4 * r + 3];
    }
}
}  // namespace

void lines_in_frustum(const LineFrustumFrame& F, int n, const double* pos, const double* normal, const float* max_distance, const float* min_distance,
                      float cos_limit, uint8_t* in_view, float* proj, int32_t* level, float* view_cos) {
    float Ow[3];                                   // mOw = -mRcw.t() * mtcw
    for (int r = 0; r < 3; ++r) {
        const float a = -F.Tcw[0 + r], b = -F.Tcw[4 + r], c = -F.Tcw[8 + r];
        Ow[r] = (float)((double)a * F.Tcw[3] + (double)b * F.Tcw[7] + (double)c * F.Tcw[11]);
    }
    for (int k = 0; k < n; ++k) {
        in_view[k] = 0; level[k] = 0; view_cos[k] = 0;
        for (int q = 0; q < 4; ++q) proj[4 * k + q] = 0;
        const float SP[3] = {(float)pos[6 * k], (float)pos[6 * k + 1], (float)pos[6 * k + 2]};
        const float EP[3] = {(float)pos[6 * k + 3], (float)pos[6 * k + 4], (float)pos[6 * k + 5]};
        float SPc[3], EPc[3];
        rt_f(F.Tcw, SP, SPc);
        rt_f(F.Tcw, EP, EPc);
        if (SPc[2] < 0.0f || EPc[2] < 0.0f) continue;
        const float invz1 = 1.0f / SPc[2];
        const float u1 = F.fx * SPc[0] * invz1 + F.cx, v1 = F.fy * SPc[1] * invz1 + F.cy;
        if (u1 < F.min_x || u1 > F.max_x) continue;
        if (v1 < F.min_y || v1 > F.max_y) continue;
        const float invz2 = 1.0f / EPc[2];
        const float u2 = F.fx * EPc[0] * invz2 + F.cx, v2 = F.fy * EPc[1] * invz2 + F.cy;
        if (u2 < F.min_x || u2 > F.max_x) continue;
        if (v2 < F.min_y || v2 > F.max_y) continue;
        const float maxD = 1.2f * max_distance[k], minD = 0.8f * min_distance[k];
        float OM[3];
        for (int c = 0; c < 3; ++c) OM[c] = 0.5f * (SP[c] + EP[c]) - Ow[c];
        const float dist = (float)std::sqrt((double)OM[0] * OM[0] + (double)OM[1] * OM[1] + (double)OM[2] * OM[2]);
        if (dist < minD || dist > maxD) continue;
        const float pn[3] = {(float)normal[3 * k], (float)normal[3 * k + 1], (float)normal[3 * k + 2]};
        const float vc = (float)(((double)OM[0] * pn[0] + (double)OM[1] * pn[1] + (double)OM[2] * pn[2]) / dist);
        if (vc < cos_limit) continue;
        const float ratio = max_distance[k] / dist;
        level[k] = (int)std::ceil((float)std::log((double)ratio) / F.log_scale_factor);       // no clamping in MapLine::PredictScale
        in_view[k] = 1; view_cos[k] = vc;
        proj[4 * k] = u1; proj[4 * k + 1] = v1; proj[4 * k + 2] = u2; proj[4 * k + 3] = v2;
    }
}
}  // namespace oracle
