// TEST INFRASTRUCTURE ONLY — see linesearch.h.
#include "linesearch.h"

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
        const float rr = r * F.scale_factors[nPredictLevel];
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
