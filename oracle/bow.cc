// TEST INFRASTRUCTURE ONLY — see bow.h.
#include "bow.h"

#include <algorithm>
#include <cmath>
#include <vector>

#include "match.h"

namespace oracle {

int search_by_bow(const BowSide& KF, const uint8_t* kf_has_mp, const BowSide& F, float nnratio, bool check_orientation, int32_t* match) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    for (int i = 0; i < F.n; ++i) match[i] = -1;
    int nmatches = 0;
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    int a = 0, b = 0;
    while (a < KF.n_nodes && b < F.n_nodes) {
        if (KF.node_id[a] == F.node_id[b]) {
            for (int q = KF.node_off[a]; q < KF.node_off[a + 1]; ++q) {
                const int realIdxKF = KF.node_feat[q];
                if (!kf_has_mp[realIdxKF]) continue;
                int bestDist1 = 256, bestIdxF = -1, bestDist2 = 256;
                for (int p = F.node_off[b]; p < F.node_off[b + 1]; ++p) {
                    const int realIdxF = F.node_feat[p];
                    if (match[realIdxF] >= 0) continue;
                    const int dist = descriptor_distance(KF.desc + 32 * realIdxKF, F.desc + 32 * realIdxF);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdxF = realIdxF; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 <= TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    match[bestIdxF] = realIdxKF;
                    if (check_orientation) {
                        float rot = KF.angle[realIdxKF] - F.angle[bestIdxF];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(bestIdxF);
                    }
                    nmatches++;
                }
            }
            ++a; ++b;
        } else if (KF.node_id[a] < F.node_id[b]) {
            a = (int)(std::lower_bound(KF.node_id, KF.node_id + KF.n_nodes, F.node_id[b]) - KF.node_id);
        } else {
            b = (int)(std::lower_bound(F.node_id, F.node_id + F.n_nodes, KF.node_id[a]) - F.node_id);
        }
    }
    if (check_orientation) {
        // ComputeThreeMaxima (src/ORBmatcher.cc:1666-1707)
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace oracle
