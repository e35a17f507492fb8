// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// PlaneMatcher::SearchMapByCoefficients src/PlaneMatcher.cpp:10-67, PointDistanceFromPlane :68-80,
// Frame::ComputePlaneWorldCoeff src/Frame.cc:815-820 (pi_w = Tcw^T pi_c as a float cv::Mat product: double accumulation).
// Float expressions are evaluated left to right without FMA.
#include <cmath>
#include <cstdint>

extern "C" int orc_plane_match(const float* Tcw, int n_frame, const float* frame_coef /*[n][4]*/, int n_map, const float* map_coef /*[m][4]*/,
                               const uint8_t* map_bad, const int32_t* pts_off /*[m+1]*/, const float* pts /*[..][3]*/, float dTh, float aTh,
                               float verTh, float parTh, int32_t* match, int32_t* ver, int32_t* par) {
    int nmatches = 0;
    for (int i = 0; i < n_frame; ++i) {
        float pM[4];
        for (int r = 0; r < 4; ++r) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += (double)Tcw[k * 4 + r] * (double)frame_coef[4 * i + k];     // (Tcw^T)[r][k] = Tcw[k][r]
            pM[r] = (float)s;
        }
        float ldTh = dTh, lverTh = verTh, lparTh = parTh;
        bool found = false;
        match[i] = ver[i] = par[i] = -1;
        for (int j = 0; j < n_map; ++j) {
            if (map_bad[j]) continue;
            const float* pW = map_coef + 4 * j;
            const float angle = pM[0] * pW[0] + pM[1] * pW[1] + pM[2] * pW[2];
            if (angle > aTh || angle < -aTh) {
                double res = 100;
                for (int p = pts_off[j]; p < pts_off[j + 1]; ++p) {
                    const double dis = std::fabs(pM[0] * pts[3 * p] + pM[1] * pts[3 * p + 1] + pM[2] * pts[3 * p + 2] + pM[3]);   // float expression
                    if (dis < res) res = dis;
                }
                if (res < ldTh) { ldTh = (float)res; match[i] = j; found = true; continue; }
            }
            if (angle < lverTh && angle > -lverTh) { lverTh = std::fabs(angle); ver[i] = j; continue; }
            if (angle > lparTh || angle < -lparTh) { lparTh = std::fabs(angle); par[i] = j; }
        }
        if (found) ++nmatches;
    }
    return nmatches;
}
