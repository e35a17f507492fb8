// TEST INFRASTRUCTURE ONLY — see search.h.
#include "search.h"

#include <cmath>

#include "match.h"

namespace oracle {
namespace {

const int GRID_COLS = 64, GRID_ROWS = 48;      // include/Frame.h:37-38

struct Grid {
    std::vector<int> cell[GRID_COLS][GRID_ROWS];
    float inv_w, inv_h;
};

void build_grid(const FrameView& F, Grid& G) {
    G.inv_w = (float)GRID_COLS / (F.max_x - F.min_x);
    G.inv_h = (float)GRID_ROWS / (F.max_y - F.min_y);
    for (int i = 0; i < F.n; ++i) {
        const int px = (int)std::round((F.keys_un[i].x - F.min_x) * G.inv_w);      // roundf: half away from zero
        const int py = (int)std::round((F.keys_un[i].y - F.min_y) * G.inv_h);
        if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
        G.cell[px][py].push_back(i);
    }
}

void features_in_area(const FrameView& F, const Grid& G, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) {
    out.clear();
    const int cx0 = std::max(0, (int)std::floor((x - F.min_x - r) * G.inv_w));
    if (cx0 >= GRID_COLS) return;
    const int cx1 = std::min(GRID_COLS - 1, (int)std::ceil((x - F.min_x + r) * G.inv_w));
    if (cx1 < 0) return;
    const int cy0 = std::max(0, (int)std::floor((y - F.min_y - r) * G.inv_h));
    if (cy0 >= GRID_ROWS) return;
    const int cy1 = std::min(GRID_ROWS - 1, (int)std::ceil((y - F.min_y + r) * G.inv_h));
    if (cy1 < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = cx0; ix <= cx1; ++ix)
        for (int iy = cy0; iy <= cy1; ++iy)
            for (int j : G.cell[ix][iy]) {
                const KeyPoint& kp = F.keys_un[j];
                if (check) {
                    if (kp.octave < minLevel) continue;
                    if (maxLevel >= 0 && kp.octave > maxLevel) continue;
                }
                const float dx = kp.x - x, dy = kp.y - y;
                if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
            }
}

// float 3x3 * 3-vector + 3-vector as cv::Mat CV_32F arithmetic does it (double accumulation, one rounding per op)
inline void rt(const float* T, const float* P, float out[3]) {
    for (int r = 0; r < 3; ++r) {
        const float m = (float)((double)T[4 * r] * P[0] + (double)T[4 * r + 1] * P[1] + (double)T[4 * r + 2] * P[2]);
        out[r] = m + T[4 * r + 3];
    }
}
inline void camera_center(const float* T, float Ow[3]) {       // mOw = -mRcw.t() * mtcw
    for (int r = 0; r < 3; ++r) {
        const float nr0 = -T[0 + r], nr1 = -T[4 + r], nr2 = -T[8 + r];      // (-R^T) row r
        Ow[r] = (float)((double)nr0 * T[3] + (double)nr1 * T[7] + (double)nr2 * T[11]);
    }
}

struct Proj { bool in_view; float u, v, ur, view_cos; int level; };

Proj in_frustum(const FrameView& F, const MapPointsView& M, int k, const float Ow[3], float cos_limit) {
    Proj p{false, 0, 0, 0, 0, 0};
    const float* P = M.pos + 3 * k;
    float Pc[3];
    rt(F.Tcw, P, Pc);
    if (Pc[2] < 0.0f) return p;
    const float invz = 1.0f / Pc[2];
    const float u = F.fx * Pc[0] * invz + F.cx, v = F.fy * Pc[1] * invz + F.cy;
    if (u < F.min_x || u > F.max_x) return p;
    if (v < F.min_y || v > F.max_y) return p;
    const float maxD = 1.2f * M.max_distance[k], minD = 0.8f * M.min_distance[k];
    const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
    const float dist = (float)std::sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
    if (dist < minD || dist > maxD) return p;
    const float* Pn = M.normal + 3 * k;
    const float view_cos = (float)(((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) / dist);
    if (view_cos < cos_limit) return p;
    const float ratio = M.max_distance[k] / dist;
    int lvl = (int)std::ceil((float)std::log((double)ratio) / F.log_scale_factor);
    if (lvl < 0) lvl = 0; else if (lvl >= F.n_levels) lvl = F.n_levels - 1;
    p.in_view = true; p.u = u; p.v = v; p.ur = u - F.bf * invz; p.view_cos = view_cos; p.level = lvl;
    return p;
}

}  // namespace

int search_by_projection_map(const FrameView& F, const MapPointsView& M, float th, float nnratio, int32_t* matches, uint8_t* in_view) {
    Grid G;
    build_grid(F, G);
    float Ow[3];
    camera_center(F.Tcw, Ow);
    int nmatches = 0;
    const bool bFactor = th != 1.0f;
    std::vector<int> cand;
    for (int k = 0; k < M.n; ++k) {
        in_view[k] = 0;
        if (M.skip[k]) continue;                                  // Tracking::SearchLocalPoints :2306-2309
        const Proj p = in_frustum(F, M, k, Ow, 0.5f);
        if (!p.in_view) continue;
        in_view[k] = 1;
        float r = p.view_cos > 0.998 ? 2.5f : 4.0f;               // RadiusByViewingCos (double compare of a float)
        if (bFactor) r *= th;
        const float rs = r * F.scale_factors[p.level];
        features_in_area(F, G, p.u, p.v, rs, p.level - 1, p.level, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
        for (int idx : cand) {
            if (matches[idx] >= 0 && M.has_obs[matches[idx]]) continue;
            if (F.u_right[idx] > 0) {
                const float er = std::fabs(p.ur - F.u_right[idx]);
                if (er > rs) continue;
            }
            const int dist = descriptor_distance(M.desc + 32 * k, F.desc + 32 * idx);
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = F.keys_un[idx].octave; bestIdx = idx; }
            else if (dist < bestDist2) { bestLevel2 = F.keys_un[idx].octave; bestDist2 = dist; }
        }
        if (bestDist <= 100) {                                    // TH_HIGH
            if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
            matches[bestIdx] = k;
            ++nmatches;
        }
    }
    return nmatches;
}

int search_by_projection_last(const FrameView& C, const LastFrameView& L, const MapPointsView& M, float th, bool mono, bool check_ori,
                              int32_t* matches) {
    Grid G;
    build_grid(C, G);
    const int HISTO = 30;
    std::vector<int> rot[HISTO];
    const float factor = 1.0f / HISTO;
    // tlc = Rlw * twc + tlw, twc = -Rcw^T tcw
    float twc[3];
    camera_center(C.Tcw, twc);
    float tlc[3];
    rt(L.Tcw, twc, tlc);
    const float mb = C.bf / C.fx;
    const bool fwd = tlc[2] > mb && !mono, bwd = -tlc[2] > mb && !mono;
    int nmatches = 0;
    std::vector<int> cand;
    for (int i = 0; i < L.n; ++i) {
        const int mp = L.map_point[i];
        if (mp < 0 || L.outlier[i]) continue;
        float xc[3];
        rt(C.Tcw, M.pos + 3 * mp, xc);
        const float invzc = (float)(1.0 / xc[2]);
        if (invzc < 0) continue;
        const float u = C.fx * xc[0] * invzc + C.cx, v = C.fy * xc[1] * invzc + C.cy;
        if (u < C.min_x || u > C.max_x) continue;
        if (v < C.min_y || v > C.max_y) continue;
        const int oct = L.keys[i].octave;
        const float radius = th * C.scale_factors[oct];
        if (fwd) features_in_area(C, G, u, v, radius, oct, -1, cand);
        else if (bwd) features_in_area(C, G, u, v, radius, 0, oct, cand);
        else features_in_area(C, G, u, v, radius, oct - 1, oct + 1, cand);
        if (cand.empty()) continue;
        int bestDist = 256, bestIdx2 = -1;
        for (int i2 : cand) {
            if (matches[i2] >= 0 && M.has_obs[matches[i2]]) continue;
            if (C.u_right[i2] > 0) {
                const float ur = u - C.bf * invzc;
                const float er = std::fabs(ur - C.u_right[i2]);
                if (er > radius) continue;
            }
            const int d = descriptor_distance(M.desc + 32 * mp, C.desc + 32 * i2);
            if (d < bestDist) { bestDist = d; bestIdx2 = i2; }
        }
        if (bestDist <= 100) {
            matches[bestIdx2] = mp;
            ++nmatches;
            if (check_ori) {
                float rotd = L.keys[i].angle - C.keys_un[bestIdx2].angle;
                if (rotd < 0.0) rotd += 360.0f;
                int bin = (int)std::round(rotd * factor);
                if (bin == HISTO) bin = 0;
                rot[bin].push_back(bestIdx2);
            }
        }
    }
    if (check_ori) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO; ++i) {
            const int s = (int)rot[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
        else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO; ++i)
            if (i != ind1 && i != ind2 && i != ind3)
                for (int j : rot[i]) { matches[j] = -1; --nmatches; }
    }
    return nmatches;
}

}  // namespace oracle

extern "C" {
struct orc_frame_view {
    int32_t n; const oracle::KeyPoint* keys_un; const float* u_right; const uint8_t* desc; float Tcw[16];
    float fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y; int32_t n_levels; const float* scale_factors; float log_scale_factor;
};
struct orc_map_points { int32_t n; const float *pos, *normal, *max_distance, *min_distance; const uint8_t *desc, *skip, *has_obs; };
struct orc_last_frame { int32_t n; const oracle::KeyPoint* keys; const int32_t* map_point; const uint8_t* outlier; float Tcw[16]; };
static oracle::FrameView fv(const orc_frame_view* f) {
    oracle::FrameView F;
    F.n = f->n; F.keys_un = f->keys_un; F.u_right = f->u_right; F.desc = f->desc;
    for (int i = 0; i < 16; ++i) F.Tcw[i] = f->Tcw[i];
    F.fx = f->fx; F.fy = f->fy; F.cx = f->cx; F.cy = f->cy; F.bf = f->bf; F.min_x = f->min_x; F.max_x = f->max_x; F.min_y = f->min_y; F.max_y = f->max_y;
    F.n_levels = f->n_levels; F.scale_factors = f->scale_factors; F.log_scale_factor = f->log_scale_factor;
    return F;
}
static oracle::MapPointsView mv(const orc_map_points* m) {
    oracle::MapPointsView M;
    M.n = m->n; M.pos = m->pos; M.normal = m->normal; M.max_distance = m->max_distance; M.min_distance = m->min_distance;
    M.desc = m->desc; M.skip = m->skip; M.has_obs = m->has_obs;
    return M;
}
int orc_search_by_projection_map(const orc_frame_view* f, const orc_map_points* m, float th, float nnratio, int32_t* matches, uint8_t* in_view) {
    return oracle::search_by_projection_map(fv(f), mv(m), th, nnratio, matches, in_view);
}
int orc_search_by_projection_last(const orc_frame_view* f, const orc_last_frame* l, const orc_map_points* m, float th, int mono, int check_ori,
                                  int32_t* matches) {
    oracle::LastFrameView L;
    L.n = l->n; L.keys = l->keys; L.map_point = l->map_point; L.outlier = l->outlier;
    for (int i = 0; i < 16; ++i) L.Tcw[i] = l->Tcw[i];
    return oracle::search_by_projection_last(fv(f), L, mv(m), th, mono != 0, check_ori != 0, matches);
}
}
