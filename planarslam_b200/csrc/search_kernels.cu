// Projection-guided descriptor search for sm_100a (one frame + one map snapshot per call; frames of a replayed
// sequence are issued back to back on the context's stream).
//
// Reference semantics: Frame::AssignFeaturesToGrid / PosInGrid src/Frame.cc:155-168,526-535; Frame::GetFeaturesInArea :440-489;
// Frame::isInFrustum :312-367; MapPoint::PredictScale src/MapPoint.cc:419-434; ORBmatcher::SearchByProjection(Frame&,
// vector<MapPoint*>&, th) src/ORBmatcher.cc:46-130; ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono)
// :1396-1535; ComputeThreeMaxima :1666-1707.  Arithmetic conventions as in oracle/search.h (float cv::Mat products and
// cv::norm accumulate in double; log() of PredictScale in double; no FMA: the file is built with --fmad=false).
//
// Work decomposition:
//   k_search_grid        one CTA: sort keypoints by (cell, index) -> the reference's per-cell lists in push order
//   k_candidates_*       one warp per map point / last-frame keypoint: projection, window cells in (ix, iy) order, level and
//                        stereo gates, 256-bit Hamming distance by popcount; ordered candidate list per point (ballot compaction)
//   k_resolve_*          one warp: the reference's sequential greedy assignment (a keypoint that already holds a map point
//                        with observations is skipped by later points), top-2 / top-1 by warp-shuffle merge per point
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <vector>

#include "pslam_internal.h"

namespace pslam {

#define SG_COLS 64
#define SG_ROWS 48
#define SG_CELLS (SG_COLS * SG_ROWS)
#define SEARCH_CAND_CAP 128
#define SEARCH_MAX_KP 4096

struct SearchFrameDev {
    int n;
    const pslam_keypoint* keys_un; const float* u_right; const uint8_t* desc;
    float Tcw[16];
    float fx, fy, cx, cy, bf, min_x, max_x, min_y, max_y;
    int n_levels; float scale[PSLAM_MAX_LEVELS]; float log_scale_factor;
    float inv_w, inv_h;
};
struct SearchMapDev {
    int n;
    const float *pos, *normal, *max_distance, *min_distance; const uint8_t *desc, *skip, *has_obs;
};

// ---- grid: sorted (cell << 16 | index) keys; keypoints outside the grid get cell = SG_CELLS (sorted to the end) ----
__global__ void __launch_bounds__(1024) k_search_grid(SearchFrameDev F, int32_t* __restrict__ cell_start /*[SG_CELLS+1]*/, int32_t* __restrict__ items /*[n]*/) {
    __shared__ uint32_t keys[SEARCH_MAX_KP];
    const int tid = threadIdx.x;
    for (int i = tid; i < SEARCH_MAX_KP; i += 1024) {
        uint32_t k = 0xffffffffu;
        if (i < F.n) {
            const int px = (int)roundf((F.keys_un[i].x - F.min_x) * F.inv_w), py = (int)roundf((F.keys_un[i].y - F.min_y) * F.inv_h);
            const int cell = (px < 0 || px >= SG_COLS || py < 0 || py >= SG_ROWS) ? SG_CELLS : px * SG_ROWS + py;     // mGrid[x][y]
            k = ((uint32_t)cell << 16) | (uint32_t)i;
        }
        keys[i] = k;
    }
    __syncthreads();
    for (int size = 2; size <= SEARCH_MAX_KP; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < SEARCH_MAX_KP / 2; i += 1024) {
                const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
                const bool up = ((lo & size) == 0);
                const uint32_t a = keys[lo], b = keys[hi];
                if ((a > b) == up) { keys[lo] = b; keys[hi] = a; }
            }
            __syncthreads();
        }
    for (int i = tid; i < F.n; i += 1024) items[i] = (int)(keys[i] & 0xffff);
    for (int c = tid; c <= SG_CELLS; c += 1024) {      // cell_start[c] = first position whose cell >= c (binary search)
        int lo = 0, hi = F.n;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if ((int)(keys[mid] >> 16) < c) lo = mid + 1; else hi = mid; }
        cell_start[c] = lo;
    }
}

__device__ __forceinline__ void mat_rt(const float* T, const float* P, float out[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float m = (float)((double)T[4 * r] * P[0] + (double)T[4 * r + 1] * P[1] + (double)T[4 * r + 2] * P[2]);
        out[r] = m + T[4 * r + 3];
    }
}
__device__ __forceinline__ void camera_center(const float* T, float Ow[3]) {
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const float a = -T[0 + r], b = -T[4 + r], c = -T[8 + r];
        Ow[r] = (float)((double)a * T[3] + (double)b * T[7] + (double)c * T[11]);
    }
}
__device__ __forceinline__ int hamming32(const uint8_t* a, const uint8_t* b) {
    const uint4* pa = reinterpret_cast<const uint4*>(a);
    const uint4* pb = reinterpret_cast<const uint4*>(b);
    const uint4 a0 = pa[0], a1 = pa[1], b0 = pb[0], b1 = pb[1];
    return __popc(a0.x ^ b0.x) + __popc(a0.y ^ b0.y) + __popc(a0.z ^ b0.z) + __popc(a0.w ^ b0.w) + __popc(a1.x ^ b1.x) + __popc(a1.y ^ b1.y) +
           __popc(a1.z ^ b1.z) + __popc(a1.w ^ b1.w);
}

// Ordered candidate gathering shared by both searches: all keypoints in the window cells (ix outer, iy inner, push order inside
// a cell) that pass the level gate, the |dx|,|dy| < r gate and the stereo gate; entry = idx | dist << 16 | octave << 26.
__device__ __forceinline__ int gather_candidates(const SearchFrameDev& F, const int32_t* cell_start, const int32_t* items, float x, float y, float r,
                                                 int minLevel, int maxLevel, bool stereo_gate, float ur_proj, float er_max, const uint8_t* qdesc,
                                                 uint32_t* out, bool& overflow) {
    const int lane = threadIdx.x & 31;
    const int cx0 = max(0, (int)floorf((x - F.min_x - r) * F.inv_w));
    if (cx0 >= SG_COLS) return 0;
    const int cx1 = min(SG_COLS - 1, (int)ceilf((x - F.min_x + r) * F.inv_w));
    if (cx1 < 0) return 0;
    const int cy0 = max(0, (int)floorf((y - F.min_y - r) * F.inv_h));
    if (cy0 >= SG_ROWS) return 0;
    const int cy1 = min(SG_ROWS - 1, (int)ceilf((y - F.min_y + r) * F.inv_h));
    if (cy1 < 0) return 0;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    int cnt = 0;
    for (int ix = cx0; ix <= cx1; ++ix) {
        // cells (ix, cy0..cy1) are contiguous in the sorted order: one run per column
        const int beg = cell_start[ix * SG_ROWS + cy0], end = cell_start[ix * SG_ROWS + cy1 + 1];
        for (int p0 = beg; p0 < end; p0 += 32) {
            const int p = p0 + lane;
            bool ok = p < end;
            int j = 0, d = 0, oct = 0;
            if (ok) {
                j = items[p];
                const pslam_keypoint kp = F.keys_un[j];
                oct = kp.octave;
                if (check) { if (oct < minLevel) ok = false; if (maxLevel >= 0 && oct > maxLevel) ok = false; }
                const float dx = kp.x - x, dy = kp.y - y;
                if (!(fabsf(dx) < r && fabsf(dy) < r)) ok = false;
                if (ok && stereo_gate) { const float urj = F.u_right[j]; if (urj > 0 && fabsf(ur_proj - urj) > er_max) ok = false; }
                if (ok) d = hamming32(qdesc, F.desc + (size_t)j * 32);
            }
            const uint32_t m = __ballot_sync(0xffffffffu, ok);
            if (ok) {
                const int pos = cnt + __popc(m & ((1u << lane) - 1));
                if (pos < SEARCH_CAND_CAP) out[pos] = (uint32_t)j | ((uint32_t)d << 16) | ((uint32_t)oct << 26); else overflow = true;
            }
            cnt += __popc(m);
        }
    }
    return min(cnt, SEARCH_CAND_CAP);
}

// ---- SearchByProjection(Frame, local map points): candidates ----
__global__ void __launch_bounds__(256) k_candidates_map(SearchFrameDev F, SearchMapDev M, float th, const int32_t* __restrict__ cell_start,
                                                        const int32_t* __restrict__ items, uint32_t* __restrict__ cand, int32_t* __restrict__ cand_n,
                                                        uint8_t* __restrict__ in_view, int32_t* __restrict__ status) {
    const int k = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (k >= M.n) return;
    int n_c = 0;
    bool view = false;
    if (!M.skip[k]) {
        float Ow[3], Pc[3];
        camera_center(F.Tcw, Ow);
        const float* P = M.pos + 3 * k;
        mat_rt(F.Tcw, P, Pc);
        if (!(Pc[2] < 0.0f)) {
            const float invz = 1.0f / Pc[2];
            const float u = F.fx * Pc[0] * invz + F.cx, v = F.fy * Pc[1] * invz + F.cy;
            if (!(u < F.min_x || u > F.max_x) && !(v < F.min_y || v > F.max_y)) {
                const float maxD = 1.2f * M.max_distance[k], minD = 0.8f * M.min_distance[k];
                const float PO[3] = {P[0] - Ow[0], P[1] - Ow[1], P[2] - Ow[2]};
                const float dist = (float)sqrt((double)PO[0] * PO[0] + (double)PO[1] * PO[1] + (double)PO[2] * PO[2]);
                if (!(dist < minD || dist > maxD)) {
                    const float* Pn = M.normal + 3 * k;
                    const float view_cos = (float)(((double)PO[0] * Pn[0] + (double)PO[1] * Pn[1] + (double)PO[2] * Pn[2]) / dist);
                    if (!(view_cos < 0.5f)) {
                        const float ratio = M.max_distance[k] / dist;
                        int lvl = (int)ceilf((float)log((double)ratio) / F.log_scale_factor);
                        if (lvl < 0) lvl = 0; else if (lvl >= F.n_levels) lvl = F.n_levels - 1;
                        view = true;
                        float r = ((double)view_cos > 0.998) ? 2.5f : 4.0f;
                        if (th != 1.0f) r *= th;
                        const float rs = r * F.scale[lvl];
                        bool overflow = false;
                        n_c = gather_candidates(F, cell_start, items, u, v, rs, lvl - 1, lvl, true, u - F.bf * invz, rs, M.desc + (size_t)k * 32,
                                                cand + (size_t)k * SEARCH_CAND_CAP, overflow);
                        if (__any_sync(0xffffffffu, overflow) && lane == 0) atomicOr(status, 64);
                    }
                }
            }
        }
    }
    if (lane == 0) { cand_n[k] = n_c; in_view[k] = view ? 1 : 0; }
}

// warp top-2 over keys (dist << 8 | position): returns the two smallest
__device__ __forceinline__ void warp_top2(uint32_t& k0, uint32_t& k1) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const uint32_t lo = min(k0, o0), hi = max(k0, o0);
        k1 = min(hi, min(k1, o1));
        k0 = lo;
    }
}

__global__ void __launch_bounds__(32) k_resolve_map(SearchMapDev M, float nnratio, const uint32_t* __restrict__ cand, const int32_t* __restrict__ cand_n,
                                                    int32_t* __restrict__ matches, int32_t* __restrict__ n_matches) {
    const int lane = threadIdx.x;
    const uint32_t NONE = 0xffffffffu;
    int nm = 0;
    for (int k = 0; k < M.n; ++k) {
        const int nc = cand_n[k];
        if (nc == 0) continue;
        uint32_t k0 = NONE, k1 = NONE;
        for (int p = lane; p < nc; p += 32) {
            const uint32_t e = cand[(size_t)k * SEARCH_CAND_CAP + p];
            const int idx = e & 0xffff;
            const int cur = matches[idx];
            if (cur >= 0 && M.has_obs[cur]) continue;                 // already holds a map point with observations (:83-85)
            const uint32_t key = (((e >> 16) & 0x3ff) << 8) | (uint32_t)p;
            if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
        }
        warp_top2(k0, k1);
        if (k0 == NONE) continue;
        const int bestDist = k0 >> 8;
        if (bestDist <= 100) {
            const uint32_t e0 = cand[(size_t)k * SEARCH_CAND_CAP + (k0 & 0xff)];
            const int bestLevel = e0 >> 26;
            int bestDist2 = 256, bestLevel2 = -1;
            if (k1 != NONE) { const uint32_t e1 = cand[(size_t)k * SEARCH_CAND_CAP + (k1 & 0xff)]; bestDist2 = k1 >> 8; bestLevel2 = e1 >> 26; }
            if (bestLevel == bestLevel2 && (float)bestDist > nnratio * (float)bestDist2) continue;
            if (lane == 0) matches[e0 & 0xffff] = k;
            ++nm;
            __syncwarp();
        }
    }
    if (lane == 0) *n_matches = nm;
}

// ---- SearchByProjection(current, last): candidates per last-frame keypoint ----
struct SearchLastDev { int n; const pslam_keypoint* keys; const int32_t* map_point; const uint8_t* outlier; float Tcw[16]; };

__global__ void __launch_bounds__(256) k_candidates_last(SearchFrameDev C, SearchLastDev L, SearchMapDev M, float th, int mono,
                                                         const int32_t* __restrict__ cell_start, const int32_t* __restrict__ items,
                                                         uint32_t* __restrict__ cand, int32_t* __restrict__ cand_n, int32_t* __restrict__ status) {
    const int i = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (i >= L.n) return;
    int n_c = 0;
    const int mp = L.map_point[i];
    if (mp >= 0 && !L.outlier[i]) {
        float twc[3], tlc[3], xc[3];
        camera_center(C.Tcw, twc);
        mat_rt(L.Tcw, twc, tlc);
        const float mb = C.bf / C.fx;
        const bool fwd = tlc[2] > mb && !mono, bwd = -tlc[2] > mb && !mono;
        mat_rt(C.Tcw, M.pos + 3 * mp, xc);
        const float invzc = (float)(1.0 / (double)xc[2]);
        if (!(invzc < 0)) {
            const float u = C.fx * xc[0] * invzc + C.cx, v = C.fy * xc[1] * invzc + C.cy;
            if (!(u < C.min_x || u > C.max_x) && !(v < C.min_y || v > C.max_y)) {
                const int oct = L.keys[i].octave;
                const float radius = th * C.scale[oct];
                int lo, hi;
                if (fwd) { lo = oct; hi = -1; } else if (bwd) { lo = 0; hi = oct; } else { lo = oct - 1; hi = oct + 1; }
                bool overflow = false;
                n_c = gather_candidates(C, cell_start, items, u, v, radius, lo, hi, true, u - C.bf * invzc, radius, M.desc + (size_t)mp * 32,
                                        cand + (size_t)i * SEARCH_CAND_CAP, overflow);
                if (__any_sync(0xffffffffu, overflow) && lane == 0) atomicOr(status, 64);
            }
        }
    }
    if (lane == 0) cand_n[i] = n_c;
}

__global__ void __launch_bounds__(32) k_resolve_last(SearchFrameDev C, SearchLastDev L, SearchMapDev M, int check_ori, const uint32_t* __restrict__ cand,
                                                     const int32_t* __restrict__ cand_n, int32_t* __restrict__ matches, int32_t* __restrict__ n_matches,
                                                     int32_t* __restrict__ hist_idx /*[L.n]*/, int8_t* __restrict__ hist_bin /*[L.n]*/) {
    const int lane = threadIdx.x;
    const uint32_t NONE = 0xffffffffu;
    __shared__ int s_cnt[32];
    if (lane < 30) s_cnt[lane] = 0;
    __syncwarp();
    int nm = 0, n_push = 0;
    for (int i = 0; i < L.n; ++i) {
        const int nc = cand_n[i];
        if (nc == 0) continue;
        const int mp = L.map_point[i];
        uint32_t k0 = NONE, k1 = NONE;
        for (int p = lane; p < nc; p += 32) {
            const uint32_t e = cand[(size_t)i * SEARCH_CAND_CAP + p];
            const int cur = matches[e & 0xffff];
            if (cur >= 0 && M.has_obs[cur]) continue;
            const uint32_t key = (((e >> 16) & 0x3ff) << 8) | (uint32_t)p;
            if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
        }
        warp_top2(k0, k1);
        if (k0 == NONE || (int)(k0 >> 8) > 100) continue;
        const int idx2 = cand[(size_t)i * SEARCH_CAND_CAP + (k0 & 0xff)] & 0xffff;
        if (lane == 0) matches[idx2] = mp;
        ++nm;
        if (check_ori) {
            float rot = L.keys[i].angle - C.keys_un[idx2].angle;
            if (rot < 0.0f) rot += 360.0f;
            int bin = (int)roundf(rot * (1.0f / 30));
            if (bin == 30) bin = 0;
            if (lane == 0) { hist_idx[n_push] = idx2; hist_bin[n_push] = (int8_t)bin; ++s_cnt[bin]; }
            ++n_push;
        }
        __syncwarp();
    }
    if (check_ori) {
        __syncwarp();
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int b = 0; b < 30; ++b) {
            const int s = s_cnt[b];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = b; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = b; }
            else if (s > max3) { max3 = s; ind3 = b; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        int removed = 0;
        for (int j = lane; j < n_push; j += 32) {
            const int b = hist_bin[j];
            if (b != ind1 && b != ind2 && b != ind3) { matches[hist_idx[j]] = -1; ++removed; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) removed += __shfl_xor_sync(0xffffffffu, removed, o);
        nm -= removed;
    }
    if (lane == 0) *n_matches = nm;
}

// ---------------------------------------------------------------------------------------------------------
struct SearchBuffers {
    void* blob = nullptr; size_t cap = 0;
    uint32_t* d_cand = nullptr; int32_t* d_cand_n = nullptr; size_t cap_pts = 0;
    int32_t *d_cell_start = nullptr, *d_items = nullptr, *d_matches = nullptr, *d_scalar = nullptr, *d_hist_idx = nullptr;
    int8_t* d_hist_bin = nullptr; uint8_t* d_in_view = nullptr;
};

template <typename T>
static T* carve(uint8_t*& p, size_t n) { T* r = reinterpret_cast<T*>(p); p += (n * sizeof(T) + 255) / 256 * 256; return r; }

static int ensure(pslam_ctx* c, SearchBuffers& B, size_t npts) {
    if (!B.d_cell_start) {
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cell_start, (SG_CELLS + 1) * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_items, SEARCH_MAX_KP * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_matches, SEARCH_MAX_KP * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_scalar, 4 * sizeof(int32_t)));
    }
    if (npts > B.cap_pts) {
        cudaFree(B.d_cand); cudaFree(B.d_cand_n); cudaFree(B.d_hist_idx); cudaFree(B.d_hist_bin); cudaFree(B.d_in_view);
        const size_t n = npts * 3 / 2 + 64;
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cand, n * SEARCH_CAND_CAP * sizeof(uint32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_cand_n, n * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_hist_idx, n * sizeof(int32_t)));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_hist_bin, n));
        PSLAM_CUDA(c, cudaMalloc((void**)&B.d_in_view, n));
        B.cap_pts = n;
    }
    return PSLAM_OK;
}

// uploads the frame / map arrays into one staging blob; fills the device views
static int upload(pslam_ctx* c, SearchBuffers& B, const pslam_frame_view* f, const pslam_map_points* m, const pslam_last_frame* l, SearchFrameDev& F,
                  SearchMapDev& M, SearchLastDev& Ld) {
    const size_t nf = f->n, nm = m->n, nl = l ? l->n : 0;
    const size_t need = 4096 + nf * (28 + 4 + 32) + nm * (12 + 12 + 4 + 4 + 32 + 1 + 1) + nl * (28 + 4 + 1) + 256 * 16;
    if (need > B.cap) { cudaFree(B.blob); B.blob = nullptr; B.cap = 0; PSLAM_CUDA(c, cudaMalloc(&B.blob, need * 3 / 2)); B.cap = need * 3 / 2; }
    uint8_t* p = (uint8_t*)B.blob;
    cudaStream_t st = c->stream;
#define UP(dst, src, type, count) { type* d_ = carve<type>(p, (count) ? (count) : 1); if (count) PSLAM_CUDA(c, cudaMemcpyAsync(d_, src, (size_t)(count) * sizeof(type), cudaMemcpyHostToDevice, st)); dst = d_; }
    F.n = (int)nf;
    UP(F.keys_un, f->keys_un, pslam_keypoint, nf); UP(F.u_right, f->u_right, float, nf); UP(F.desc, f->desc, uint8_t, nf * 32);
    std::memcpy(F.Tcw, f->Tcw, sizeof F.Tcw);
    F.fx = f->fx; F.fy = f->fy; F.cx = f->cx; F.cy = f->cy; F.bf = f->bf; F.min_x = f->min_x; F.max_x = f->max_x; F.min_y = f->min_y; F.max_y = f->max_y;
    F.n_levels = f->n_levels; F.log_scale_factor = f->log_scale_factor;
    for (int i = 0; i < f->n_levels && i < PSLAM_MAX_LEVELS; ++i) F.scale[i] = f->scale_factors[i];
    F.inv_w = (float)SG_COLS / (f->max_x - f->min_x); F.inv_h = (float)SG_ROWS / (f->max_y - f->min_y);
    M.n = (int)nm;
    UP(M.pos, m->pos, float, nm * 3); UP(M.normal, m->normal, float, nm * 3); UP(M.max_distance, m->max_distance, float, nm);
    UP(M.min_distance, m->min_distance, float, nm); UP(M.desc, m->desc, uint8_t, nm * 32); UP(M.skip, m->skip, uint8_t, nm); UP(M.has_obs, m->has_obs, uint8_t, nm);
    if (l) {
        Ld.n = (int)nl;
        UP(Ld.keys, l->keys, pslam_keypoint, nl); UP(Ld.map_point, l->map_point, int32_t, nl); UP(Ld.outlier, l->outlier, uint8_t, nl);
        std::memcpy(Ld.Tcw, l->Tcw, sizeof Ld.Tcw);
    }
#undef UP
    return PSLAM_OK;
}

static int validate(pslam_ctx* c, const pslam_frame_view* f, const pslam_map_points* m) {
    if (!f || !m || f->n < 0 || m->n < 0) return set_error(c, PSLAM_E_INVALID, "null view");
    if (f->n > SEARCH_MAX_KP || f->n > 65535) return set_error(c, PSLAM_E_INVALID, "more than 4096 keypoints in the frame view");
    if (f->n_levels < 1 || f->n_levels > PSLAM_MAX_LEVELS || !(f->max_x > f->min_x) || !(f->max_y > f->min_y)) return set_error(c, PSLAM_E_INVALID, "bad frame view");
    return PSLAM_OK;
}

}  // namespace pslam

using namespace pslam;

extern "C" {

int pslam_search_by_projection_map(pslam_ctx* c, const pslam_frame_view* f, const pslam_map_points* m, float th, float nnratio, int32_t* matches_io,
                                   uint8_t* in_view) {
    if (!c) return PSLAM_E_INVALID;
    int rc = validate(c, f, m);
    if (rc != PSLAM_OK) return rc;
    if (!matches_io) return set_error(c, PSLAM_E_INVALID, "null matches");
    if (f->n == 0 || m->n == 0) { if (in_view && m->n) std::memset(in_view, 0, m->n); return 0; }
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    if (!c->search) c->search = new SearchBuffers();
    SearchBuffers& B = *c->search;
    if ((rc = ensure(c, B, (size_t)m->n)) != PSLAM_OK) return rc;
    SearchFrameDev F; SearchMapDev M; SearchLastDev Ld;
    if ((rc = upload(c, B, f, m, nullptr, F, M, Ld)) != PSLAM_OK) return rc;
    cudaStream_t st = c->stream;
    PSLAM_CUDA(c, cudaMemcpyAsync(B.d_matches, matches_io, f->n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PSLAM_CUDA(c, cudaMemsetAsync(B.d_scalar, 0, 4 * sizeof(int32_t), st));
    PSLAM_LAUNCH(c, "search_grid", k_search_grid<<<1, 1024, 0, st>>>(F, B.d_cell_start, B.d_items));
    PSLAM_LAUNCH(c, "search_candidates_map", k_candidates_map<<<(M.n + 7) / 8, 256, 0, st>>>(F, M, th, B.d_cell_start, B.d_items, B.d_cand, B.d_cand_n,
                 B.d_in_view, B.d_scalar + 1));
    PSLAM_LAUNCH(c, "search_resolve_map", k_resolve_map<<<1, 32, 0, st>>>(M, nnratio, B.d_cand, B.d_cand_n, B.d_matches, B.d_scalar));
    int32_t sc[2] = {0, 0};
    PSLAM_CUDA(c, cudaMemcpyAsync(matches_io, B.d_matches, f->n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    if (in_view) PSLAM_CUDA(c, cudaMemcpyAsync(in_view, B.d_in_view, m->n, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaMemcpyAsync(sc, B.d_scalar, sizeof sc, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaStreamSynchronize(st));
    if (sc[1]) return set_error(c, PSLAM_E_CAPACITY, "more than 128 candidates in one search window");
    return sc[0];
}

int pslam_search_by_projection_last(pslam_ctx* c, const pslam_frame_view* cur, const pslam_last_frame* last, const pslam_map_points* m, float th,
                                    int mono, int check_orientation, int32_t* matches_io) {
    if (!c) return PSLAM_E_INVALID;
    int rc = validate(c, cur, m);
    if (rc != PSLAM_OK) return rc;
    if (!last || last->n < 0 || !matches_io) return set_error(c, PSLAM_E_INVALID, "null last-frame view or matches");
    if (cur->n == 0 || last->n == 0) return 0;
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    if (!c->search) c->search = new SearchBuffers();
    SearchBuffers& B = *c->search;
    if ((rc = ensure(c, B, (size_t)last->n)) != PSLAM_OK) return rc;
    SearchFrameDev F; SearchMapDev M; SearchLastDev Ld;
    if ((rc = upload(c, B, cur, m, last, F, M, Ld)) != PSLAM_OK) return rc;
    cudaStream_t st = c->stream;
    PSLAM_CUDA(c, cudaMemcpyAsync(B.d_matches, matches_io, cur->n * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    PSLAM_CUDA(c, cudaMemsetAsync(B.d_scalar, 0, 4 * sizeof(int32_t), st));
    PSLAM_LAUNCH(c, "search_grid", k_search_grid<<<1, 1024, 0, st>>>(F, B.d_cell_start, B.d_items));
    PSLAM_LAUNCH(c, "search_candidates_last", k_candidates_last<<<(Ld.n + 7) / 8, 256, 0, st>>>(F, Ld, M, th, mono, B.d_cell_start, B.d_items, B.d_cand,
                 B.d_cand_n, B.d_scalar + 1));
    PSLAM_LAUNCH(c, "search_resolve_last", k_resolve_last<<<1, 32, 0, st>>>(F, Ld, M, check_orientation, B.d_cand, B.d_cand_n, B.d_matches, B.d_scalar,
                 B.d_hist_idx, B.d_hist_bin));
    int32_t sc[2] = {0, 0};
    PSLAM_CUDA(c, cudaMemcpyAsync(matches_io, B.d_matches, cur->n * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaMemcpyAsync(sc, B.d_scalar, sizeof sc, cudaMemcpyDeviceToHost, st));
    PSLAM_CUDA(c, cudaStreamSynchronize(st));
    if (sc[1]) return set_error(c, PSLAM_E_CAPACITY, "more than 128 candidates in one search window");
    return sc[0];
}

}  // extern "C"

namespace pslam {
void search_free(pslam_ctx* c) {
    if (!c->search) return;
    SearchBuffers& B = *c->search;
    cudaFree(B.blob); cudaFree(B.d_cand); cudaFree(B.d_cand_n); cudaFree(B.d_cell_start); cudaFree(B.d_items); cudaFree(B.d_matches);
    cudaFree(B.d_scalar); cudaFree(B.d_hist_idx); cudaFree(B.d_hist_bin); cudaFree(B.d_in_view);
    delete c->search;
    c->search = nullptr;
}
}  // namespace pslam
