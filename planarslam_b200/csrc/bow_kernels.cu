// ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&) on sm_100a (src/ORBmatcher.cc:160-292).  A frame feature belongs
// to exactly one vocabulary node, so the common nodes of the two feature vectors are independent sub-problems: one warp per
// common node walks the key-frame features of the node in order (an earlier match takes its frame feature away from later
// ones), the lanes evaluate the 256-bit Hamming distances to the node's frame features, a warp top-2 gives the reference's
// best / second-best, and a second kernel applies the rotation-histogram filter (ComputeThreeMaxima, :1666-1707).
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdint>
#include <vector>

#include "pslam_internal.h"

namespace pslam {

__device__ __forceinline__ void bow_top2(uint32_t& k0, uint32_t& k1) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        const uint32_t o0 = __shfl_xor_sync(0xffffffffu, k0, o), o1 = __shfl_xor_sync(0xffffffffu, k1, o);
        const uint32_t lo = min(k0, o0), hi = max(k0, o0);
        k1 = min(hi, min(k1, o1));
        k0 = lo;
    }
}

__global__ void __launch_bounds__(128) k_bow_nodes(int n_pairs, const int2* __restrict__ pairs, const uint8_t* __restrict__ kf_desc, const float* __restrict__ kf_angle,
                                                   const uint8_t* __restrict__ kf_has_mp, const int32_t* __restrict__ kf_off, const int32_t* __restrict__ kf_feat,
                                                   const uint8_t* __restrict__ f_desc, const float* __restrict__ f_angle, const int32_t* __restrict__ f_off,
                                                   const int32_t* __restrict__ f_feat, float nnratio, int check_ori, int32_t* __restrict__ match,
                                                   int8_t* __restrict__ bin_of, int32_t* __restrict__ hist, int32_t* __restrict__ nmatches) {
    const int lane = threadIdx.x & 31, w = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (w >= n_pairs) return;
    const int a = pairs[w].x, b = pairs[w].y;
    const int f0 = f_off[b], nfn = f_off[b + 1] - f0;
    const uint32_t NONE = 0xffffffffu;
    int nm = 0;
    for (int q = kf_off[a]; q < kf_off[a + 1]; ++q) {
        const int ik = kf_feat[q];
        if (!kf_has_mp[ik]) continue;
        uint32_t dk[8];
#pragma unroll
        for (int t = 0; t < 8; ++t) dk[t] = reinterpret_cast<const uint32_t*>(kf_desc)[8 * ik + t];
        uint32_t k0 = NONE, k1 = NONE;
        for (int p = lane; p < nfn; p += 32) {
            const int jf = f_feat[f0 + p];
            if (match[jf] >= 0) continue;
            int d = 0;
#pragma unroll
            for (int t = 0; t < 8; ++t) d += __popc(dk[t] ^ reinterpret_cast<const uint32_t*>(f_desc)[8 * jf + t]);
            const uint32_t key = ((uint32_t)d << 16) | (uint32_t)p;      // ties: the first frame feature of the node wins (strict <)
            if (key < k0) { k1 = k0; k0 = key; } else if (key < k1) k1 = key;
        }
        bow_top2(k0, k1);
        if (k0 == NONE) continue;
        const int bestDist1 = (int)(k0 >> 16), bestDist2 = k1 == NONE ? 256 : (int)(k1 >> 16);
        if (bestDist1 <= 50 && (float)bestDist1 < __fmul_rn(nnratio, (float)bestDist2)) {
            const int jf = f_feat[f0 + (k0 & 0xffff)];
            if (lane == 0) {
                match[jf] = ik;
                if (check_ori) {
                    float rot = __fsub_rn(kf_angle[ik], f_angle[jf]);
                    if (rot < 0.0f) rot = __fadd_rn(rot, 360.0f);
                    int bin = (int)roundf(__fmul_rn(rot, 1.0f / 30));
                    if (bin == 30) bin = 0;
                    bin_of[jf] = (int8_t)bin;
                    atomicAdd(&hist[bin], 1);
                }
            }
            ++nm;
            __syncwarp();
        }
    }
    if (lane == 0 && nm) atomicAdd(nmatches, nm);
}

__global__ void __launch_bounds__(256) k_bow_orientation(int nf, int32_t* __restrict__ match, const int8_t* __restrict__ bin_of, const int32_t* __restrict__ hist,
                                                         int32_t* __restrict__ nmatches) {
    __shared__ int s_ind[3];
    if (threadIdx.x == 0) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < 30; ++i) {
            const int s = hist[i];
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if ((float)max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if ((float)max3 < 0.1f * (float)max1) ind3 = -1;
        s_ind[0] = ind1; s_ind[1] = ind2; s_ind[2] = ind3;
    }
    __syncthreads();
    int removed = 0;
    for (int j = blockIdx.x * 256 + threadIdx.x; j < nf; j += gridDim.x * 256) {
        if (match[j] < 0) continue;
        const int b = bin_of[j];
        if (b != s_ind[0] && b != s_ind[1] && b != s_ind[2]) { match[j] = -1; ++removed; }
    }
    if (removed) atomicSub(nmatches, removed);
}

}  // namespace pslam

using namespace pslam;

extern "C" int pslam_search_by_bow(pslam_ctx* c, int nkf, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_has_mp, int kf_nodes,
                                   const int32_t* kf_node_id, const int32_t* kf_node_off, const int32_t* kf_node_feat, int nf, const uint8_t* f_desc,
                                   const float* f_angle, int f_nodes, const int32_t* f_node_id, const int32_t* f_node_off, const int32_t* f_node_feat,
                                   float nnratio, int check_orientation, int32_t* match) {
    if (!c) return PSLAM_E_INVALID;
    if (nkf < 0 || nf < 0 || kf_nodes < 0 || f_nodes < 0 || (nf && !match) || (nkf && (!kf_desc || !kf_angle || !kf_has_mp)) || (nf && (!f_desc || !f_angle)) ||
        (kf_nodes && (!kf_node_id || !kf_node_off || !kf_node_feat)) || (f_nodes && (!f_node_id || !f_node_off || !f_node_feat)))
        return set_error(c, PSLAM_E_INVALID, "bad SearchByBoW arrays");
    for (int i = 0; i < nf; ++i) match[i] = -1;
    // merge-join of the two (ascending) node-id lists, like the reference's two map iterators
    std::vector<int2> pairs;
    for (int a = 0, b = 0; a < kf_nodes && b < f_nodes;) {
        if (kf_node_id[a] == f_node_id[b]) { pairs.push_back(make_int2(a, b)); ++a; ++b; }
        else if (kf_node_id[a] < f_node_id[b]) a = (int)(std::lower_bound(kf_node_id, kf_node_id + kf_nodes, f_node_id[b]) - kf_node_id);
        else b = (int)(std::lower_bound(f_node_id, f_node_id + f_nodes, kf_node_id[a]) - f_node_id);
    }
    if (pairs.empty() || nf == 0 || nkf == 0) return 0;
    const int nkfeat = kf_node_off[kf_nodes], nffeat = f_node_off[f_nodes];
    for (int i = 0; i < nkfeat; ++i) if (kf_node_feat[i] < 0 || kf_node_feat[i] >= nkf) return set_error(c, PSLAM_E_INVALID, "key-frame feature index out of range");
    for (int i = 0; i < nffeat; ++i) if (f_node_feat[i] < 0 || f_node_feat[i] >= nf) return set_error(c, PSLAM_E_INVALID, "frame feature index out of range");
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = c->stream;
    const size_t sz[] = {pairs.size() * 8, (size_t)nkf * 32, (size_t)nkf * 4, (size_t)nkf, (size_t)(kf_nodes + 1) * 4, (size_t)nkfeat * 4, (size_t)nf * 32,
                         (size_t)nf * 4, (size_t)(f_nodes + 1) * 4, (size_t)nffeat * 4, (size_t)nf * 4, (size_t)nf, 31 * 4};
    const void* src[] = {pairs.data(), kf_desc, kf_angle, kf_has_mp, kf_node_off, kf_node_feat, f_desc, f_angle, f_node_off, f_node_feat, match, nullptr, nullptr};
    size_t off[14]; off[0] = 0;
    for (int i = 0; i < 13; ++i) off[i + 1] = (off[i] + sz[i] + 15) & ~(size_t)15;
    uint8_t* d = nullptr;
    PSLAM_CUDA(c, cudaMalloc((void**)&d, off[13]));
    cudaError_t e = cudaMemsetAsync(d + off[12], 0, 31 * 4, st);
    for (int i = 0; i < 11 && e == cudaSuccess; ++i) if (sz[i]) e = cudaMemcpyAsync(d + off[i], src[i], sz[i], cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFree(d); return check_cuda(c, e, "SearchByBoW upload"); }
    int32_t* d_hist = (int32_t*)(d + off[12]);
    PSLAM_LAUNCH(c, "bow_nodes", k_bow_nodes<<<((int)pairs.size() + 3) / 4, 128, 0, st>>>((int)pairs.size(), (const int2*)(d + off[0]), d + off[1], (const float*)(d + off[2]),
                 d + off[3], (const int32_t*)(d + off[4]), (const int32_t*)(d + off[5]), d + off[6], (const float*)(d + off[7]), (const int32_t*)(d + off[8]),
                 (const int32_t*)(d + off[9]), nnratio, check_orientation, (int32_t*)(d + off[10]), (int8_t*)(d + off[11]), d_hist, d_hist + 30));
    if (check_orientation)
        PSLAM_LAUNCH(c, "bow_orientation", k_bow_orientation<<<1, 256, 0, st>>>(nf, (int32_t*)(d + off[10]), (const int8_t*)(d + off[11]), d_hist, d_hist + 30));
    int32_t n = 0;
    e = cudaMemcpyAsync(match, d + off[10], (size_t)nf * 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(&n, d_hist + 30, 4, cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return check_cuda(c, e, "SearchByBoW");
    return n;
}
