// Manhattan-frame tracking step on sm_100a - Tracking::TrackManhattanFrame (src/Tracking.cc:763-1157).
// One thread per frame, 8 frames per block (manhattan_body.h: six ordered passes over the frame's surface normals; the sums keep the
// reference's order).  ~0.1 MB of normals per frame are streamed six times from L2; latency-bound, hidden behind the detector kernels
// when thousands of frames are in flight.  Round 2: a warp per frame with an ordered tree reduction once it can be timed.
#include <cstdint>

#include "manhattan_body.h"
#include "pslam_internal.h"

namespace pslam {

#define MH_BLOCK 8
static_assert(sizeof(MhResult) == sizeof(pslam_manhattan_result) && sizeof(MhResult) == 92, "pslam_manhattan_result layout");

__global__ void __launch_bounds__(MH_BLOCK) k_track_manhattan(const float* __restrict__ R_last, const float* __restrict__ normals, const int32_t* __restrict__ n_normals,
                                                              int max_normals, const double* __restrict__ dirs, const int32_t* __restrict__ n_dirs, int max_dirs,
                                                              int nframes, pslam_manhattan_result* __restrict__ res, uint8_t* __restrict__ nmask,
                                                              uint8_t* __restrict__ dmask) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    int n = n_normals[f], m = n_dirs[f];
    n = n < 0 ? 0 : (n > max_normals ? max_normals : n);
    m = m < 0 ? 0 : (m > max_dirs ? max_dirs : m);
    MhResult r;
    mh_track(R_last + 9 * (size_t)f, normals + 3 * (size_t)f * max_normals, n, dirs + 3 * (size_t)f * max_dirs, m, r, nmask + (size_t)f * max_normals,
             dmask + (size_t)f * max_dirs);
    for (int i = n; i < max_normals; ++i) nmask[(size_t)f * max_normals + i] = 0;
    for (int i = m; i < max_dirs; ++i) dmask[(size_t)f * max_dirs + i] = 0;
    *reinterpret_cast<MhResult*>(res + f) = r;
}

}  // namespace pslam

using namespace pslam;

extern "C" {

int pslam_track_manhattan_batch_dev(pslam_ctx* c, const float* R_last, const float* normals, const int32_t* n_normals, int max_normals, const double* dirs,
                                    const int32_t* n_dirs, int max_dirs, int nframes, pslam_manhattan_result* res, uint8_t* normal_mask, uint8_t* dir_mask) {
    if (!c) return PSLAM_E_INVALID;
    if (!R_last || !normals || !n_normals || !dirs || !n_dirs || !res || !normal_mask || !dir_mask || nframes < 1 || max_normals < 1 || max_dirs < 1)
        return set_error(c, PSLAM_E_INVALID, "bad manhattan arguments");
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    PSLAM_LAUNCH(c, "track_manhattan", k_track_manhattan<<<(nframes + MH_BLOCK - 1) / MH_BLOCK, MH_BLOCK, 0, c->stream>>>(R_last, normals, n_normals, max_normals, dirs,
                 n_dirs, max_dirs, nframes, res, normal_mask, dir_mask));
    return PSLAM_OK;
}

int pslam_track_manhattan_batch(pslam_ctx* c, const float* R_last, const float* normals, const int32_t* n_normals, int max_normals, const double* dirs,
                                const int32_t* n_dirs, int max_dirs, int nframes, pslam_manhattan_result* res, uint8_t* normal_mask, uint8_t* dir_mask) {
    if (!c) return PSLAM_E_INVALID;
    if (!R_last || !normals || !n_normals || !dirs || !n_dirs || !res || !normal_mask || !dir_mask || nframes < 1 || max_normals < 1 || max_dirs < 1)
        return set_error(c, PSLAM_E_INVALID, "bad manhattan arguments");
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = c->stream;
    const size_t nf = (size_t)nframes;
    const size_t sz[] = {nf * 36, nf * max_normals * 12, nf * 4, nf * max_dirs * 24, nf * 4, nf * sizeof(pslam_manhattan_result), nf * max_normals, nf * max_dirs};
    const void* src[] = {R_last, normals, n_normals, dirs, n_dirs, nullptr, nullptr, nullptr};
    void* dst[] = {nullptr, nullptr, nullptr, nullptr, nullptr, res, normal_mask, dir_mask};
    size_t off[9]; off[0] = 0;
    for (int i = 0; i < 8; ++i) off[i + 1] = (off[i] + sz[i] + 255) & ~(size_t)255;
    uint8_t* d = nullptr;
    PSLAM_CUDA(c, cudaMalloc((void**)&d, off[8]));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 5 && e == cudaSuccess; ++i) e = cudaMemcpyAsync(d + off[i], src[i], sz[i], cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFree(d); return check_cuda(c, e, "manhattan upload"); }
    const int rc = pslam_track_manhattan_batch_dev(c, (const float*)(d + off[0]), (const float*)(d + off[1]), (const int32_t*)(d + off[2]), max_normals,
                                                   (const double*)(d + off[3]), (const int32_t*)(d + off[4]), max_dirs, nframes, (pslam_manhattan_result*)(d + off[5]),
                                                   d + off[6], d + off[7]);
    if (rc != PSLAM_OK) { cudaFree(d); return rc; }
    for (int i = 5; i < 8 && e == cudaSuccess; ++i) e = cudaMemcpyAsync(dst[i], d + off[i], sz[i], cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return check_cuda(c, e, "manhattan");
    return PSLAM_OK;
}

}
