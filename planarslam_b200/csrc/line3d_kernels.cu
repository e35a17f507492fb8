// Per-frame 3-D line fit on sm_100a - Frame::isLineGood (src/Frame.cc:189-267) and the LineExtractor.cpp routines it calls.
// One thread per frame (line3d_body.h explains why the lines of a frame are a sequential chain); a batch of frames is one launch.
// The per-thread scratch (51 points x 96 B + a 3 x 51 SVD panel) lives in local memory.  Latency-bound, about 3-4 MFLOP per frame
// (2 040 3x3 Jacobi SVDs dominate); with thousands of frames in flight it is hidden behind the detector kernels.
#include <cstdint>

#include "line3d_body.h"
#include "pslam_internal.h"

namespace pslam {

// frames per block: the threads of a warp run different frames and diverge (RANSAC early exits, Jacobi sweep counts), so a warp
// carries only 8 frames; 2 368 frames -> 296 blocks, two per SM
#define L3D_BLOCK 8

static_assert(sizeof(L3dKeyLine) == sizeof(pslam_keyline), "KeyLine layout");
static_assert(sizeof(pslam_line3d) == 96, "pslam_line3d layout");

__global__ void __launch_bounds__(L3D_BLOCK) k_lines3d(const pslam_keyline* __restrict__ kl, const int32_t* __restrict__ n_lines, int max_lines,
                                                const uint16_t* __restrict__ depth, int nframes, L3dCam cam, const uint32_t* __restrict__ seed,
                                                const int32_t* __restrict__ skip, pslam_line3d* __restrict__ out, int32_t* __restrict__ n_drawn) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= nframes) return;
    L3dPoint pts[L3D_MAX_PTS];
    double At[3 * L3D_MAX_PTS];
    int n = n_lines[f];
    if (n < 0) n = 0;
    if (n > max_lines) n = max_lines;
    L3dRand rng;
    l3d_srand(rng, seed[f], skip ? skip[f] : 0);
    const uint16_t* dframe = depth + (size_t)f * cam.w * cam.h;
    for (int i = 0; i < max_lines; ++i) {
        pslam_line3d o;
        if (i < n) {
            L3dLineOut R;
            l3d_line(reinterpret_cast<const L3dKeyLine*>(kl)[(size_t)f * max_lines + i], dframe, cam, rng, pts, At, R);
            for (int c = 0; c < 3; ++c) { o.A[c] = R.A[c]; o.B[c] = R.B[c]; o.director[c] = R.director[c]; }
            o.inliers = R.inliers; o.depth = R.depth; o.n_points = R.n_points; o.n_inliers = R.n_inliers; o.valid = R.valid;
        } else {
            for (int c = 0; c < 3; ++c) { o.A[c] = 0; o.B[c] = 0; o.director[c] = 0; }
            o.inliers = 0; o.depth = -1.0f; o.n_points = 0; o.n_inliers = 0; o.valid = 0;
        }
        out[(size_t)f * max_lines + i] = o;
    }
    n_drawn[f] = rng.drawn;
}

static int lines3d_launch(pslam_ctx* c, const pslam_keyline* d_kl, const int32_t* d_nl, int max_lines, const uint16_t* d_depth, int nframes, float depth_factor,
                          const float* cam4, const uint32_t* d_seed, const int32_t* d_skip, pslam_line3d* d_out, int32_t* d_drawn) {
    L3dCam cam;
    cam.w = c->cfg.width; cam.h = c->cfg.height;
    cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    cam.invfx = 1.0f / cam.fx; cam.invfy = 1.0f / cam.fy;                      // src/Frame.cc:77-78
    cam.depth_factor = depth_factor;
    PSLAM_LAUNCH(c, "lines3d", k_lines3d<<<(nframes + L3D_BLOCK - 1) / L3D_BLOCK, L3D_BLOCK, 0, c->stream>>>(d_kl, d_nl, max_lines, d_depth, nframes, cam, d_seed, d_skip, d_out, d_drawn));
    return PSLAM_OK;
}

}  // namespace pslam

using namespace pslam;

extern "C" {

int pslam_lines3d_batch_dev(pslam_ctx* c, const pslam_keyline* keylines, const int32_t* n_lines, int max_lines, const uint16_t* depth, int nframes,
                            float depth_factor, const float* cam, const uint32_t* seed, const int32_t* skip, pslam_line3d* out, int32_t* n_drawn) {
    if (!c) return PSLAM_E_INVALID;
    if (!keylines || !n_lines || !depth || !cam || !seed || !out || !n_drawn || nframes < 1 || max_lines < 1 || cam[0] == 0 || cam[1] == 0)        // fy < 0 is legal (Examples/RGB-D/ICL.yaml:9)
        return set_error(c, PSLAM_E_INVALID, "bad lines3d arguments");
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    return lines3d_launch(c, keylines, n_lines, max_lines, depth, nframes, depth_factor, cam, seed, skip, out, n_drawn);
}

int pslam_lines3d_batch(pslam_ctx* c, const pslam_keyline* keylines, const int32_t* n_lines, int max_lines, const uint16_t* depth, int nframes, float depth_factor,
                        const float* cam, const uint32_t* seed, const int32_t* skip, pslam_line3d* out, int32_t* n_drawn) {
    if (!c) return PSLAM_E_INVALID;
    if (!keylines || !n_lines || !depth || !cam || !seed || !out || !n_drawn || nframes < 1 || max_lines < 1 || cam[0] == 0 || cam[1] == 0)        // fy < 0 is legal (Examples/RGB-D/ICL.yaml:9)
        return set_error(c, PSLAM_E_INVALID, "bad lines3d arguments");
    PSLAM_CUDA(c, cudaSetDevice(c->cfg.device));
    cudaStream_t st = c->stream;
    const size_t npx = (size_t)c->cfg.width * c->cfg.height;
    const size_t sz[] = {(size_t)nframes * max_lines * sizeof(pslam_keyline), (size_t)nframes * 4, (size_t)nframes * npx * 2, (size_t)nframes * 4, (size_t)nframes * 4,
                         (size_t)nframes * max_lines * sizeof(pslam_line3d), (size_t)nframes * 4};
    const void* src[] = {keylines, n_lines, depth, seed, skip, nullptr, nullptr};
    size_t off[8]; off[0] = 0;
    for (int i = 0; i < 7; ++i) off[i + 1] = (off[i] + sz[i] + 255) & ~(size_t)255;
    uint8_t* d = nullptr;
    PSLAM_CUDA(c, cudaMalloc((void**)&d, off[7]));
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 5 && e == cudaSuccess; ++i)
        if (src[i]) e = cudaMemcpyAsync(d + off[i], src[i], sz[i], cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { cudaFree(d); return check_cuda(c, e, "lines3d upload"); }
    const int rc = lines3d_launch(c, (const pslam_keyline*)(d + off[0]), (const int32_t*)(d + off[1]), max_lines, (const uint16_t*)(d + off[2]), nframes, depth_factor, cam,
                                  (const uint32_t*)(d + off[3]), skip ? (const int32_t*)(d + off[4]) : nullptr, (pslam_line3d*)(d + off[5]), (int32_t*)(d + off[6]));
    if (rc != PSLAM_OK) { cudaFree(d); return rc; }
    e = cudaMemcpyAsync(out, d + off[5], sz[5], cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(n_drawn, d + off[6], sz[6], cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    cudaFree(d);
    if (e != cudaSuccess) return check_cuda(c, e, "lines3d");
    return PSLAM_OK;
}

}
