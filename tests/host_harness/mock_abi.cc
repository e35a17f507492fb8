// TEST INFRASTRUCTURE ONLY.  CPU stand-in for the entry points of libpslam_b200.so that were added after the round-1 GPU budget was
// spent, built from the SAME shared host/device bodies the CUDA kernels call (planarslam_b200/csrc/*_body.h).  It includes
// include/pslam_abi.h, so the signatures are checked against the real ABI at compile time.  tests/test_new_kernels_mock_abi.py loads
// it in place of the product library to run the Python mirrors and the GPU test bodies end to end on the CPU (argument order, array
// shapes, dtypes).  It is never shipped and nothing under planarslam_b200/ refers to it; the product still has no CPU path.
#include <cstdint>
#include <cstring>
#include <vector>

#include "linefrustum_body.h"
#include "line3d_body.h"
#include "manhattan_body.h"
#include "pslam_abi.h"

struct pslam_ctx { int width, height; };

extern "C" {

pslam_ctx* mock_create(int width, int height) { return new pslam_ctx{width, height}; }
void mock_destroy(pslam_ctx* c) { delete c; }

int pslam_lines3d_batch(pslam_ctx* c, const pslam_keyline* keylines, const int32_t* n_lines, int max_lines, const uint16_t* depth, int nframes, float depth_factor,
                        const float* cam, const uint32_t* seed, const int32_t* skip, pslam_line3d* out, int32_t* n_drawn) {
    static_assert(sizeof(L3dKeyLine) == sizeof(pslam_keyline), "KeyLine layout");
    L3dCam cm;
    cm.w = c->width; cm.h = c->height; cm.fx = cam[0]; cm.fy = cam[1]; cm.cx = cam[2]; cm.cy = cam[3];
    cm.invfx = 1.0f / cm.fx; cm.invfy = 1.0f / cm.fy; cm.depth_factor = depth_factor;
    std::vector<L3dPoint> pts(L3D_MAX_PTS);
    std::vector<double> At(3 * L3D_MAX_PTS);
    for (int f = 0; f < nframes; ++f) {
        int n = n_lines[f];
        n = n < 0 ? 0 : (n > max_lines ? max_lines : n);
        L3dRand rng;
        l3d_srand(rng, seed[f], skip ? skip[f] : 0);
        const uint16_t* d = depth + (size_t)f * cm.w * cm.h;
        for (int i = 0; i < max_lines; ++i) {
            pslam_line3d o;
            std::memset(&o, 0, sizeof(o));
            o.depth = -1.0f;
            if (i < n) {
                L3dLineOut R;
                l3d_line(reinterpret_cast<const L3dKeyLine*>(keylines)[(size_t)f * max_lines + i], d, cm, rng, pts.data(), At.data(), R);
                for (int k = 0; k < 3; ++k) { o.A[k] = R.A[k]; o.B[k] = R.B[k]; o.director[k] = R.director[k]; }
                o.inliers = R.inliers; o.depth = R.depth; o.n_points = R.n_points; o.n_inliers = R.n_inliers; o.valid = R.valid;
            }
            out[(size_t)f * max_lines + i] = o;
        }
        n_drawn[f] = rng.drawn;
    }
    return PSLAM_OK;
}

int pslam_track_manhattan_batch(pslam_ctx*, const float* R_last, const float* normals, const int32_t* n_normals, int max_normals, const double* dirs,
                                const int32_t* n_dirs, int max_dirs, int nframes, pslam_manhattan_result* res, uint8_t* normal_mask, uint8_t* dir_mask) {
    static_assert(sizeof(MhResult) == sizeof(pslam_manhattan_result), "result layout");
    for (int f = 0; f < nframes; ++f) {
        std::memset(normal_mask + (size_t)f * max_normals, 0, max_normals);
        std::memset(dir_mask + (size_t)f * max_dirs, 0, max_dirs);
        mh_track(R_last + 9 * (size_t)f, normals + 3 * (size_t)f * max_normals, n_normals[f], dirs + 3 * (size_t)f * max_dirs, n_dirs[f],
                 *reinterpret_cast<MhResult*>(res + f), normal_mask + (size_t)f * max_normals, dir_mask + (size_t)f * max_dirs);
    }
    return PSLAM_OK;
}

int pslam_lines_in_frustum(pslam_ctx*, const pslam_line_frustum_frame* frame, int n, const double* pos, const double* normal, const float* max_distance,
                           const float* min_distance, float cos_limit, uint8_t* in_view, float* proj, int32_t* level, float* view_cos) {
    LfFrame F;
    for (int i = 0; i < 16; ++i) F.Tcw[i] = frame->Tcw[i];
    F.fx = frame->fx; F.fy = frame->fy; F.cx = frame->cx; F.cy = frame->cy; F.min_x = frame->min_x; F.max_x = frame->max_x; F.min_y = frame->min_y; F.max_y = frame->max_y;
    F.log_scale_factor = frame->log_scale_factor;
    lf_camera_center(F);
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        const bool ok = lf_line_in_frustum(F, pos + 6 * (size_t)k, normal + 3 * (size_t)k, max_distance[k], min_distance[k], cos_limit, proj + 4 * (size_t)k, level[k], view_cos[k]);
        in_view[k] = ok;
        cnt += ok;
    }
    return cnt;
}

int pslam_compute_stereo_from_rgbd_batch(pslam_ctx* c, const pslam_keypoint* keys, const pslam_keypoint* keys_un, const int32_t* n, int cap, const uint16_t* depth,
                                         int nframes, float depth_factor, float bf, float* u_right, float* depth_out) {
    for (int f = 0; f < nframes; ++f)
        for (int i = 0; i < cap; ++i) {
            const size_t o = (size_t)f * cap + i;
            float ur = -1.0f, dz = -1.0f;
            if (i < n[f]) {
                const float d = (float)depth[((size_t)f * c->height + (int)keys[o].y) * c->width + (int)keys[o].x] * depth_factor;
                if (d > 0) { dz = d; ur = keys_un[o].x - bf / d; }
            }
            u_right[o] = ur; depth_out[o] = dz;
        }
    return PSLAM_OK;
}

const char* pslam_last_error(const pslam_ctx*) { return "mock"; }

}
