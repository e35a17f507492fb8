// Host build of planarslam_b200/csrc/line3d_body.h (the code k_lines3d runs, one thread per frame) for tests/test_line3d_host.py.
#include <cstdint>
#include <cstring>

#include "line3d_body.h"

// One frame.  Output arrays sized n_lines: lines3d f64[6], director f64[3], inliers u64, depth f32, n_points / n_inliers / valid i32.
extern "C" int host_lines3d_frame(const void* keylines, int n_lines, const uint16_t* depth16, int w, int h, float depth_factor, const float* cam4, uint32_t seed,
                                  int skip, double* lines3d, double* director, uint64_t* inliers, float* depth_line, int32_t* n_points, int32_t* n_inliers,
                                  int32_t* valid) {
    L3dCam cam;
    cam.w = w; cam.h = h; cam.fx = cam4[0]; cam.fy = cam4[1]; cam.cx = cam4[2]; cam.cy = cam4[3];
    cam.invfx = 1.0f / cam.fx; cam.invfy = 1.0f / cam.fy; cam.depth_factor = depth_factor;
    static L3dPoint pts[L3D_MAX_PTS];
    static double At[3 * L3D_MAX_PTS];
    L3dRand rng;
    l3d_srand(rng, seed, skip);
    const L3dKeyLine* kl = (const L3dKeyLine*)keylines;
    for (int i = 0; i < n_lines; ++i) {
        L3dLineOut R;
        l3d_line(kl[i], depth16, cam, rng, pts, At, R);
        std::memcpy(lines3d + 6 * i, R.A, 24); std::memcpy(lines3d + 6 * i + 3, R.B, 24); std::memcpy(director + 3 * i, R.director, 24);
        inliers[i] = R.inliers; depth_line[i] = R.depth; n_points[i] = R.n_points; n_inliers[i] = R.n_inliers; valid[i] = R.valid;
    }
    return rng.drawn;
}

extern "C" void host_glibc_rand(uint32_t seed, int skip, int n, int32_t* out) {
    L3dRand g;
    l3d_srand(g, seed, skip);
    for (int i = 0; i < n; ++i) out[i] = l3d_rand(g);
}
