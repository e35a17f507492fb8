// TEST INFRASTRUCTURE ONLY — C entry points (ctypes) onto the oracle 3-D line fit, glibc rand() and cv::SVD restatements.
#include <cstring>

#include "cvsvd.h"
#include "line3d.h"
using namespace oracle;
extern "C" {
void orc_glibc_rand(uint32_t seed, int n, int32_t* out) {
    GlibcRand g(seed);
    for (int i = 0; i < n; ++i) out[i] = g.rand();
}
// w: min(m,n); u: m x min(m,n); vt: min(m,n) x n
void orc_cv_svd64(const double* A, int m, int n, double* w, double* u, double* vt) { cv_svd<double>(A, m, n, w, u, vt); }
void orc_cv_svd32(const float* A, int m, int n, float* w, float* u, float* vt) { cv_svd<float>(A, m, n, w, u, vt); }

// One frame.  keylines: KeyLine[n_lines] (68 B); depth: float [h][w]; cam: fx fy cx cy; out arrays sized n_lines:
// valid u8, depth_line f32, lines3d f64[6], director f64[3], n_points i32, n_inliers i32, inliers u64.  Returns rand() calls made.
int orc_lines3d_frame(const void* keylines, int n_lines, const float* depth, int w, int h, const float* cam, uint32_t seed, int skip, uint8_t* valid,
                      float* depth_line, double* lines3d, double* director, int32_t* n_points, int32_t* n_inliers, uint64_t* inliers) {
    Line3dCam c;
    c.w = w; c.h = h; c.fx = cam[0]; c.fy = cam[1]; c.cx = cam[2]; c.cy = cam[3]; c.invfx = 1.0f / c.fx; c.invfy = 1.0f / c.fy;
    GlibcRand g(seed);
    for (int i = 0; i < skip; ++i) (void)g.rand();
    g.drawn = 0;
    std::vector<Line3dResult> r(n_lines);
    lines3d_frame((const KeyLine*)keylines, n_lines, depth, c, g, r.data());
    for (int i = 0; i < n_lines; ++i) {
        valid[i] = r[i].valid; depth_line[i] = r[i].depth;
        std::memcpy(lines3d + 6 * i, r[i].A, 24); std::memcpy(lines3d + 6 * i + 3, r[i].B, 24); std::memcpy(director + 3 * i, r[i].director, 24);
        n_points[i] = r[i].n_points; n_inliers[i] = r[i].n_inliers; inliers[i] = r[i].inliers;
    }
    return (int)g.drawn;
}
}

#include "manhattan.h"
// out_i: found[3], n_cone[3], n_selected[3], min_num, svd_applied (11 ints); out_f: R[9], density[3] (12 floats)
extern "C" void orc_track_manhattan_frame(const float* R_last, const float* normals, int n, const double* dirs, int m, int32_t* out_i, float* out_f,
                                          uint8_t* normal_mask, uint8_t* dir_mask) {
    oracle::ManhattanResult r;
    oracle::track_manhattan_frame(R_last, normals, n, dirs, m, r, normal_mask, dir_mask);
    for (int i = 0; i < 3; ++i) { out_i[i] = r.found[i]; out_i[3 + i] = r.n_cone[i]; out_i[6 + i] = r.n_selected[i]; out_f[9 + i] = r.density[i]; }
    out_i[9] = r.min_num; out_i[10] = r.svd_applied;
    for (int i = 0; i < 9; ++i) out_f[i] = r.R[i];
}
