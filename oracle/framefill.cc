// TEST INFRASTRUCTURE ONLY — see framefill.h
#include "framefill.h"
#include <cstddef>
namespace oracle {
void compute_stereo_from_rgbd(int n, const float* keys, const float* keys_un, const float* depth, int w, float bf, float* u_right, float* out_depth) {
    for (int i = 0; i < n; ++i) {
        u_right[i] = -1; out_depth[i] = -1;
        const float u = keys[2 * i], v = keys[2 * i + 1];
        const float d = depth[(std::size_t)(int)v * w + (int)u];          // Mat::at<float>(float, float): truncation
        if (d > 0) { out_depth[i] = d; u_right[i] = keys_un[2 * i] - bf / d; }
    }
}
}  // namespace oracle
extern "C" void orc_compute_stereo_from_rgbd(int n, const float* keys, const float* keys_un, const float* depth, int w, float bf, float* u_right, float* out_depth) {
    oracle::compute_stereo_from_rgbd(n, keys, keys_un, depth, w, bf, u_right, out_depth);
}
