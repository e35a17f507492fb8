// Host build of planarslam_b200/csrc/linefrustum_body.h (the code k_lines_in_frustum runs per map line) for tests/test_linefrustum_host.py.
#include <cstdint>

#include "linefrustum_body.h"

// frame: Tcw[16], fx, fy, cx, cy, min_x, max_x, min_y, max_y, log_scale_factor (25 floats)
extern "C" int host_lines_in_frustum(const float* frame, int n, const double* pos, const double* normal, const float* max_distance, const float* min_distance,
                                     float cos_limit, uint8_t* in_view, float* proj, int32_t* level, float* view_cos) {
    LfFrame F;
    for (int i = 0; i < 16; ++i) F.Tcw[i] = frame[i];
    F.fx = frame[16]; F.fy = frame[17]; F.cx = frame[18]; F.cy = frame[19]; F.min_x = frame[20]; F.max_x = frame[21]; F.min_y = frame[22]; F.max_y = frame[23];
    F.log_scale_factor = frame[24];
    lf_camera_center(F);
    int cnt = 0;
    for (int k = 0; k < n; ++k) {
        const bool ok = lf_line_in_frustum(F, pos + 6 * k, normal + 3 * k, max_distance[k], min_distance[k], cos_limit, proj + 4 * k, level[k], view_cos[k]);
        in_view[k] = ok;
        cnt += ok;
    }
    return cnt;
}
