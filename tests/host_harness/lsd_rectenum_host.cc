// Host build of planarslam_b200/csrc/lsd_rectenum.h (the header the CUDA NFA validation uses for the OpenCV 4.x rectangle
// enumeration), for tests/test_lsd_rectenum_host.py: g++ compiles the very source nvcc compiles for the device.
#include <cstdint>

#include "lsd_rectenum.h"

// rect: x1 y1 x2 y2 width dx dy.  Writes up to cap rows {y, xa, xb} (clamped to the W x H image, empty rows skipped) and returns
// the number of rows the scan visits inside the image with a non-empty span.
extern "C" int host_lsd_cv4_spans(const double* rect, int W, int H, int32_t* rows, int cap) {
    LsdRowScan S;
    lsd_cv4_setup(rect[0], rect[1], rect[2], rect[3], rect[4], rect[5], rect[6], S);
    int m = 0;
    const int ya = S.y0 < 0 ? 0 : S.y0, yb = S.c2 < H - 1 ? S.c2 : H - 1;
    for (int y = ya; y <= yb; ++y) {
        int xa, xb;
        lsd_cv4_row(S, y, xa, xb);
        if (xa < 0) xa = 0;
        if (xb > W - 1) xb = W - 1;
        if (xb < xa) continue;
        if (m < cap) { rows[3 * m] = y; rows[3 * m + 1] = xa; rows[3 * m + 2] = xb; }
        ++m;
    }
    return m;
}

// planarslam_b200/csrc/lsd_detsincos.h (the deterministic sincos the device code and the host-built seed table share), compiled for the host
#include "lsd_detsincos.h"
extern "C" void host_lsd_sincos(const double* x, int n, double* s, double* c) {
    for (int i = 0; i < n; ++i) lsd_sincos_body(x[i], s[i], c[i]);
}
