// Row spans of the rectangle enumeration that cv2 4.x's LineSegmentDetectorImpl::rect_nfa visits (OpenCV imgproc lsd.cpp after the
// 4.5 rewrite; the reference reaches it through LSDDetector::detect, src/LSDextractor.cpp:16).  Plain double arithmetic only, so
// the same source gives the same spans compiled by nvcc for the device (--fmad=false) and by g++ for the host: the CPU suite
// checks this header, compiled for the host, against the oracle's independent statement of the enumeration
// (tests/test_lsd_rectenum_host.py) - the device code in lsd_kernels.cuh only adds the pixel loop.
//
// The enumeration: the four corners as doubles, rotated so that the first is the top one (smallest y, ties -> smallest x); rows
// from ceil(top.y) to ceil(bottom.y) INCLUSIVE; the left bound follows top -> v1 -> bottom and switches to the second edge AFTER
// row ceil(v1.y), the right bound follows top -> v3 -> bottom and switches AT row ceil(v3.y); an edge whose two ends round up to
// the same row has slope 0; columns from ceil(left) to trunc(right).
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define LSD_HD __host__ __device__ __forceinline__
#else
#define LSD_HD inline
#endif

struct LsdRowScan {
    double v0x, v0y, v1x, v1y, v3x, v3y;
    double s01, s12, s03, s32;
    int y0, c1, c2, c3;
};

// double -> int the way cvttsd2si does it (the reference is x86 code): out-of-range values give INT_MIN
LSD_HD int lsd_x86_trunc(double v) { return (v > -2147483649.0 && v < 2147483648.0) ? (int)v : (-2147483647 - 1); }

LSD_HD void lsd_cv4_setup(double x1, double y1, double x2, double y2, double width, double dx, double dy, LsdRowScan& S) {
    const double half = 0.5 * width, dyhw = dy * half, dxhw = half * dx;
    const double c0x = x1 - dyhw, c0y = y1 + dxhw, c1x = x2 - dyhw, c1y = y2 + dxhw;
    const double c2x = x2 + dyhw, c2y = y2 - dxhw, c3x = x1 + dyhw, c3y = y1 - dxhw;
    int off = 0;
    double mx = c0x, my = c0y;
    if (c1y < my || (c1y == my && mx > c1x)) { off = 1; mx = c1x; my = c1y; }
    if (c2y < my || (c2y == my && mx > c2x)) { off = 2; mx = c2x; my = c2y; }
    if (c3y < my || (c3y == my && mx > c3x)) { off = 3; mx = c3x; my = c3y; }
    double v2x, v2y;
    if (off == 0)      { S.v0x = c0x; S.v0y = c0y; S.v1x = c1x; S.v1y = c1y; v2x = c2x; v2y = c2y; S.v3x = c3x; S.v3y = c3y; }
    else if (off == 1) { S.v0x = c1x; S.v0y = c1y; S.v1x = c2x; S.v1y = c2y; v2x = c3x; v2y = c3y; S.v3x = c0x; S.v3y = c0y; }
    else if (off == 2) { S.v0x = c2x; S.v0y = c2y; S.v1x = c3x; S.v1y = c3y; v2x = c0x; v2y = c0y; S.v3x = c1x; S.v3y = c1y; }
    else               { S.v0x = c3x; S.v0y = c3y; S.v1x = c0x; S.v1y = c0y; v2x = c1x; v2y = c1y; S.v3x = c2x; S.v3y = c2y; }
    S.y0 = lsd_x86_trunc(ceil(S.v0y)); S.c1 = lsd_x86_trunc(ceil(S.v1y)); S.c2 = lsd_x86_trunc(ceil(v2y)); S.c3 = lsd_x86_trunc(ceil(S.v3y));
    S.s01 = S.c1 != S.y0 ? (S.v1x - S.v0x) / (S.v1y - S.v0y) : 0.0;
    S.s12 = S.c2 != S.c1 ? (v2x - S.v1x) / (v2y - S.v1y) : 0.0;
    S.s03 = S.c3 != S.y0 ? (S.v3x - S.v0x) / (S.v3y - S.v0y) : 0.0;
    S.s32 = S.c3 != S.c2 ? (v2x - S.v3x) / (v2y - S.v3y) : 0.0;
}

// columns [xa, xb] of row y (S.y0 <= y <= S.c2); empty when xb < xa.  Not clamped to the image.
LSD_HD void lsd_cv4_row(const LsdRowScan& S, int y, int& xa, int& xb) {
    const double yd = (double)y;
    const double left = y > S.c1 ? (yd - S.v1y) * S.s12 + S.v1x : (yd - S.v0y) * S.s01 + S.v0x;
    const double right = y >= S.c3 ? (yd - S.v3y) * S.s32 + S.v3x : (yd - S.v0y) * S.s03 + S.v0x;
    xa = lsd_x86_trunc(ceil(left));
    xb = lsd_x86_trunc(right);
}
