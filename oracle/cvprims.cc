// TEST INFRASTRUCTURE ONLY — see cvprims.h. OpenCV primitives restated from OpenCV's published algorithms
// (imgproc resize / smooth / features2d FAST / core fastAtan2); pinned against cv2 4.13 by the tests.
#include "cvprims.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace oracle {

int cv_round(double v) { return (int)std::nearbyint(v); }  // default FE_TONEAREST = half-to-even

int reflect101(int i, int n) {
    if (n == 1) return 0;
    while (i < 0 || i >= n) i = (i < 0) ? -i : 2 * (n - 1) - i;
    return i;
}

static inline short sat_short(float v) {
    int r = cv_round(v);
    return (short)std::min(32767, std::max(-32768, r));
}

// cv::resize, INTER_LINEAR, CV_8UC1: 11-bit fixed-point coefficients (INTER_RESIZE_COEF_BITS = 11),
// horizontal pass into int, vertical pass ((b*(S>>4))>>16 ... +2)>>2.
void resize_linear_u8(const Img8& s, uint8_t* dst, int dw, int dh) {
    const double scale_x = (double)s.w / dw, scale_y = (double)s.h / dh;
    std::vector<int> xofs(dw), yofs(dh);
    std::vector<short> alpha(2 * dw), beta(2 * dh);
    for (int dx = 0; dx < dw; ++dx) {
        float fx = (float)((dx + 0.5) * scale_x - 0.5);
        int sx = (int)std::floor(fx);
        fx -= sx;
        if (sx < 0) { fx = 0; sx = 0; }
        if (sx >= s.w - 1) { fx = 0; sx = s.w - 1; }
        xofs[dx] = sx;
        alpha[2 * dx] = sat_short((1.f - fx) * 2048.f);
        alpha[2 * dx + 1] = sat_short(fx * 2048.f);
    }
    for (int dy = 0; dy < dh; ++dy) {
        float fy = (float)((dy + 0.5) * scale_y - 0.5);
        int sy = (int)std::floor(fy);
        fy -= sy;
        yofs[dy] = sy;
        beta[2 * dy] = sat_short((1.f - fy) * 2048.f);
        beta[2 * dy + 1] = sat_short(fy * 2048.f);
    }
    std::vector<int> r0(dw), r1(dw);
    auto hrow = [&](int sy, std::vector<int>& out) {
        sy = std::min(std::max(sy, 0), s.h - 1);
        const uint8_t* S = s.p + (size_t)sy * s.stride;
        for (int dx = 0; dx < dw; ++dx) {
            int sx = xofs[dx];
            int sx1 = std::min(sx + 1, s.w - 1);
            out[dx] = S[sx] * alpha[2 * dx] + S[sx1] * alpha[2 * dx + 1];
        }
    };
    for (int dy = 0; dy < dh; ++dy) {
        hrow(yofs[dy], r0);
        hrow(yofs[dy] + 1, r1);
        const int b0 = beta[2 * dy], b1 = beta[2 * dy + 1];
        uint8_t* D = dst + (size_t)dy * dw;
        for (int dx = 0; dx < dw; ++dx) {
            int v = (((b0 * (r0[dx] >> 4)) >> 16) + ((b1 * (r1[dx] >> 4)) >> 16) + 2) >> 2;
            D[dx] = (uint8_t)std::min(255, std::max(0, v));
        }
    }
}

void copy_make_border_reflect101(const Img8& s, uint8_t* dst, int b) {
    const int W = s.w + 2 * b, H = s.h + 2 * b;
    for (int y = 0; y < H; ++y) {
        int sy = reflect101(y - b, s.h);
        for (int x = 0; x < W; ++x) dst[(size_t)y * W + x] = s.at(sy, reflect101(x - b, s.w));
    }
}

// cv::GaussianBlur(Size(7,7), 2, 2, BORDER_REFLECT_101) on CV_8UC1: OpenCV's bit-exact fixed-point path.
// Kernel in unsigned 8.8 fixed point (sums to 256); horizontal pass keeps 8.8, vertical pass accumulates
// 16.16 and rounds half-up.  The tap values are what getGaussianKernel's fixed-point variant yields for
// n=7, sigma=2 (pinned by tests/test_oracle_cvprims.py against cv2).
const int kGauss7S2[7] = {18, 34, 48, 56, 48, 34, 18};

void gaussian_blur_7x7_s2_u8(const Img8& s, uint8_t* dst) {
    const int w = s.w, h = s.h;
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int k = -3; k <= 3; ++k) acc += kGauss7S2[k + 3] * s.at(y, reflect101(x + k, w));
            tmp[(size_t)y * w + x] = (uint16_t)acc;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = -3; k <= 3; ++k) acc += (uint32_t)kGauss7S2[k + 3] * tmp[(size_t)reflect101(y + k, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)std::min<uint32_t>(255u, (acc + 32768u) >> 16);
        }
}

// Bresenham circle of radius 3, clockwise from (0,3).
static const int kCircle[16][2] = {{0, 3},  {1, 3},   {2, 2},   {3, 1},   {3, 0},  {3, -1}, {2, -2}, {1, -3},
                                   {0, -3}, {-1, -3}, {-2, -2}, {-3, -1}, {-3, 0}, {-3, 1}, {-2, 2}, {-1, 3}};

// M-1 where M = max over the 16 arcs of 9 contiguous circle pixels and both polarities of the smallest
// signed difference; the pixel is a FAST-9 corner at threshold t  <=>  M > t  <=>  score >= t.
// Returns <= 0 when there is no arc with a uniform sign (never a corner for any t >= 0).
int fast_score_9_16(const Img8& im, int x, int y) {
    int d[16];
    const int v = im.at(y, x);
    for (int k = 0; k < 16; ++k) d[k] = v - im.at(y + kCircle[k][1], x + kCircle[k][0]);
    int best = 0;
    for (int s0 = 0; s0 < 16; ++s0) {
        int mn = 255, mx = -255;
        for (int k = 0; k < 9; ++k) {
            int dv = d[(s0 + k) & 15];
            mn = std::min(mn, dv);
            mx = std::max(mx, dv);
        }
        best = std::max(best, std::max(mn, -mx));
    }
    return best - 1;
}

void fast_detect(const Img8& sub, int thr, std::vector<FastKp>& out) {
    out.clear();
    const int w = sub.w, h = sub.h;
    if (w < 7 || h < 7) return;
    std::vector<int> sc((size_t)w * h, 0);
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            // exact early-out: any arc of 9 of the 16 circle pixels holds at least two of the four compass points, so a corner at threshold
            // thr needs two compass points darker than v - thr or two brighter than v + thr (the high-speed test of the published FAST)
            const int v = sub.at(y, x);
            int nd = 0, nb = 0;
            for (int k = 0; k < 16; k += 4) { const int dv = v - sub.at(y + kCircle[k][1], x + kCircle[k][0]); nd += dv > thr; nb += -dv > thr; }
            if (nd < 2 && nb < 2) continue;
            int s = fast_score_9_16(sub, x, y);
            sc[(size_t)y * w + x] = (s >= thr) ? s : 0;
        }
    for (int y = 3; y < h - 3; ++y)
        for (int x = 3; x < w - 3; ++x) {
            int s = sc[(size_t)y * w + x];
            if (s <= 0) continue;
            bool mx = true;
            for (int dy = -1; dy <= 1 && mx; ++dy)
                for (int dx = -1; dx <= 1; ++dx)
                    if ((dx || dy) && sc[(size_t)(y + dy) * w + (x + dx)] >= s) { mx = false; break; }
            if (mx) out.push_back({x, y, s});
        }
}

// cv::fastAtan2 (scalar path): degree-valued 7th-order odd polynomial on the min/max ratio.
float fast_atan2_deg(float y, float x) {
    static const float p1 = 0.9997878412794807f * (float)(180 / M_PI);
    static const float p3 = -0.3258083974640975f * (float)(180 / M_PI);
    static const float p5 = 0.1555786518463281f * (float)(180 / M_PI);
    static const float p7 = -0.04432655554792128f * (float)(180 / M_PI);
    float ax = std::fabs(x), ay = std::fabs(y), a, c, c2;
    if (ax >= ay) {
        c = ay / (ax + (float)DBL_EPSILON);
        c2 = c * c;
        a = (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    } else {
        c = ax / (ay + (float)DBL_EPSILON);
        c2 = c * c;
        a = 90.f - (((p7 * c2 + p5) * c2 + p3) * c2 + p1) * c;
    }
    if (x < 0) a = 180.f - a;
    if (y < 0) a = 360.f - a;
    return a;
}

}  // namespace oracle
