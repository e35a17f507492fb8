// TEST INFRASTRUCTURE ONLY — C entry points (ctypes) onto the oracle's OpenCV-primitive restatements.
#include "cvprims.h"
#include <cstring>
using namespace oracle;
extern "C" {
void orc_resize_linear_u8(const uint8_t* src, int sw, int sh, int sstride, uint8_t* dst, int dw, int dh) {
    resize_linear_u8(Img8{src, sw, sh, sstride}, dst, dw, dh);
}
void orc_border_reflect101(const uint8_t* src, int w, int h, int stride, uint8_t* dst, int b) {
    copy_make_border_reflect101(Img8{src, w, h, stride}, dst, b);
}
void orc_gaussian_blur_7x7_s2(const uint8_t* src, int w, int h, int stride, uint8_t* dst) {
    gaussian_blur_7x7_s2_u8(Img8{src, w, h, stride}, dst);
}
// returns number of keypoints; xys = int[3*cap] (x, y, score)
int orc_fast_detect(const uint8_t* src, int w, int h, int stride, int thr, int* xys, int cap) {
    std::vector<FastKp> kps;
    fast_detect(Img8{src, w, h, stride}, thr, kps);
    int n = (int)kps.size();
    for (int i = 0; i < n && i < cap; ++i) { xys[3*i] = kps[i].x; xys[3*i+1] = kps[i].y; xys[3*i+2] = kps[i].score; }
    return n;
}
float orc_fast_atan2(float y, float x) { return fast_atan2_deg(y, x); }
int orc_cv_round(double v) { return cv_round(v); }
}
