// TEST INFRASTRUCTURE ONLY — CPU oracle for the PlanarSLAM hot path. Nothing under planarslam_b200/ may
// include, link or call this; only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg do.
//
// Restatement of the OpenCV primitives the reference's ORB extractor calls (OpenCV is NOT vendored in
// /root/reference, so these follow OpenCV's published algorithms and are pinned bit-for-bit against the
// in-container cv2 4.13 by tests/test_oracle_cvprims.py):
//   cv::resize(INTER_LINEAR) u8      <- src/ORBextractor.cc:1120
//   cv::copyMakeBorder(REFLECT_101)  <- src/ORBextractor.cc:1122,1127
//   cv::GaussianBlur(7x7, s=2) u8    <- src/ORBextractor.cc:1086
//   cv::FAST(thr, nonmax=true)       <- src/ORBextractor.cc:809,814
//   cv::fastAtan2                    <- src/ORBextractor.cc:103
//   cvRound                          <- src/ORBextractor.cc:81,115,119,442,1112
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

struct Img8 {                       // borrowed view of an 8-bit single-channel image
    const uint8_t* p; int w, h, stride;
    uint8_t at(int y, int x) const { return p[(size_t)y * stride + x]; }
};

int cv_round(double v);             // round-half-to-even, like cvRound on SSE2 builds
int reflect101(int i, int n);       // gfedcb|abcdefgh|gfedcba

// dst must be dw*dh bytes (stride dw).
void resize_linear_u8(const Img8& src, uint8_t* dst, int dw, int dh);
// dst is (w+2b)x(h+2b), stride w+2b.
void copy_make_border_reflect101(const Img8& src, uint8_t* dst, int b);
// In-place-safe 7x7 sigma=2 Gaussian (OpenCV fixed-point 8.8 separable path), BORDER_REFLECT_101.
void gaussian_blur_7x7_s2_u8(const Img8& src, uint8_t* dst /*stride = src.w*/);

// FAST-9/16 corner score of pixel (x,y) (needs a 3-px margin): the largest t for which the pixel is
// still a corner at threshold t, or 0 when it is not a corner even at t = 1... (see .cc)
int fast_score_9_16(const Img8& im, int x, int y);

struct FastKp { int x, y, score; };
// cv::FAST(sub-image, thr, nonmaxSuppression=true, TYPE_9_16): row-major keypoints with coordinates
// relative to the sub-image, evaluated on its 3-px-inset interior only.
void fast_detect(const Img8& sub, int thr, std::vector<FastKp>& out);

float fast_atan2_deg(float y, float x);

}  // namespace oracle
