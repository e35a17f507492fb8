// TEST INFRASTRUCTURE ONLY - CPU oracle (see cvprims.h header note).
// Restatement of the LBD line descriptor the reference computes per frame:
//   LineSegment::ExtractLineSegment    src/LSDextractor.cpp:14,28   lbd = BinaryDescriptor::createBinaryDescriptor(); lbd->compute(img, keylines, ldesc)
//   cv::line_descriptor::BinaryDescriptor::compute (opencv_contrib 3.4.1 line_descriptor, NOT in /root/reference and not in the cv2 of this image):
//     octave image = GaussianBlur(image, 5x5, sigma 1); dx / dy = Sobel(CV_16S, ksize 3); per key line the 9-band x 7-row line-support region is walked
//     along the line direction, the gradients are projected on the line / its normal, positive and negative parts are summed per row with the global
//     Gaussian weight, spread over the band and its two neighbours with the local Gaussian weight; per band mean and standard deviation of the four sums
//     -> 72 floats, normalised (means and deviations separately), clipped at 0.4, renormalised; 32 bytes = 8 comparisons for each of 32 band pairs
//     (Zhang & Koch, "An efficient and robust line segment matching approach based on LBD descriptor and pairwise geometric consistency", JVCI 2013).
// PARITY UNPINNED for the descriptor logic: the upstream source is absent and no binary of it exists here; this file follows the published algorithm and the
// structure of the upstream implementation as recalled (row / band accumulation order, the integer-division sigmas of the two Gaussian windows, the 32 band
// pairs).  What IS pinned: its two OpenCV primitives, GaussianBlur(5x5, s=1) on 8-bit data and Sobel(CV_16S, 3), bit-for-bit against cv2 4.13
// (tests/test_oracle_lbd.py).  The CUDA path is held bit-exact to this restatement.
#pragma once
#include <cstdint>
#include <vector>

#include "cvprims.h"
#include "lsd.h"

namespace oracle {

// cv::GaussianBlur(src, dst, Size(5, 5), 1) on CV_8U (8.8 fixed point, taps 14 62 104 62 14, REFLECT_101)
void gaussian_blur_5x5_s1_u8(const Img8& src, uint8_t* dst /* stride = src.w */);
// cv::Sobel(src, dst, CV_16S, dx, dy, 3) (REFLECT_101)
void sobel3_s16(const uint8_t* src, int w, int h, int16_t* dx, int16_t* dy);
// the 72-float LBD vector and the 32-byte binary descriptor of every key line (octave 0)
void lbd_compute(const Img8& img, const KeyLine* kl, int n, float* lbd72 /* [n][72], may be null */, uint8_t* desc /* [n][32] */);

}  // namespace oracle
