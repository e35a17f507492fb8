// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the per-line 3-D fit the reference runs on every frame after the line detector:
//   Frame::isLineGood            src/Frame.cc:189-267          (<= 51 samples per 2-D line, depth look-up, back-projection, accept test)
//   compPt3dCov / depthStdDev    src/LineExtractor.cpp:1182-1248 (per-point covariance J diag(1,1,s^2) J^T, cv::SVD, D U^T)
//   extract3dline_mahdist        src/LineExtractor.cpp:1265-1359 (<= 10 RANSAC iterations on Mahalanobis distance, SVD refit loop, end points)
//   verify3dLine                 src/LineExtractor.cpp:1361-1415 (10-cell support test)
//   mah_dist3d_pt_line           src/LineExtractor.cpp:1418-1470
//   computeLine3d_svd            src/LineExtractor.cpp:1157-1179
//   random_unique                include/LSDextractor.h:239-251  (partial Fisher-Yates on libc rand())
// Third-party arithmetic: cv::SVD (oracle/cvsvd.h - pinned to cv2.SVDecomp to 1e-12 / sign only, see there) and glibc rand()
// (TYPE_3 additive feedback generator, restated below and PINNED bit-for-bit against the container's libc in
// tests/test_oracle_line3d.py).  The reference draws from the process-wide libc stream (unseeded, shared with other threads:
// src/MapPlane.cc:23-25); here every frame gets its own stream (seed, skip) - with the state the libc stream has at the start of
// the frame the draws are the reference's.
// PARITY UNPINNED beyond that: the reference ships no test vectors for this path and cannot be compiled here.
#pragma once
#include <cstdint>
#include <vector>

#include "lsd.h"

namespace oracle {

// glibc random_r TYPE_3 (what rand() uses): r[i] = r[i-3] + r[i-31], output >> 1
struct GlibcRand {
    int f = 3, b = 0;          // front / rear positions in the 31-word state
    int32_t state[31];
    uint64_t drawn = 0;
    explicit GlibcRand(uint32_t seed = 1) { srand(seed); }
    void srand(uint32_t seed);
    int32_t rand();
};

struct Line3dCam { int w, h; float fx, fy, cx, cy, invfx, invfy; };

struct Line3dResult {
    uint8_t valid = 0;         // the accept test of isLineGood passed
    float depth = -1.0f;       // mvDepthLine[i]
    double A[3] = {0, 0, 0}, B[3] = {0, 0, 0};       // mvLines3D[i]
    double director[3] = {0, 0, 0};                  // FrameLine::direction
    int32_t n_points = 0;      // samples with depth
    int32_t n_inliers = 0;     // RandomLine3d::pts.size()
    uint64_t inliers = 0;      // bit j: sample point j (in sampling order, after dropping samples without depth) supports the line
};

// depth: float metres [h][w] (the reference's imDepth after convertTo(CV_32F, mDepthMapFactor))
void lines3d_frame(const KeyLine* kl, int n_lines, const float* depth, const Line3dCam& cam, GlibcRand& rng, Line3dResult* out);

}  // namespace oracle
