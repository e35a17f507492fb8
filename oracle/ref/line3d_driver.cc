// TEST INFRASTRUCTURE ONLY.  C entry point onto the REFERENCE's own 3-D line fit - compPt3dCov, extract3dline_mahdist, verify3dLine,
// mah_dist3d_pt_line, computeLine3d_svd (src/LineExtractor.cpp, compiled unmodified from /root/reference with the stand-ins of
// oracle/ref/shims/: cv::Mat algebra with cv::gemm's summation order, cv::SVD = OpenCV's Jacobi algorithm of oracle/cvsvd.h) and libc's rand().
// The per-line loop around them - Frame::isLineGood (src/Frame.cc:189-267: sampling, depth look-up, back-projection, accept test) - lives in a
// translation unit that cannot be compiled here (Frame.h needs PCL / g2o / DBoW2 / the whole object graph) and is restated below, clearly
// separated.  Built into oracle/_ref/libline3d_ref.so by `make -C oracle ref`; pins oracle/line3d.cc (tests/test_oracle_line3d_ref.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "LSDextractor.h"

RandomPoint3d compPt3dCov(cv::Point3d pt, cv::Mat K, double time_diff_sec);      // src/LineExtractor.cpp:1196
RandomLine3d extract3dline_mahdist(const std::vector<RandomPoint3d>& pts);       // src/LineExtractor.cpp:1265

extern "C" int ref_lines3d_frame(const void* keylines, int n_lines, const float* depth, int w, int h, const float* cam, uint32_t seed, int skip, uint8_t* valid,
                                 float* depth_line, double* lines3d, double* director, int32_t* n_points, int32_t* n_inliers, uint64_t* inliers) {
    const cv::line_descriptor::KeyLine* kl = (const cv::line_descriptor::KeyLine*)keylines;
    static_assert(sizeof(cv::line_descriptor::KeyLine) == 68, "KeyLine layout");
    const float fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3], invfx = 1.0f / fx, invfy = 1.0f / fy;
    cv::Mat K = (cv::Mat_<double>(3, 3) << fx, 0, cx, 0, fy, cy, 0, 0, 1);       // tmpK, src/Frame.cc:84-86
    srand(seed);
    for (int i = 0; i < skip; ++i) (void)rand();
    for (int i = 0; i < n_lines; ++i) {
        valid[i] = 0; depth_line[i] = -1.0f; n_points[i] = 0; n_inliers[i] = 0; inliers[i] = 0;
        for (int c = 0; c < 6; ++c) lines3d[6 * i + c] = 0;
        for (int c = 0; c < 3; ++c) director[3 * i + c] = 0;
        // ---- restated from Frame::isLineGood (src/Frame.cc:193-236): sampling and back-projection
        const cv::Point2f sp = kl[i].getStartPoint(), ep = kl[i].getEndPoint();
        const double len = cv::norm(sp - ep);
        const double numSmp = (double)std::min((int)len, 50);
        std::vector<cv::Point3d> pts3d;
        if (numSmp >= 1)
            for (int j = 0; j <= numSmp; ++j) {
                const cv::Point2f a = sp * (1 - j / numSmp), b = ep * (j / numSmp);
                const cv::Point2f s = a + b;
                const cv::Point2d pt(s.x, s.y);
                if (pt.x < 0 || pt.y < 0 || pt.x >= w || pt.y >= h) continue;
                int row, col;
                if ((floor(pt.x) == pt.x) && (floor(pt.y) == pt.y)) { col = std::max(int(pt.x - 1), 0); row = std::max(int(pt.y - 1), 0); }
                else { col = int(pt.x); row = int(pt.y); }
                const float d = depth[(size_t)row * w + col];
                if (d <= 0.01) continue;
                cv::Point3d p;
                p.z = d;
                p.x = (col - cx) * p.z * invfx;
                p.y = (row - cy) * p.z * invfy;
                pts3d.push_back(p);
            }
        n_points[i] = (int)pts3d.size();
        if (pts3d.size() < 10.0) continue;
        // ---- the reference's own code from here ...
        std::vector<RandomPoint3d> rndpts3d;
        rndpts3d.reserve(pts3d.size());
        for (size_t j = 0; j < pts3d.size(); ++j) rndpts3d.push_back(compPt3dCov(pts3d[j], K, 1));
        RandomLine3d tmpLine = extract3dline_mahdist(rndpts3d);
        // ---- ... to here; accept test restated from src/Frame.cc:246-264
        n_inliers[i] = (int)tmpLine.pts.size();
        for (const RandomPoint3d& q : tmpLine.pts)
            for (size_t j = 0; j < rndpts3d.size(); ++j)
                if (!((inliers[i] >> j) & 1) && q.pos.x == rndpts3d[j].pos.x && q.pos.y == rndpts3d[j].pos.y && q.pos.z == rndpts3d[j].pos.z) { inliers[i] |= (uint64_t)1 << j; break; }
        director[3 * i] = tmpLine.director.x; director[3 * i + 1] = tmpLine.director.y; director[3 * i + 2] = tmpLine.director.z;
        if (tmpLine.pts.size() / len > 0.4 && cv::norm(tmpLine.A - tmpLine.B) > 0.02) {
            valid[i] = 1;
            depth_line[i] = std::min(depth[(size_t)(int)kl[i].endPointY * w + (int)kl[i].endPointX], depth[(size_t)(int)kl[i].startPointY * w + (int)kl[i].startPointX]);
            lines3d[6 * i] = tmpLine.A.x; lines3d[6 * i + 1] = tmpLine.A.y; lines3d[6 * i + 2] = tmpLine.A.z;
            lines3d[6 * i + 3] = tmpLine.B.x; lines3d[6 * i + 4] = tmpLine.B.y; lines3d[6 * i + 5] = tmpLine.B.z;
        }
    }
    return 0;
}
