// TEST INFRASTRUCTURE ONLY.  C entry points onto the REFERENCE's own plane extractor - PlaneDetection (src/PlaneExtractor.cpp,
// include/PlaneExtractor.h) driving the vendored PEAC (include/peac/*.hpp) - compiled unmodified from /root/reference with the
// container stand-ins of oracle/ref/shims/ (see the notes there: only the 3x3 eigen-solver is not the reference's code).
// Built into oracle/_ref/libpeac_ref.so by `make -C oracle ref`; used by tests/test_oracle_peac_ref.py to pin oracle/peac.cc.
#include <cstdint>
#include <cstring>

#include "PlaneExtractor.h"

struct RefPeac { PlaneDetection pd; };

extern "C" {
// Frame::ComputePlanes lines src/Frame.cc:648-650: readDepthImage(Depth, K, depthMapFactor); runPlaneDetection(rows, cols)
void* ref_peac_run(const uint16_t* depth, int w, int h, float fx, float fy, float cx, float cy, float scale) {
    RefPeac* r = new RefPeac();
    cv::Mat d(h, w, CV_16UC1, (void*)depth);
    cv::Mat K(3, 3, CV_32F);
    K.setTo(0);
    K.at<float>(0, 0) = fx; K.at<float>(1, 1) = fy; K.at<float>(0, 2) = cx; K.at<float>(1, 2) = cy; K.at<float>(2, 2) = 1;
    r->pd.readDepthImage(d, K, scale);
    r->pd.runPlaneDetection(h, w);
    return r;
}
void ref_peac_free(void* p) { delete (RefPeac*)p; }
int ref_peac_num_planes(void* p) { return ((RefPeac*)p)->pd.plane_num_; }
void ref_peac_labels(void* p, int32_t* out) {
    RefPeac* r = (RefPeac*)p;
    cv::Mat& m = r->pd.plane_filter.membershipImg;
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) out[(size_t)i * m.cols + j] = m.at<int>(i, j);
}
// normal[3], center[3], mse, curvature; N
void ref_peac_plane(void* p, int i, double* d8, int* n) {
    auto& q = ((RefPeac*)p)->pd.plane_filter.extractedPlanes[i];
    for (int k = 0; k < 3; ++k) { d8[k] = q->normal[k]; d8[3 + k] = q->center[k]; }
    d8[6] = q->mse; d8[7] = q->curvature; *n = q->N;
}
int ref_peac_membership(void* p, int i, int32_t* out, int cap) {
    const auto& m = ((RefPeac*)p)->pd.plane_vertices_[i];
    for (int k = 0; k < (int)m.size() && k < cap; ++k) out[k] = m[k];
    return (int)m.size();
}
}
