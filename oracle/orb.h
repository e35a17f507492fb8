// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the reference ORB extractor, src/ORBextractor.cc (ctor :410-470, ComputePyramid :1107-1132,
// ComputeKeyPointsOctTree :765-853, DistributeOctTree :539-763, IC_Angle :77-104, computeOrbDescriptor
// :108-147, operator() :1043-1105).
//
// Parity conventions fixed here (the reference leaves them to the allocator / compiler; DESIGN.md lists them):
//  * quadtree "largest node first" ties (reference sorts pair<count, node*>, :684) are broken by node
//    creation order (a later-created node counts as the larger pointer);
//  * the steered-BRIEF coordinate expressions x*b + y*a and x*a - y*b are evaluated WITHOUT fused
//    multiply-add (as a stock x86-64 OpenCV build does).
#pragma once
#include <cstdint>
#include <vector>

#include "cvprims.h"

namespace oracle {

struct KeyPoint {  // layout-compatible with cv::KeyPoint (28 bytes)
    float x, y, size, angle, response;
    int octave, class_id;
};

struct OrbParams {
    int nfeatures = 1000;
    float scale_factor = 1.2f;
    int nlevels = 8;
    int ini_th_fast = 20;
    int min_th_fast = 7;
};

struct Candidate { int x, y, score; };  // relative to (minBorderX, minBorderY) = (16,16), like :822-823

class OrbExtractor {
public:
    explicit OrbExtractor(const OrbParams& p);
    // Full extraction. desc is N x 32 bytes.
    void extract(const Img8& gray, std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc);

    // ---- stage outputs kept for stage-by-stage parity tests ----
    struct Level { int w = 0, h = 0; std::vector<uint8_t> px; };  // borderless
    std::vector<Level> pyramid;                        // after extract()/compute_pyramid()
    std::vector<std::vector<Candidate>> candidates;    // per level, cell-major / row-major order
    std::vector<std::vector<KeyPoint>> level_kps;      // per level after quadtree + orientation (level coords)

    void compute_pyramid(const Img8& gray);
    void detect_candidates(int level, std::vector<Candidate>& out) const;
    // returns selected candidates (in final list order)
    std::vector<Candidate> distribute_quadtree(const std::vector<Candidate>& cand, int minX, int maxX, int minY,
                                               int maxY, int N) const;
    float ic_angle(const Level& L, int x, int y) const;
    void describe(const Level& blurred, const KeyPoint& kp, uint8_t* out32) const;

    OrbParams prm;
    std::vector<float> scale, inv_scale, sigma2, inv_sigma2;
    std::vector<int> features_per_level;
    std::vector<int> umax;
};

}  // namespace oracle
