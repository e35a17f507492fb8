// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Descriptor matching on the tracking path:
//   ORBmatcher::DescriptorDistance           src/ORBmatcher.cc:1712-1728 (= LSDmatcher::DescriptorDistance src/LSDmatcher.cpp:316-332)
//   ORBmatcher::MatchORBPoints               src/ORBmatcher.cc:1332-1394  (cv::BFMatcher(NORM_HAMMING).match + distance gate)
//   LSDmatcher::SearchByDescriptor(KF,Frame) src/LSDmatcher.cpp:242-279   (BFMatcher knnMatch k=2 + ratio 1/1.5)
// cv::BFMatcher is not vendored; its brute-force semantics (first minimum wins, k nearest in ascending distance with ties
// in ascending train index) are pinned against cv2 4.13 by tests/test_oracle_match.py.
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {
int descriptor_distance(const uint8_t* a, const uint8_t* b);                      // 256-bit Hamming
// 1-NN for every query row: train index and distance (-1 / 256 when there are no train rows)
void bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);
// 2-NN: idx[2*i], idx[2*i+1]; missing neighbours are -1 with distance 256
void bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist);
// MatchORBPoints' gate: keep match i when dist < max(2*min_dist, 15); returns the kept query indices in order
std::vector<int> match_orb_points_gate(const int32_t* dist, int nq);
}  // namespace oracle
