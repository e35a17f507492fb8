// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of ORBmatcher::SearchByBoW(KeyFrame*, Frame&, vector<MapPoint*>&)  src/ORBmatcher.cc:160-292 (merge-join of the two
// DBoW2 feature vectors, per common node an ordered greedy best / second-best Hamming match with TH_LOW = 50 and the ratio
// test, rotation histogram + ComputeThreeMaxima :1666-1707).  The DBoW2 transform that produces the feature vectors
// (Thirdparty/DBoW2) is an input here: feature vectors arrive as CSR arrays (node ids ascending like the
// std::map they come from).  Pinned: identical match lists to the reference's own src/ORBmatcher.cc
// compiled here (oracle/ref/match_driver.cc -> oracle/_ref/libmatch_ref.so, tests/test_oracle_match_ref.py).
#pragma once
#include <cstdint>

namespace oracle {
struct BowSide {
    int n = 0;                         // features
    const uint8_t* desc = nullptr;     // [n][32]
    const float* angle = nullptr;      // [n] key-point angle (KF: mvKeysUn, frame: mvKeys)
    int n_nodes = 0;
    const int32_t* node_id = nullptr;  // [n_nodes] ascending
    const int32_t* node_off = nullptr; // [n_nodes + 1]
    const int32_t* node_feat = nullptr;// feature indices, in the order DBoW2 appended them
};
// kf_has_mp[i] = (vpMapPointsKF[i] && !isBad()).  match[j] = KF feature whose map point is stored into vpMapPointMatches[j], -1 none.
int search_by_bow(const BowSide& KF, const uint8_t* kf_has_mp, const BowSide& F, float nnratio, bool check_orientation, int32_t* match);
}  // namespace oracle
