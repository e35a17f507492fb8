// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>::transform(features, BowVector&, FeatureVector&, levelsup)
// Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1125-1193 with the per-feature tree descent :1213-1252, FORB::distance
// FORB.cpp:81-101, BowVector::addWeight / normalize BowVector.cpp:34-46,62-84 and FeatureVector::addFeature, as
// Frame::ComputeBoW / KeyFrame::ComputeBoW call it (src/Frame.cc, src/KeyFrame.cc:66-76: TF_IDF weighting, L1 norm, levelsup 4).
// DBoW2 is vendored in the reference, so the algorithm is fully specified in-tree; pinned: byte-identical to the reference's
// Thirdparty/DBoW2 compiled here (oracle/_ref/libbow_ref.so, tests/test_oracle_bow_ref.py).  The vocabulary is passed as flat arrays (node table in DBoW2's m_nodes order).
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {
struct BowVocabulary {
    int n_nodes = 0, L = 0;                 // depth levels
    const uint8_t* desc = nullptr;          // [n_nodes][32] node descriptors (root's is unused)
    const int32_t* child_off = nullptr;     // [n_nodes + 1] children of node i = child_id[child_off[i] .. child_off[i+1])
    const int32_t* child_id = nullptr;
    const int32_t* word_id = nullptr;       // [n_nodes] for leaves
    const double* weight = nullptr;         // [n_nodes] word weight (idf under TF_IDF)
};
struct BowResult {
    std::vector<int32_t> word_id; std::vector<double> word_val;                     // BowVector in std::map order
    std::vector<int32_t> node_id, node_off, node_feat;                              // FeatureVector as CSR
    std::vector<int32_t> feat_word, feat_node;                                      // per feature (debug)
};
void bow_transform(const BowVocabulary& V, const uint8_t* features, int n, int levelsup, BowResult& out);
// L1Scoring::score (ScoringObject.cpp:23-71)
double bow_score_l1(const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2);
}  // namespace oracle
