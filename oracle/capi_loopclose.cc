// TEST INFRASTRUCTURE ONLY — C entry points of the loop-closure / relocalisation candidate oracle (loopclose.h) for the ctypes tests.
#include <cstdint>

#include "loopclose.h"

namespace {
oracle::BowDatabase make_db(int n_kf, const int32_t* off, const int32_t* word, const double* val, const int32_t* covis, int covis_stride) {
    oracle::BowDatabase db;
    db.n_kf = n_kf; db.off = off; db.word = word; db.val = val; db.covis = covis; db.covis_stride = covis_stride;
    return db;
}
}  // namespace

extern "C" int orc_detect_loop_candidates(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word, const double* val,
                                          const int32_t* covis, int covis_stride, const uint8_t* connected, float min_score, int32_t* cand, int32_t* common_words,
                                          float* score) {
    return oracle::detect_loop_candidates(q_word, q_val, n_q, make_db(n_kf, off, word, val, covis, covis_stride), connected, min_score, cand, common_words, score);
}
extern "C" int orc_detect_relocalization_candidates(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word,
                                                    const double* val, const int32_t* covis, int covis_stride, float* reloc_score_io, int32_t* cand,
                                                    int32_t* common_words) {
    return oracle::detect_relocalization_candidates(q_word, q_val, n_q, make_db(n_kf, off, word, val, covis, covis_stride), reloc_score_io, cand, common_words);
}
extern "C" int orc_search_by_bow_kf(int n1, const uint8_t* desc1, const float* angle1, const uint8_t* has_mp1, int nodes1, const int32_t* node_id1, const int32_t* node_off1,
                                    const int32_t* node_feat1, int n2, const uint8_t* desc2, const float* angle2, const uint8_t* has_mp2, int nodes2,
                                    const int32_t* node_id2, const int32_t* node_off2, const int32_t* node_feat2, float nnratio, int check_ori, int32_t* match12) {
    oracle::BowSide A, B;
    A.n = n1; A.desc = desc1; A.angle = angle1; A.n_nodes = nodes1; A.node_id = node_id1; A.node_off = node_off1; A.node_feat = node_feat1;
    B.n = n2; B.desc = desc2; B.angle = angle2; B.n_nodes = nodes2; B.node_id = node_id2; B.node_off = node_off2; B.node_feat = node_feat2;
    return oracle::search_by_bow_kf(A, has_mp1, B, has_mp2, nnratio, check_ori != 0, match12);
}
