// Host harness: planarslam_b200/csrc/bowdb_select.h (the list logic behind pslam_detect_loop_candidates / pslam_detect_relocalization_candidates) fed with
// per-key-frame triples computed by a plain scalar loop in the kernel's arithmetic (ordered sum of (|v - w| - |v|) - |w| over the shared words, float(-sum / 2)).
#include <cmath>
#include <cstdint>
#include <vector>

#include "bowdb_select.h"

namespace {
void triples(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word, const double* val, std::vector<int32_t>& common,
             std::vector<int32_t>& first, std::vector<float>& score) {
    common.assign(n_kf, 0); first.assign(n_kf, -1); score.assign(n_kf, 0.f);
    for (int k = 0; k < n_kf; ++k) {
        double sum = 0;
        for (int i = off[k]; i < off[k + 1]; ++i) {
            const int32_t* p = std::lower_bound(q_word, q_word + n_q, word[i]);
            if (p == q_word + n_q || *p != word[i]) continue;
            const int qi = (int)(p - q_word);
            if (first[k] < 0) first[k] = qi;
            ++common[k];
            const double vi = q_val[qi], wi = val[i];
            sum += (std::fabs(vi - wi) - std::fabs(vi)) - std::fabs(wi);
        }
        score[k] = (float)(-sum / 2.0);
    }
}
}  // namespace

extern "C" int host_detect_loop_candidates(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word, const double* val,
                                           const int32_t* covis, int covis_stride, const uint8_t* connected, float min_score, int32_t* cand, int32_t* common_words,
                                           float* score) {
    std::vector<int32_t> common, first; std::vector<float> l1;
    triples(q_word, q_val, n_q, n_kf, off, word, val, common, first, l1);
    if (!pslam::bowdb_covis_ok(n_kf, covis, covis_stride)) return -1;
    return pslam::bowdb_select_loop(n_kf, common.data(), first.data(), l1.data(), covis, covis_stride, connected, min_score, cand, common_words, score);
}
extern "C" int host_detect_relocalization_candidates(const int32_t* q_word, const double* q_val, int n_q, int n_kf, const int32_t* off, const int32_t* word,
                                                     const double* val, const int32_t* covis, int covis_stride, float* reloc_score_io, int32_t* cand,
                                                     int32_t* common_words) {
    std::vector<int32_t> common, first; std::vector<float> l1;
    triples(q_word, q_val, n_q, n_kf, off, word, val, common, first, l1);
    if (!pslam::bowdb_covis_ok(n_kf, covis, covis_stride)) return -1;
    return pslam::bowdb_select_reloc(n_kf, common.data(), first.data(), l1.data(), covis, covis_stride, reloc_score_io, cand, common_words);
}
