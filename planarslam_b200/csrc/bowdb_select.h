// Host-side list logic of KeyFrameDatabase::DetectLoopCandidates / DetectRelocalizationCandidates (src/KeyFrameDatabase.cc:76-305) over the per-key-frame
// triples k_bow_db_scores produces (shared-word count, position of the first shared word in the query, L1 score as float): the common-word gate, the
// covisibility accumulation and the 0.75 x best cut touch a few dozen key frames.  Plain C++ so that tests/test_bowdb_select_host.py can run it without a GPU.
#pragma once
#include <algorithm>
#include <cstdint>
#include <utility>
#include <vector>

namespace pslam {

inline bool bowdb_covis_ok(int n_kf, const int32_t* covis, int covis_stride) {
    if (covis_stride < 0 || (covis_stride && !covis)) return false;
    for (size_t i = 0; i < (size_t)n_kf * covis_stride; ++i)
        if (covis[i] >= n_kf) return false;
    return true;
}

// lKFsSharingWords: key frames in the order the scan of the inverted file meets them - query words ascending, each word's list in database order
inline void bowdb_sharing_order(int n_kf, const int32_t* first, const uint8_t* connected, std::vector<int>& listed) {
    listed.clear();
    for (int k = 0; k < n_kf; ++k)
        if (first[k] >= 0 && !(connected && connected[k])) listed.push_back(k);
    std::stable_sort(listed.begin(), listed.end(), [&](int a, int b) { return first[a] < first[b]; });
}

// the tail both detectors share (:176-196 / :284-304): keep accumulated scores above 0.75 x best, first occurrence of each key frame
inline int bowdb_retain(const std::vector<std::pair<float, int>>& acc, float best_acc, int n_kf, int32_t* candidates) {
    const float min_retain = 0.75f * best_acc;
    std::vector<char> added(n_kf, 0);
    int n = 0;
    for (const auto& a : acc)
        if (a.first > min_retain && !added[a.second]) { candidates[n++] = a.second; added[a.second] = 1; }
    return n;
}

inline int bowdb_select_loop(int n_kf, const int32_t* common, const int32_t* first, const float* l1, const int32_t* covis, int covis_stride, const uint8_t* connected,
                             float min_score, int32_t* candidates, int32_t* common_words, float* score) {
    std::vector<int> listed;
    bowdb_sharing_order(n_kf, first, connected, listed);
    // a connected key frame's counter restarts at every shared word (its mnLoopQuery is never set, :93-103)
    if (common_words) for (int k = 0; k < n_kf; ++k) common_words[k] = (connected && connected[k] && common[k] > 0) ? 1 : common[k];
    if (listed.empty()) return 0;
    int max_common = 0;
    for (int k : listed) max_common = std::max(max_common, common[k]);
    const int min_common = (int)(max_common * 0.8f);
    std::vector<char> scored(n_kf, 0);                       // mnLoopQuery == query id && mnLoopWords > minCommonWords
    std::vector<std::pair<float, int>> above;
    for (int k : listed) {
        if (common[k] <= min_common) continue;
        scored[k] = 1;
        if (score) score[k] = l1[k];
        if (l1[k] >= min_score) above.emplace_back(l1[k], k);
    }
    if (above.empty()) return 0;
    std::vector<std::pair<float, int>> acc;
    float best_acc = min_score;
    for (const auto& sm : above) {
        float best = sm.first, sum = sm.first;
        int best_kf = sm.second;
        for (int j = 0; j < covis_stride; ++j) {
            const int k2 = covis[(size_t)sm.second * covis_stride + j];
            if (k2 < 0) break;
            if (!scored[k2]) continue;
            sum += l1[k2];
            if (l1[k2] > best) { best_kf = k2; best = l1[k2]; }
        }
        acc.emplace_back(sum, best_kf);
        if (sum > best_acc) best_acc = sum;
    }
    return bowdb_retain(acc, best_acc, n_kf, candidates);
}

inline int bowdb_select_reloc(int n_kf, const int32_t* common, const int32_t* first, const float* l1, const int32_t* covis, int covis_stride, float* reloc_score_io,
                              int32_t* candidates, int32_t* common_words) {
    std::vector<int> listed;
    bowdb_sharing_order(n_kf, first, nullptr, listed);
    if (common_words) for (int k = 0; k < n_kf; ++k) common_words[k] = common[k];
    if (listed.empty()) return 0;
    int max_common = 0;
    for (int k : listed) max_common = std::max(max_common, common[k]);
    const int min_common = (int)(max_common * 0.8f);
    std::vector<std::pair<float, int>> evaluated;
    for (int k : listed) {
        if (common[k] <= min_common) continue;
        reloc_score_io[k] = l1[k];
        evaluated.emplace_back(l1[k], k);
    }
    if (evaluated.empty()) return 0;
    std::vector<std::pair<float, int>> acc;
    float best_acc = 0;
    for (const auto& sm : evaluated) {
        float best = sm.first, sum = sm.first;
        int best_kf = sm.second;
        for (int j = 0; j < covis_stride; ++j) {
            const int k2 = covis[(size_t)sm.second * covis_stride + j];
            if (k2 < 0) break;
            if (first[k2] < 0) continue;                      // mnRelocQuery != F->mnId: shares no word with the frame
            sum += reloc_score_io[k2];                        // also a score left by an earlier query (the reference reads mRelocScore unconditionally)
            if (reloc_score_io[k2] > best) { best_kf = k2; best = reloc_score_io[k2]; }
        }
        acc.emplace_back(sum, best_kf);
        if (sum > best_acc) best_acc = sum;
    }
    return bowdb_retain(acc, best_acc, n_kf, candidates);
}

}  // namespace pslam
