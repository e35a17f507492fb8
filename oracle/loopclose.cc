// TEST INFRASTRUCTURE ONLY — see loopclose.h.
#include "loopclose.h"

#include <algorithm>
#include <cmath>
#include <utility>
#include <vector>

#include "bow_transform.h"
#include "match.h"

namespace oracle {
namespace {

// the scan of the inverted file (src/KeyFrameDatabase.cc:85-105 / :206-223): words of the query in ascending order, each word's list in insertion order
// (= database order).  listed[] = the order in which key frames enter lKFsSharingWords.
void shared_words(const int32_t* q_word, int n_q, const BowDatabase& db, const uint8_t* connected, std::vector<int>& words, std::vector<int>& listed) {
    words.assign(db.n_kf, 0);
    std::vector<int> first(db.n_kf, -1);
    for (int k = 0; k < db.n_kf; ++k) {
        const int32_t* w = db.word + db.off[k];
        const int n = db.off[k + 1] - db.off[k];
        int a = 0, b = 0;
        while (a < n_q && b < n) {
            if (q_word[a] == w[b]) { if (first[k] < 0) first[k] = a; ++words[k]; ++a; ++b; }
            else if (q_word[a] < w[b]) ++a;
            else ++b;
        }
    }
    std::vector<std::pair<int, int>> order;
    for (int k = 0; k < db.n_kf; ++k) {
        if (first[k] < 0) continue;
        if (connected && connected[k]) { words[k] = 1; continue; }      // mnLoopQuery is never set for a connected key frame: its counter restarts at every word
        order.emplace_back(first[k], k);
    }
    std::sort(order.begin(), order.end());
    listed.clear();
    for (auto& p : order) listed.push_back(p.second);
}

}  // namespace

int detect_loop_candidates(const int32_t* q_word, const double* q_val, int n_q, const BowDatabase& db, const uint8_t* connected, float min_score,
                           int32_t* cand, int32_t* common_words, float* score) {
    std::vector<int> words, listed;
    shared_words(q_word, n_q, db, connected, words, listed);
    if (common_words) for (int k = 0; k < db.n_kf; ++k) common_words[k] = words[k];
    if (listed.empty()) return 0;
    int maxCommonWords = 0;
    for (int k : listed) maxCommonWords = std::max(maxCommonWords, words[k]);
    const int minCommonWords = (int)(maxCommonWords * 0.8f);
    std::vector<char> is_listed(db.n_kf, 0);
    for (int k : listed) is_listed[k] = 1;
    std::vector<float> loop_score(db.n_kf, 0.f);
    std::vector<std::pair<float, int>> lScoreAndMatch;
    for (int k : listed) {
        if (words[k] > minCommonWords) {
            const float si = (float)bow_score_l1(q_word, q_val, n_q, db.word + db.off[k], db.val + db.off[k], db.off[k + 1] - db.off[k]);
            loop_score[k] = si;
            if (score) score[k] = si;
            if (si >= min_score) lScoreAndMatch.emplace_back(si, k);
        }
    }
    if (lScoreAndMatch.empty()) return 0;
    std::vector<std::pair<float, int>> lAcc;
    float bestAccScore = min_score;
    for (auto& sm : lScoreAndMatch) {
        float bestScore = sm.first, accScore = sm.first;
        int best = sm.second;
        for (int j = 0; j < db.covis_stride; ++j) {
            const int k2 = db.covis ? db.covis[(size_t)sm.second * db.covis_stride + j] : -1;
            if (k2 < 0) break;
            if (is_listed[k2] && words[k2] > minCommonWords) {
                accScore += loop_score[k2];
                if (loop_score[k2] > bestScore) { best = k2; bestScore = loop_score[k2]; }
            }
        }
        lAcc.emplace_back(accScore, best);
        if (accScore > bestAccScore) bestAccScore = accScore;
    }
    const float minScoreToRetain = 0.75f * bestAccScore;
    std::vector<char> added(db.n_kf, 0);
    int n = 0;
    for (auto& a : lAcc)
        if (a.first > minScoreToRetain && !added[a.second]) { cand[n++] = a.second; added[a.second] = 1; }
    return n;
}

int detect_relocalization_candidates(const int32_t* q_word, const double* q_val, int n_q, const BowDatabase& db, float* reloc_score, int32_t* cand,
                                     int32_t* common_words) {
    std::vector<int> words, listed;
    shared_words(q_word, n_q, db, nullptr, words, listed);
    if (common_words) for (int k = 0; k < db.n_kf; ++k) common_words[k] = words[k];
    if (listed.empty()) return 0;
    int maxCommonWords = 0;
    for (int k : listed) maxCommonWords = std::max(maxCommonWords, words[k]);
    const int minCommonWords = (int)(maxCommonWords * 0.8f);
    std::vector<char> is_listed(db.n_kf, 0);
    for (int k : listed) is_listed[k] = 1;
    std::vector<std::pair<float, int>> lScoreAndMatch;
    for (int k : listed) {
        if (words[k] > minCommonWords) {
            const float si = (float)bow_score_l1(q_word, q_val, n_q, db.word + db.off[k], db.val + db.off[k], db.off[k + 1] - db.off[k]);
            reloc_score[k] = si;
            lScoreAndMatch.emplace_back(si, k);
        }
    }
    if (lScoreAndMatch.empty()) return 0;
    std::vector<std::pair<float, int>> lAcc;
    float bestAccScore = 0;
    for (auto& sm : lScoreAndMatch) {
        float bestScore = sm.first, accScore = bestScore;
        int best = sm.second;
        for (int j = 0; j < db.covis_stride; ++j) {
            const int k2 = db.covis ? db.covis[(size_t)sm.second * db.covis_stride + j] : -1;
            if (k2 < 0) break;
            if (!is_listed[k2]) continue;                   // mnRelocQuery != F->mnId
            accScore += reloc_score[k2];
            if (reloc_score[k2] > bestScore) { best = k2; bestScore = reloc_score[k2]; }
        }
        lAcc.emplace_back(accScore, best);
        if (accScore > bestAccScore) bestAccScore = accScore;
    }
    const float minScoreToRetain = 0.75f * bestAccScore;
    std::vector<char> added(db.n_kf, 0);
    int n = 0;
    for (auto& a : lAcc)
        if (a.first > minScoreToRetain && !added[a.second]) { cand[n++] = a.second; added[a.second] = 1; }
    return n;
}

int search_by_bow_kf(const BowSide& K1, const uint8_t* has_mp1, const BowSide& K2, const uint8_t* has_mp2, float nnratio, bool check_orientation, int32_t* match12) {
    const int TH_LOW = 50, HISTO_LENGTH = 30;
    for (int i = 0; i < K1.n; ++i) match12[i] = -1;
    std::vector<char> matched2(K2.n, 0);
    std::vector<int> rotHist[30];
    const float factor = 1.0f / HISTO_LENGTH;
    int nmatches = 0, a = 0, b = 0;
    while (a < K1.n_nodes && b < K2.n_nodes) {
        if (K1.node_id[a] == K2.node_id[b]) {
            for (int q = K1.node_off[a]; q < K1.node_off[a + 1]; ++q) {
                const int idx1 = K1.node_feat[q];
                if (!has_mp1[idx1]) continue;
                int bestDist1 = 256, bestIdx2 = -1, bestDist2 = 256;
                for (int p = K2.node_off[b]; p < K2.node_off[b + 1]; ++p) {
                    const int idx2 = K2.node_feat[p];
                    if (matched2[idx2] || !has_mp2[idx2]) continue;
                    const int dist = descriptor_distance(K1.desc + 32 * idx1, K2.desc + 32 * idx2);
                    if (dist < bestDist1) { bestDist2 = bestDist1; bestDist1 = dist; bestIdx2 = idx2; }
                    else if (dist < bestDist2) bestDist2 = dist;
                }
                if (bestDist1 < TH_LOW && static_cast<float>(bestDist1) < nnratio * static_cast<float>(bestDist2)) {
                    match12[idx1] = bestIdx2;
                    matched2[bestIdx2] = 1;
                    if (check_orientation) {
                        float rot = K1.angle[idx1] - K2.angle[bestIdx2];
                        if (rot < 0.0) rot += 360.0f;
                        int bin = (int)std::round(rot * factor);
                        if (bin == HISTO_LENGTH) bin = 0;
                        rotHist[bin].push_back(idx1);
                    }
                    nmatches++;
                }
            }
            ++a; ++b;
        } else if (K1.node_id[a] < K2.node_id[b]) {
            a = (int)(std::lower_bound(K1.node_id, K1.node_id + K1.n_nodes, K2.node_id[b]) - K1.node_id);
        } else {
            b = (int)(std::lower_bound(K2.node_id, K2.node_id + K2.n_nodes, K1.node_id[a]) - K2.node_id);
        }
    }
    if (check_orientation) {
        int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            const int s = (int)rotHist[i].size();
            if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
            else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
            else if (s > max3) { max3 = s; ind3 = i; }
        }
        if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; } else if (max3 < 0.1f * (float)max1) ind3 = -1;
        for (int i = 0; i < HISTO_LENGTH; ++i) {
            if (i == ind1 || i == ind2 || i == ind3) continue;
            for (int j : rotHist[i]) { match12[j] = -1; nmatches--; }
        }
    }
    return nmatches;
}

}  // namespace oracle
