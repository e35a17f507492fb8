// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of the loop-closure / relocalisation candidate search that consumes the key-frame descriptor exchange:
//   KeyFrameDatabase::DetectLoopCandidates(KeyFrame*, float minScore)          src/KeyFrameDatabase.cc:76-197
//   KeyFrameDatabase::DetectRelocalizationCandidates(Frame*)                   src/KeyFrameDatabase.cc:199-305
//   DBoW2 L1 score (bow_transform.h: bow_score_l1)                              Thirdparty/DBoW2/DBoW2/ScoringObject.cpp:23-68
//   ORBmatcher::SearchByBoW(KeyFrame*, KeyFrame*, vector<MapPoint*>&)          src/ORBmatcher.cc:526-659
// The database arrives as plain arrays: key frames in the order KeyFrameDatabase::add saw them (every inverted-file list keeps that order,
// src/KeyFrameDatabase.cc:38-44), their BowVectors as CSR (word ids ascending like the std::map), the covisibility neighbours
// KeyFrame::GetBestCovisibilityKeyFrames(10) as an index table.  Pinned: identical candidate lists / match lists to the reference's own
// src/KeyFrameDatabase.cc + src/ORBmatcher.cc + Thirdparty/DBoW2 compiled here (oracle/ref/match_driver.cc, tests/test_oracle_loopclose_ref.py).
#pragma once
#include <cstdint>

#include "bow.h"

namespace oracle {

struct BowDatabase {
    int n_kf = 0;
    const int32_t* off = nullptr;        // [n_kf + 1] into word / val
    const int32_t* word = nullptr;       // word ids, ascending per key frame
    const double* val = nullptr;
    const int32_t* covis = nullptr;      // [n_kf][covis_stride] best covisibility neighbours (database indices, -1 ends a row)
    int covis_stride = 10;
};

// Per key frame: number of words shared with the query (mnLoopWords / mnRelocWords), position of its first appearance in the scan of the inverted file
// (rank in lKFsSharingWords, -1: not listed), L1 score as the float the reference stores (only where the reference evaluates it, else untouched).
// connected[i] != 0: key frame i is in pKF->GetConnectedKeyFrames() (never listed).  Returns the number of candidates written to cand (reference order).
int detect_loop_candidates(const int32_t* q_word, const double* q_val, int n_q, const BowDatabase& db, const uint8_t* connected, float min_score,
                           int32_t* cand, int32_t* common_words, float* score);
// reloc_score_io[i] = KeyFrame::mRelocScore: the reference adds a neighbour's score whenever that neighbour shares a word with the frame, also when the
// score was not evaluated for this query (too few common words) - the value of an earlier query then.  Stale values in, updated values out.
int detect_relocalization_candidates(const int32_t* q_word, const double* q_val, int n_q, const BowDatabase& db, float* reloc_score_io, int32_t* cand,
                                     int32_t* common_words);
// match12[i1] = feature of key frame 2 whose map point the call stores into vpMatches12[i1] (-1: NULL).  has_mp = map point exists and is not bad.
int search_by_bow_kf(const BowSide& K1, const uint8_t* has_mp1, const BowSide& K2, const uint8_t* has_mp2, float nnratio, bool check_orientation, int32_t* match12);

}  // namespace oracle
