// TEST INFRASTRUCTURE ONLY.  C entry points onto the REFERENCE's own DBoW2 (Thirdparty/DBoW2/DBoW2: TemplatedVocabulary.h, FORB.cpp,
// BowVector.cpp, FeatureVector.cpp, ScoringObject.cpp; DUtils/Random.cpp), compiled unmodified from /root/reference with the container
// stand-ins of oracle/ref/shims/.  The vocabulary is read by the reference's own text loader (loadFromTextFile, the path ORB-SLAM's
// ORBvoc.txt takes).  Built into oracle/_ref/libbow_ref.so by `make -C oracle ref`; pins oracle/bow_transform.cc (tests/test_oracle_bow_ref.py).
#include <cstdint>
#include <cstring>
#include <vector>

#include "Thirdparty/DBoW2/DBoW2/FORB.h"
#include "Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;      // include/ORBVocabulary.h:30-31

extern "C" {
void* ref_voc_load(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void ref_voc_free(void* v) { delete (ORBVocabulary*)v; }
int ref_voc_size(void* v) { return (int)((ORBVocabulary*)v)->size(); }
// Frame::ComputeBoW (src/Frame.cc:538-544): transform(Converter::toDescriptorVector(mDescriptors), mBowVec, mFeatVec, levelsup = 4)
// word_id / word_val: the BowVector in map order; node_id / node_off / node_feat: the FeatureVector as CSR.  cnt = {n_words, n_nodes}.
void ref_bow_transform(void* voc, const uint8_t* desc, int n, int levelsup, int32_t* word_id, double* word_val, int32_t* node_id, int32_t* node_off,
                       int32_t* node_feat, int32_t* cnt) {
    std::vector<cv::Mat> features;
    features.reserve(n);
    for (int i = 0; i < n; ++i) features.push_back(cv::Mat(1, 32, CV_8U, (void*)(desc + 32 * (size_t)i)));      // Converter::toDescriptorVector: one row each
    DBoW2::BowVector bv;
    DBoW2::FeatureVector fv;
    ((ORBVocabulary*)voc)->transform(features, bv, fv, levelsup);
    int w = 0;
    for (const auto& e : bv) { word_id[w] = (int32_t)e.first; word_val[w] = e.second; ++w; }
    int k = 0, off = 0;
    for (const auto& e : fv) {
        node_id[k] = (int32_t)e.first; node_off[k] = off;
        for (unsigned f : e.second) node_feat[off++] = (int32_t)f;
        ++k;
    }
    node_off[k] = off;
    cnt[0] = w; cnt[1] = k;
}
// ORBVocabulary::score (L1): BowVector a vs b given as (id, value) arrays
double ref_bow_score(void* voc, const int32_t* ida, const double* va, int na, const int32_t* idb, const double* vb, int nb) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < na; ++i) a.insert(a.end(), std::make_pair((DBoW2::WordId)ida[i], va[i]));
    for (int i = 0; i < nb; ++i) b.insert(b.end(), std::make_pair((DBoW2::WordId)idb[i], vb[i]));
    return ((ORBVocabulary*)voc)->score(a, b);
}
}
