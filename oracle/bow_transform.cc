// TEST INFRASTRUCTURE ONLY — see bow_transform.h.
#include "bow_transform.h"

#include <algorithm>
#include <cmath>
#include <map>

#include "match.h"

namespace oracle {

void bow_transform(const BowVocabulary& V, const uint8_t* features, int n, int levelsup, BowResult& out) {
    std::map<int32_t, double> v;
    std::map<int32_t, std::vector<int32_t>> fv;
    out.feat_word.assign(n, -1); out.feat_node.assign(n, 0);
    const int nid_level = V.L - levelsup;
    for (int i = 0; i < n; ++i) {
        const uint8_t* f = features + 32 * i;
        int nid = 0, final_id = 0, current_level = 0;
        do {
            ++current_level;
            const int c0 = V.child_off[final_id], c1 = V.child_off[final_id + 1];
            final_id = V.child_id[c0];
            double best_d = descriptor_distance(f, V.desc + 32 * final_id);
            for (int c = c0 + 1; c < c1; ++c) {
                const int id = V.child_id[c];
                const double d = descriptor_distance(f, V.desc + 32 * id);
                if (d < best_d) { best_d = d; final_id = id; }
            }
            if (current_level == nid_level) nid = final_id;
        } while (V.child_off[final_id + 1] > V.child_off[final_id]);          // !isLeaf()
        const int wid = V.word_id[final_id];
        const double w = V.weight[final_id];
        out.feat_word[i] = wid; out.feat_node[i] = nid;
        if (w > 0) {                                                              // not stopped
            auto it = v.lower_bound(wid);
            if (it != v.end() && it->first == wid) it->second += w; else v.insert(it, {wid, w});
            fv[nid].push_back(i);
        }
    }
    // L1 scoring object: mustNormalize -> BowVector::normalize(L1)
    double norm = 0.0;
    for (auto& kv : v) norm += std::fabs(kv.second);
    if (norm > 0.0) for (auto& kv : v) kv.second /= norm;
    out.word_id.clear(); out.word_val.clear();
    for (auto& kv : v) { out.word_id.push_back(kv.first); out.word_val.push_back(kv.second); }
    out.node_id.clear(); out.node_off.assign(1, 0); out.node_feat.clear();
    for (auto& kv : fv) {
        out.node_id.push_back(kv.first);
        out.node_feat.insert(out.node_feat.end(), kv.second.begin(), kv.second.end());
        out.node_off.push_back((int32_t)out.node_feat.size());
    }
}

double bow_score_l1(const int32_t* id1, const double* v1, int n1, const int32_t* id2, const double* v2, int n2) {
    double score = 0;
    int a = 0, b = 0;
    while (a < n1 && b < n2) {
        if (id1[a] == id2[b]) { score += std::fabs(v1[a] - v2[b]) - std::fabs(v1[a]) - std::fabs(v2[b]); ++a; ++b; }
        else if (id1[a] < id2[b]) a = (int)(std::lower_bound(id1, id1 + n1, id2[b]) - id1);
        else b = (int)(std::lower_bound(id2, id2 + n2, id1[a]) - id2);
    }
    return -score / 2.0;
}

}  // namespace oracle
