// TEST INFRASTRUCTURE ONLY — see match.h.
#include "match.h"

#include <algorithm>

namespace oracle {

int descriptor_distance(const uint8_t* a, const uint8_t* b) {      // SWAR popcount over 8 x 32 bit, like the reference
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(a);
    const uint32_t* pb = reinterpret_cast<const uint32_t*>(b);
    int dist = 0;
    for (int i = 0; i < 8; ++i) {
        uint32_t v = pa[i] ^ pb[i];
        v = v - ((v >> 1) & 0x55555555);
        v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
        dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
    }
    return dist;
}

void bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
    for (int i = 0; i < nq; ++i) {
        int best = 256 + 1, bi = -1;
        for (int j = 0; j < nt; ++j) {
            const int d = descriptor_distance(q + 32 * i, t + 32 * j);
            if (d < best) { best = d; bi = j; }
        }
        idx[i] = bi; dist[i] = bi < 0 ? 256 : best;
    }
}

void bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) {
    for (int i = 0; i < nq; ++i) {
        int d0 = 257, d1 = 257, i0 = -1, i1 = -1;
        for (int j = 0; j < nt; ++j) {
            const int d = descriptor_distance(q + 32 * i, t + 32 * j);
            if (d < d0) { d1 = d0; i1 = i0; d0 = d; i0 = j; }
            else if (d < d1) { d1 = d; i1 = j; }
        }
        idx[2 * i] = i0; idx[2 * i + 1] = i1;
        dist[2 * i] = i0 < 0 ? 256 : d0; dist[2 * i + 1] = i1 < 0 ? 256 : d1;
    }
}

std::vector<int> match_orb_points_gate(const int32_t* dist, int nq) {
    double min_dist = 1000;
    for (int i = 0; i < nq; ++i) if (dist[i] < min_dist) min_dist = dist[i];
    std::vector<int> keep;
    for (int i = 0; i < nq; ++i) if (dist[i] < std::max(2 * min_dist, 15.0)) keep.push_back(i);
    return keep;
}

}  // namespace oracle

extern "C" {
int orc_descriptor_distance(const uint8_t* a, const uint8_t* b) { return oracle::descriptor_distance(a, b); }
void orc_bf_match(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) { oracle::bf_match(q, nq, t, nt, idx, dist); }
void orc_bf_knn2(const uint8_t* q, int nq, const uint8_t* t, int nt, int32_t* idx, int32_t* dist) { oracle::bf_knn2(q, nq, t, nt, idx, dist); }
int orc_match_gate(const int32_t* dist, int nq, int32_t* keep) {
    std::vector<int> k = oracle::match_orb_points_gate(dist, nq);
    for (size_t i = 0; i < k.size(); ++i) keep[i] = k[i];
    return (int)k.size();
}
}
