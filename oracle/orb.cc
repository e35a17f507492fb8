// TEST INFRASTRUCTURE ONLY — CPU oracle restating the reference ORB extractor (see orb.h for citations).
#include "orb.h"

#include <algorithm>
#include <cmath>
#include <list>

namespace oracle {

static const int8_t kPattern[1024] = {
#include "orb_pattern.inc"
};

static const int kPatch = 31, kHalfPatch = 15, kEdge = 19;

// ---- ctor: scale tables, per-level quotas, disc half-widths (reference :410-470) ----
OrbExtractor::OrbExtractor(const OrbParams& p) : prm(p) {
    const int L = p.nlevels;
    scale.assign(L, 1.f); sigma2.assign(L, 1.f); inv_scale.resize(L); inv_sigma2.resize(L);
    for (int i = 1; i < L; ++i) {
        scale[i] = scale[i - 1] * p.scale_factor;
        sigma2[i] = scale[i] * scale[i];
    }
    for (int i = 0; i < L; ++i) { inv_scale[i] = 1.0f / scale[i]; inv_sigma2[i] = 1.0f / sigma2[i]; }

    features_per_level.resize(L);
    const float factor = 1.0f / p.scale_factor;
    float want = p.nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)L));
    int sum = 0;
    for (int l = 0; l < L - 1; ++l) {
        features_per_level[l] = cv_round(want);
        sum += features_per_level[l];
        want *= factor;
    }
    features_per_level[L - 1] = std::max(p.nfeatures - sum, 0);

    umax.assign(kHalfPatch + 1, 0);
    const int vmax = (int)std::floor(kHalfPatch * std::sqrt(2.f) / 2 + 1);
    const int vmin = (int)std::ceil(kHalfPatch * std::sqrt(2.f) / 2);
    const double hp2 = kHalfPatch * kHalfPatch;
    for (int v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (int v = kHalfPatch, v0 = 0; v >= vmin; --v) {
        while (umax[v0] == umax[v0 + 1]) ++v0;
        umax[v] = v0;
        ++v0;
    }
}

// ---- pyramid (reference :1107-1132); borders are not materialised: no consumer on this path reads them ----
void OrbExtractor::compute_pyramid(const Img8& gray) {
    pyramid.assign(prm.nlevels, Level());
    for (int l = 0; l < prm.nlevels; ++l) {
        Level& L = pyramid[l];
        L.w = cv_round((float)gray.w * inv_scale[l]);
        L.h = cv_round((float)gray.h * inv_scale[l]);
        L.px.resize((size_t)L.w * L.h);
        if (l == 0) {
            for (int y = 0; y < L.h; ++y) std::copy(gray.p + (size_t)y * gray.stride, gray.p + (size_t)y * gray.stride + L.w, &L.px[(size_t)y * L.w]);
        } else {
            const Level& P = pyramid[l - 1];
            resize_linear_u8(Img8{P.px.data(), P.w, P.h, P.w}, L.px.data(), L.w, L.h);
        }
    }
}

// ---- per-cell FAST with threshold fallback (reference :771-829) ----
void OrbExtractor::detect_candidates(int level, std::vector<Candidate>& out) const {
    out.clear();
    const Level& L = pyramid[level];
    const int minBX = kEdge - 3, minBY = minBX, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
    const float width = (float)(maxBX - minBX), height = (float)(maxBY - minBY);
    const int nCols = (int)(width / 30.f), nRows = (int)(height / 30.f);
    if (nCols <= 0 || nRows <= 0) return;
    const int wCell = (int)std::ceil(width / nCols), hCell = (int)std::ceil(height / nRows);
    std::vector<FastKp> cell;
    for (int i = 0; i < nRows; ++i) {
        const float iniY = (float)(minBY + i * hCell);
        float maxY = iniY + hCell + 6;
        if (iniY >= maxBY - 3) continue;
        if (maxY > maxBY) maxY = (float)maxBY;
        for (int j = 0; j < nCols; ++j) {
            const float iniX = (float)(minBX + j * wCell);
            float maxX = iniX + wCell + 6;
            if (iniX >= maxBX - 6) continue;
            if (maxX > maxBX) maxX = (float)maxBX;
            const int x0 = (int)iniX, y0 = (int)iniY, x1 = (int)maxX, y1 = (int)maxY;
            Img8 sub{L.px.data() + (size_t)y0 * L.w + x0, x1 - x0, y1 - y0, L.w};
            fast_detect(sub, prm.ini_th_fast, cell);
            if (cell.empty()) fast_detect(sub, prm.min_th_fast, cell);
            for (const FastKp& k : cell) out.push_back({k.x + j * wCell, k.y + i * hCell, k.score});
        }
    }
}

// ---- quadtree distribution (reference DistributeOctTree :539-763, DivideNode :481-537) ----
namespace {
struct QNode {
    int ulx, uly, urx, bry;          // UL=(ulx,uly) UR=(urx,uly) BL=(ulx,bry) BR=(urx,bry)
    std::vector<int> keys;           // indices into the candidate array, original relative order
    bool no_more = false;
    std::list<int>::iterator it;     // position in the ordered node list
};
}  // namespace

std::vector<Candidate> OrbExtractor::distribute_quadtree(const std::vector<Candidate>& cand, int minX, int maxX,
                                                         int minY, int maxY, int N) const {
    std::vector<Candidate> result;
    const int nIni = (int)std::round((float)(maxX - minX) / (maxY - minY));
    if (nIni <= 0) return result;
    const float hX = (float)(maxX - minX) / nIni;

    std::vector<QNode> nodes;          // creation order == canonical "pointer" order
    nodes.reserve(cand.size() * 4 + 16);
    std::list<int> order;              // node ids in list order
    for (int i = 0; i < nIni; ++i) {
        QNode n;
        n.ulx = (int)(hX * (float)i); n.urx = (int)(hX * (float)(i + 1)); n.uly = 0; n.bry = maxY - minY;
        nodes.push_back(n);
        order.push_back(i);
    }
    for (int k = 0; k < (int)cand.size(); ++k) nodes[(int)((float)cand[k].x / hX)].keys.push_back(k);
    for (auto it = order.begin(); it != order.end();) {
        QNode& n = nodes[*it];
        if (n.keys.size() == 1) { n.no_more = true; ++it; }
        else if (n.keys.empty()) it = order.erase(it);
        else ++it;
    }

    // split node `id` into up to four children pushed to the list front in the order 1,2,3,4;
    // children holding more than one key are appended to `expandable` as (count, id).
    auto split = [&](int id, std::vector<std::pair<int, int>>& expandable, int& nToExpand) {
        const int ulx = nodes[id].ulx, uly = nodes[id].uly, urx = nodes[id].urx, bry = nodes[id].bry;
        const int halfX = (int)std::ceil((float)(urx - ulx) / 2), halfY = (int)std::ceil((float)(bry - uly) / 2);
        const int midx = ulx + halfX, midy = uly + halfY;
        QNode ch[4];
        ch[0].ulx = ulx;  ch[0].uly = uly;  ch[0].urx = midx; ch[0].bry = midy;
        ch[1].ulx = midx; ch[1].uly = uly;  ch[1].urx = urx;  ch[1].bry = midy;
        ch[2].ulx = ulx;  ch[2].uly = midy; ch[2].urx = midx; ch[2].bry = bry;
        ch[3].ulx = midx; ch[3].uly = midy; ch[3].urx = urx;  ch[3].bry = bry;
        for (int k : nodes[id].keys) {
            const float x = (float)cand[k].x, y = (float)cand[k].y;
            int c = (x < (float)midx) ? ((y < (float)midy) ? 0 : 2) : ((y < (float)midy) ? 1 : 3);
            ch[c].keys.push_back(k);
        }
        for (int c = 0; c < 4; ++c) {
            if (ch[c].keys.empty()) continue;
            ch[c].no_more = ch[c].keys.size() == 1;
            const int cid = (int)nodes.size();
            const int cnt = (int)ch[c].keys.size();
            nodes.push_back(std::move(ch[c]));
            order.push_front(cid);
            nodes[cid].it = order.begin();
            if (cnt > 1) { ++nToExpand; expandable.emplace_back(cnt, cid); }
        }
    };

    bool finish = false;
    std::vector<std::pair<int, int>> expandable;
    while (!finish) {
        const int prevSize = (int)order.size();
        int nToExpand = 0;
        expandable.clear();
        for (auto it = order.begin(); it != order.end();) {
            const int id = *it;
            if (nodes[id].no_more) { ++it; continue; }
            split(id, expandable, nToExpand);
            it = order.erase(it);
        }
        if ((int)order.size() >= N || (int)order.size() == prevSize) {
            finish = true;
        } else if ((int)order.size() + nToExpand * 3 > N) {
            while (!finish) {
                const int prev2 = (int)order.size();
                std::vector<std::pair<int, int>> prev = expandable;
                expandable.clear();
                std::sort(prev.begin(), prev.end());   // (count, creation id) ascending
                for (int j = (int)prev.size() - 1; j >= 0; --j) {
                    int dummy = 0;
                    const int id = prev[j].second;
                    split(id, expandable, dummy);
                    order.erase(nodes[id].it);
                    if ((int)order.size() >= N) break;
                }
                if ((int)order.size() >= N || (int)order.size() == prev2) finish = true;
            }
        }
    }

    result.reserve(order.size());
    for (int id : order) {
        const std::vector<int>& ks = nodes[id].keys;
        int best = ks[0];
        for (size_t k = 1; k < ks.size(); ++k)
            if (cand[ks[k]].score > cand[best].score) best = ks[k];
        result.push_back(cand[best]);
    }
    return result;
}

// ---- intensity-centroid orientation (reference IC_Angle :77-104) ----
float OrbExtractor::ic_angle(const Level& L, int x, int y) const {
    int m01 = 0, m10 = 0;
    const uint8_t* c = &L.px[(size_t)y * L.w + x];
    for (int u = -kHalfPatch; u <= kHalfPatch; ++u) m10 += u * c[u];
    for (int v = 1; v <= kHalfPatch; ++v) {
        int vsum = 0;
        const int d = umax[v];
        for (int u = -d; u <= d; ++u) {
            const int p = c[u + v * L.w], m = c[u - v * L.w];
            vsum += p - m;
            m10 += u * (p + m);
        }
        m01 += v * vsum;
    }
    return fast_atan2_deg((float)m01, (float)m10);
}

// ---- steered BRIEF (reference computeOrbDescriptor :108-147) ----
void OrbExtractor::describe(const Level& B, const KeyPoint& kp, uint8_t* out) const {
    const float angle = kp.angle * (float)(M_PI / 180.f);
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const uint8_t* c = &B.px[(size_t)cv_round(kp.y) * B.w + cv_round(kp.x)];
    const int8_t* pat = kPattern;
    auto tap = [&](int idx) -> int {
        const float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
        const int yy = cv_round(px * b + py * a), xx = cv_round(px * a - py * b);
        return c[yy * B.w + xx];
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
        int val = 0;
        for (int k = 0; k < 8; ++k) val |= (tap(2 * k) < tap(2 * k + 1)) << k;
        out[i] = (uint8_t)val;
    }
}

// ---- operator() (reference :1043-1105) ----
void OrbExtractor::extract(const Img8& gray, std::vector<KeyPoint>& kps, std::vector<uint8_t>& desc) {
    kps.clear(); desc.clear();
    if (!gray.p || gray.w <= 0 || gray.h <= 0) return;
    compute_pyramid(gray);
    candidates.assign(prm.nlevels, {});
    level_kps.assign(prm.nlevels, {});
    for (int l = 0; l < prm.nlevels; ++l) {
        const Level& L = pyramid[l];
        const int minB = kEdge - 3, maxBX = L.w - kEdge + 3, maxBY = L.h - kEdge + 3;
        detect_candidates(l, candidates[l]);
        std::vector<Candidate> sel = distribute_quadtree(candidates[l], minB, maxBX, minB, maxBY, features_per_level[l]);
        const int patch = (int)(kPatch * scale[l]);
        for (const Candidate& c : sel) {
            KeyPoint k;
            k.x = (float)c.x + minB; k.y = (float)c.y + minB;
            k.size = (float)patch; k.response = (float)c.score; k.octave = l; k.class_id = -1;
            k.angle = ic_angle(L, cv_round(k.x), cv_round(k.y));
            level_kps[l].push_back(k);
        }
    }
    for (int l = 0; l < prm.nlevels; ++l) {
        if (level_kps[l].empty()) continue;
        const Level& L = pyramid[l];
        Level B; B.w = L.w; B.h = L.h; B.px.resize(L.px.size());
        gaussian_blur_7x7_s2_u8(Img8{L.px.data(), L.w, L.h, L.w}, B.px.data());
        for (KeyPoint k : level_kps[l]) {
            desc.resize(desc.size() + 32);
            describe(B, k, &desc[desc.size() - 32]);
            if (l != 0) { k.x *= scale[l]; k.y *= scale[l]; }
            kps.push_back(k);
        }
    }
}

}  // namespace oracle
