// TEST INFRASTRUCTURE ONLY — CPU oracle restating the reference plane extractor (see peac.h for citations).
#include "peac.h"

#include <algorithm>
#include <cmath>
#include <limits>

namespace oracle {

// ---- symmetric 3x3 eigen-decomposition: cyclic Jacobi, eigenvalues ascending, V[:][i] <-> s[i] ----
void eig33sym_jacobi(const double K[3][3], double s[3], double V[3][3]) {
    double a[3][3], d[3], b[3], z[3];
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) { a[i][j] = K[i][j]; V[i][j] = (i == j) ? 1.0 : 0.0; }
        d[i] = b[i] = a[i][i];
        z[i] = 0.0;
    }
    for (int sweep = 0; sweep < 50; ++sweep) {
        const double sm = std::fabs(a[0][1]) + std::fabs(a[0][2]) + std::fabs(a[1][2]);
        if (sm == 0.0) break;
        const double tresh = (sweep < 3) ? 0.2 * sm / 9.0 : 0.0;
        for (int p = 0; p < 2; ++p)
            for (int q = p + 1; q < 3; ++q) {
                const double g = 100.0 * std::fabs(a[p][q]);
                if (sweep > 3 && std::fabs(d[p]) + g == std::fabs(d[p]) && std::fabs(d[q]) + g == std::fabs(d[q])) {
                    a[p][q] = 0.0;
                } else if (std::fabs(a[p][q]) > tresh) {
                    double h = d[q] - d[p], t;
                    if (std::fabs(h) + g == std::fabs(h)) {
                        t = a[p][q] / h;
                    } else {
                        const double theta = 0.5 * h / a[p][q];
                        t = 1.0 / (std::fabs(theta) + std::sqrt(1.0 + theta * theta));
                        if (theta < 0.0) t = -t;
                    }
                    const double c = 1.0 / std::sqrt(1.0 + t * t), sn = t * c, tau = sn / (1.0 + c);
                    h = t * a[p][q];
                    z[p] -= h; z[q] += h; d[p] -= h; d[q] += h;
                    a[p][q] = 0.0;
                    auto rot = [&](int i, int j, int k, int l) {
                        const double gg = a[i][j], hh = a[k][l];
                        a[i][j] = gg - sn * (hh + gg * tau);
                        a[k][l] = hh + sn * (gg - hh * tau);
                    };
                    for (int j = 0; j < p; ++j) rot(j, p, j, q);
                    for (int j = p + 1; j < q; ++j) rot(p, j, j, q);
                    for (int j = q + 1; j < 3; ++j) rot(p, j, q, j);
                    for (int j = 0; j < 3; ++j) {
                        const double gg = V[j][p], hh = V[j][q];
                        V[j][p] = gg - sn * (hh + gg * tau);
                        V[j][q] = hh + sn * (gg - hh * tau);
                    }
                }
            }
        for (int i = 0; i < 3; ++i) { b[i] += z[i]; d[i] = b[i]; z[i] = 0.0; }
    }
    // ascending order, stable selection
    int ord[3] = {0, 1, 2};
    for (int i = 0; i < 2; ++i)
        for (int j = i + 1; j < 3; ++j)
            if (d[ord[j]] < d[ord[i]]) std::swap(ord[i], ord[j]);
    double Vc[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Vc[i][j] = V[i][j];
    for (int k = 0; k < 3; ++k) { s[k] = d[ord[k]]; for (int i = 0; i < 3; ++i) V[i][k] = Vc[i][ord[k]]; }
}

namespace {

struct Thresholds {
    PeacParams p;
    double sim_merge, sim_refine;
    double t_ang_init(double z) const {         // AHCParamSet.hpp:120-127
        double cz = std::max(z, p.z_near);
        cz = std::min(cz, p.z_far);
        const double an = p.angle_near_deg * M_PI / 180.0, af = p.angle_far_deg * M_PI / 180.0;
        const double factor = (af - an) / (p.z_far - p.z_near);
        return std::cos(factor * cz + an - factor * p.z_near);
    }
    double t_mse_init(double z) const { const double v = p.depth_sigma * z * z + p.std_tol_init; return v * v; }
    double t_mse_merge(double z) const { const double v = p.depth_sigma * z * z + p.std_tol_merge; return v * v; }
    double t_dz(double z) const { return p.depth_alpha * std::fabs(z) + p.depth_change_tol; }
};

struct Node {
    PlaneStats st;
    double center[3] = {0, 0, 0}, normal[3] = {0, 0, 0}, mse = 0, curv = 0;
    int N = 0, rid = 0;
    bool nouse = false;
    std::vector<int> nbs;   // ascending node id (canonical replacement for the pointer-ordered std::set)
};

void stats_compute(const PlaneStats& s, double center[3], double normal[3], double& mse, double& curv) {   // AHCPlaneSeg.hpp:125-156
    const double sc = 1.0 / s.N;
    center[0] = s.sx * sc; center[1] = s.sy * sc; center[2] = s.sz * sc;
    double K[3][3] = {{s.sxx - s.sx * s.sx * sc, s.sxy - s.sx * s.sy * sc, s.sxz - s.sx * s.sz * sc},
                      {0, s.syy - s.sy * s.sy * sc, s.syz - s.sy * s.sz * sc},
                      {0, 0, s.szz - s.sz * s.sz * sc}};
    K[1][0] = K[0][1]; K[2][0] = K[0][2]; K[2][1] = K[1][2];
    double sv[3], V[3][3];
    eig33sym_jacobi(K, sv, V);
    if (V[0][0] * center[0] + V[1][0] * center[1] + V[2][0] * center[2] <= 0) {
        normal[0] = V[0][0]; normal[1] = V[1][0]; normal[2] = V[2][0];
    } else {
        normal[0] = -V[0][0]; normal[1] = -V[1][0]; normal[2] = -V[2][0];
    }
    mse = sv[0] * sc;
    curv = sv[0] / (sv[0] + sv[1] + sv[2]);
}

inline double similarity(const Node& a, const Node& b) {
    return std::fabs(a.normal[0] * b.normal[0] + a.normal[1] * b.normal[1] + a.normal[2] * b.normal[2]);
}

void set_insert(std::vector<int>& v, int id) {
    auto it = std::lower_bound(v.begin(), v.end(), id);
    if (it == v.end() || *it != id) v.insert(it, id);
}
void set_erase(std::vector<int>& v, int id) {
    auto it = std::lower_bound(v.begin(), v.end(), id);
    if (it != v.end() && *it == id) v.erase(it);
}

struct Dsu {                 // DisjointSet.hpp:64-92 (union by size, ties keep the first argument's root)
    std::vector<int> parent, size;
    explicit Dsu(int n) : parent(n), size(n, 1) { for (int i = 0; i < n; ++i) parent[i] = i; }
    int find(int x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; }
    int set_size(int x) { return size[find(x)]; }
    void unite(int x, int y) {
        const int xr = find(x), yr = find(y);
        if (xr == yr) return;
        if (size[xr] < size[yr]) { parent[xr] = yr; size[yr] += size[xr]; }
        else { parent[yr] = xr; size[xr] += size[yr]; }
    }
};

// libstdc++ binary heap (std::priority_queue with PlaneSegMinMSECmp, AHCPlaneFitter.hpp:99-107): top = smallest mse
struct MinMseHeap {
    const std::vector<Node>* nodes;
    std::vector<int> h;
    bool comp(int a, int b) const { return (*nodes)[b].mse < (*nodes)[a].mse; }
    void sift_up(int hole, int top, int value) {
        int parent = (hole - 1) / 2;
        while (hole > top && comp(h[parent], value)) { h[hole] = h[parent]; hole = parent; parent = (hole - 1) / 2; }
        h[hole] = value;
    }
    void push(int id) { h.push_back(id); sift_up((int)h.size() - 1, 0, id); }
    int pop() {
        const int top = h[0];
        const int value = h.back();
        h.pop_back();
        const int len = (int)h.size();
        if (len > 0) {
            int hole = 0, child = 0;
            while (child < (len - 1) / 2) {
                child = 2 * (child + 1);
                if (comp(h[child], h[child - 1])) --child;
                h[hole] = h[child]; hole = child;
            }
            if ((len & 1) == 0 && child == (len - 2) / 2) { child = 2 * (child + 1); h[hole] = h[child - 1]; hole = child - 1; }
            sift_up(hole, 0, value);
        }
        return top;
    }
    bool empty() const { return h.empty(); }
};

void disconnect_all(std::vector<Node>& nodes, int id) {
    for (int nb : nodes[id].nbs) set_erase(nodes[nb].nbs, id);
    nodes[id].nbs.clear();
}

// ahCluster (AHCPlaneFitter.hpp:983-1189); `extracted` receives node ids sorted by N descending
void ah_cluster(std::vector<Node>& nodes, MinMseHeap& Q, Dsu& ds, const Thresholds& T, std::vector<int>& extracted) {
    int step = 0;
    while (!Q.empty() && step <= T.p.max_step) {
        const int p = Q.pop();
        if (nodes[p].nouse) continue;
        int best = -1;
        Node cand;
        for (int nb : nodes[p].nbs) {
            if (similarity(nodes[p], nodes[nb]) < T.sim_merge) continue;
            Node m;
            const PlaneStats &a = nodes[p].st, &b = nodes[nb].st;
            m.st.sx = a.sx + b.sx; m.st.sy = a.sy + b.sy; m.st.sz = a.sz + b.sz;
            m.st.sxx = a.sxx + b.sxx; m.st.syy = a.syy + b.syy; m.st.szz = a.szz + b.szz;
            m.st.sxy = a.sxy + b.sxy; m.st.syz = a.syz + b.syz; m.st.sxz = a.sxz + b.sxz; m.st.N = a.N + b.N;
            m.rid = nodes[p].N >= nodes[nb].N ? nodes[p].rid : nodes[nb].rid;
            m.N = m.st.N;
            stats_compute(m.st, m.center, m.normal, m.mse, m.curv);
            if (best < 0 || cand.mse > m.mse || (cand.mse == m.mse && cand.N < m.mse)) { cand = m; best = nb; }   // sic, :1044-1045
        }
        if (best >= 0 && cand.mse < T.t_mse_merge(cand.center[2])) {
            const int id = (int)nodes.size();
            nodes.push_back(cand);
            Q.nodes = &nodes;
            Q.push(id);
            // mergeNbsFrom (AHCPlaneSeg.hpp:379-410)
            ds.unite(nodes[p].rid, nodes[best].rid);
            std::vector<int> u = nodes[p].nbs;
            for (int nb : nodes[best].nbs) set_insert(u, nb);
            set_erase(u, p); set_erase(u, best);
            disconnect_all(nodes, p);
            disconnect_all(nodes, best);
            for (int nb : u) set_insert(nodes[nb].nbs, id);
            nodes[id].nbs = u;
            nodes[p].nouse = nodes[best].nouse = true;
        } else {
            if (nodes[p].N >= T.p.min_support) extracted.push_back(p);
            disconnect_all(nodes, p);
        }
        ++step;
    }
    while (!Q.empty()) {
        const int p = Q.pop();
        if (nodes[p].N >= T.p.min_support) extracted.push_back(p);
        disconnect_all(nodes, p);
    }
    std::stable_sort(extracted.begin(), extracted.end(), [&](int a, int b) { return nodes[b].N < nodes[a].N; });
}

}  // namespace

void peac_run(const uint16_t* depth, int w, int h, float fx, float fy, float cx, float cy, float scale,
              const PeacParams& prm, PeacResult& out) {
    Thresholds T;
    T.p = prm;
    T.sim_merge = std::cos(prm.sim_merge_deg * M_PI / 180.0);
    T.sim_refine = std::cos(prm.sim_refine_deg * M_PI / 180.0);
    out = PeacResult();
    out.w = w; out.h = h;

    // ---- organised cloud (PlaneDetection::readDepthImage, src/PlaneExtractor.cpp:26-57) ----
    std::vector<double> X((size_t)w * h), Y((size_t)w * h), Z((size_t)w * h);
    for (int i = 0; i < h; ++i)
        for (int j = 0; j < w; ++j) {
            const double z = (double)depth[(size_t)i * w + j] * scale;
            Z[(size_t)i * w + j] = z;
            X[(size_t)i * w + j] = ((double)j - cx) * z / fx;
            Y[(size_t)i * w + j] = ((double)i - cy) * z / fy;
        }
    auto get = [&](int i, int j, double& x, double& y, double& z) -> bool {   // include/PlaneExtractor.h:25-33
        const size_t k = (size_t)i * w + j;
        z = Z[k];
        if (z == 0 || std::isnan(z)) return false;
        x = X[k]; y = Y[k];
        return true;
    };

    const int Nh = h / prm.win_h, Nw = w / prm.win_w;
    Dsu ds(Nh * Nw);
    std::vector<Node> nodes;
    nodes.reserve((size_t)Nh * Nw * 2 + 64);
    std::vector<int> G((size_t)Nh * Nw, -1);
    out.blocks.resize((size_t)Nh * Nw);
    MinMseHeap Q;
    Q.nodes = &nodes;

    // ---- initGraph, nodes (AHCPlaneFitter.hpp:798-883; PlaneSeg ctor AHCPlaneSeg.hpp:211-285) ----
    for (int bi = 0; bi < Nh; ++bi)
        for (int bj = 0; bj < Nw; ++bj) {
            Node n;
            n.rid = bi * Nw + bj;
            bool valid = true;
            for (int i = bi * prm.win_h, ic = 0; ic < prm.win_h && i < h && valid; ++i, ++ic)
                for (int j = bj * prm.win_w, jc = 0; jc < prm.win_w && j < w; ++j, ++jc) {
                    double x = 0, y = 0, z = 10000, xn, yn, zn;
                    if (!get(i, j, x, y, z)) { valid = false; break; }
                    if (j + 1 < w && get(i, j + 1, xn, yn, zn) && std::fabs(z - zn) > T.t_dz(z)) { valid = false; break; }
                    if (i + 1 < h && get(i + 1, j, xn, yn, zn) && std::fabs(z - zn) > T.t_dz(z)) { valid = false; break; }
                    PlaneStats& s = n.st;
                    s.sx += x; s.sy += y; s.sz += z;
                    s.sxx += x * x; s.syy += y * y; s.szz += z * z;
                    s.sxy += x * y; s.syz += y * z; s.sxz += x * z;
                    ++s.N;
                }
            if (valid) { n.nouse = false; n.N = n.st.N; }
            else { n.N = 0; n.st = PlaneStats(); n.nouse = true; }
            if (n.N < 4) n.mse = n.curv = std::numeric_limits<double>::quiet_NaN();
            else stats_compute(n.st, n.center, n.normal, n.mse, n.curv);
            PeacBlock& pb = out.blocks[(size_t)bi * Nw + bj];
            pb.stats = n.st; pb.mse = n.mse; pb.curvature = n.curv;
            for (int k = 0; k < 3; ++k) { pb.center[k] = n.center[k]; pb.normal[k] = n.normal[k]; }
            pb.valid = 0;
            if (n.mse < T.t_mse_init(n.center[2]) && !n.nouse) {
                const int id = (int)nodes.size();
                nodes.push_back(n);
                G[(size_t)bi * Nw + bj] = id;
                Q.push(id);
                pb.valid = 1;
            }
        }
    auto connect = [&](int a, int b) { set_insert(nodes[a].nbs, b); set_insert(nodes[b].nbs, a); };
    // ---- initGraph, edges (:894-954): row pass then column pass over (left, centre, right) triples ----
    for (int i = 0; i < Nh; ++i)
        for (int j = 1; j < Nw; j += 2) {
            const int c = i * Nw + j;
            if (G[c - 1] < 0) { --j; continue; }
            if (G[c] < 0) continue;
            if (j < Nw - 1 && G[c + 1] < 0) { ++j; continue; }
            const double th = T.t_ang_init(nodes[G[c]].center[2]);
            if ((j < Nw - 1 && similarity(nodes[G[c - 1]], nodes[G[c + 1]]) >= th) ||
                (j == Nw - 1 && similarity(nodes[G[c]], nodes[G[c - 1]]) >= th)) {
                connect(G[c], G[c - 1]);
                if (j < Nw - 1) connect(G[c], G[c + 1]);
            } else {
                --j;
            }
        }
    for (int j = 0; j < Nw; ++j)
        for (int i = 1; i < Nh; i += 2) {
            const int c = i * Nw + j;
            if (G[c - Nw] < 0) { --i; continue; }
            if (G[c] < 0) continue;
            if (i < Nh - 1 && G[c + Nw] < 0) { ++i; continue; }
            const double th = T.t_ang_init(nodes[G[c]].center[2]);
            if ((i < Nh - 1 && similarity(nodes[G[c - Nw]], nodes[G[c + Nw]]) >= th) ||
                (i == Nh - 1 && similarity(nodes[G[c]], nodes[G[c - Nw]]) >= th)) {
                connect(G[c], G[c - Nw]);
                if (i < Nh - 1) connect(G[c], G[c + Nw]);
            } else {
                --i;
            }
        }

    // ---- ahCluster ----
    std::vector<int> planes;     // node ids of extractedPlanes
    ah_cluster(nodes, Q, ds, T, planes);
    out.n_coarse_planes = (int)planes.size();

    // ---- refineDetails: findBlockMembership (:485-587) ----
    const int P = prm.win_w * prm.win_h;
    std::vector<int32_t>& lab = out.labels;
    lab.assign((size_t)w * h, -1);
    std::vector<int> blk(Nh * Nw, -1);
    std::vector<char> valid_plane(planes.size(), 0);
    std::vector<std::pair<int, int>> rfq;
    auto plid_of_root = [&](int root) { for (size_t k = 0; k < planes.size(); ++k) if (nodes[planes[k]].rid == root) return (int)k; return 0; };  // std::map operator[] default 0
    for (int i = 0, b = 0; i < Nh; ++i)
        for (int j = 0; j < Nw; ++j, ++b) {
            const int setid = ds.find(b);
            if (ds.set_size(setid) * P >= prm.min_support) {
                bool same = true;
                const int nb4[4] = {j > 0 ? b - 1 : -1, j < Nw - 1 ? b + 1 : -1, i > 0 ? b - Nw : -1, i < Nh - 1 ? b + Nw : -1};
                for (int k = 0; k < 4; ++k)
                    if (nb4[k] >= 0 && ds.find(nb4[k]) != setid) { same = false; break; }     // ERODE_ALL_BORDER
                const int plid = plid_of_root(setid);
                if (same) {
                    blk[b] = plid;
                    for (int y = i * prm.win_h; y < (i + 1) * prm.win_h; ++y)
                        for (int x = j * prm.win_w; x < (j + 1) * prm.win_w; ++x) lab[(size_t)y * w + x] = plid;
                    valid_plane[plid] = 1;
                } else blk[b] = -1;
            } else blk[b] = -1;
            // seeds for the region growing
            if (blk[b] < 0) {
                if (i > 0 && blk[b - Nw] >= 0) {
                    const int s0 = (i * prm.win_h - 1) * w + j * prm.win_w;
                    for (int k = 1; k < prm.win_w; ++k) rfq.emplace_back(s0 + k, blk[b - Nw]);
                }
                if (j > 0 && blk[b - 1] >= 0) {
                    const int s0 = (i * prm.win_h) * w + j * prm.win_w - 1;
                    for (int k = 0; k < prm.win_h - 1; ++k) rfq.emplace_back(s0 + k * w, blk[b - 1]);
                }
            } else {
                const int plid = blk[b];
                if (i > 0 && blk[b - Nw] != plid) {
                    const int s0 = (i * prm.win_h) * w + j * prm.win_w;
                    for (int k = 0; k < prm.win_w - 1; ++k) rfq.emplace_back(s0 + k, plid);
                }
                if (j > 0 && blk[b - 1] != plid) {
                    const int s0 = (i * prm.win_h) * w + j * prm.win_w;
                    for (int k = 1; k < prm.win_h; ++k) rfq.emplace_back(s0 + k * w, plid);
                }
            }
        }
    out.coarse_block_plane.assign(blk.begin(), blk.end());

    // ---- floodFill (:428-476) ----
    out.n_seeds = (int)rfq.size();
    {
        std::vector<float> dist((size_t)w * h, std::numeric_limits<float>::max());
        for (size_t k = 0; k < rfq.size(); ++k) {
            const int s = rfq[k].first, plid = rfq[k].second;
            const int sy = s / w, sx = s - sy * w;
            const Node& pl = nodes[planes[plid]];
            int nb[4], nn = 0;
            if (sx > 0) nb[nn++] = s - 1;
            if (sx < w - 1) nb[nn++] = s + 1;
            if (sy > 0) nb[nn++] = s - w;
            if (sy < h - 1) nb[nn++] = s + w;
            for (int t = 0; t < nn; ++t) {
                const int c = nb[t];
                int32_t& trail = lab[c];
                if (trail <= -6) continue;
                if (trail >= 0 && trail == plid) continue;
                const int cy_ = c / w, cx_ = c - cy_ * w;
                const int by = cy_ / prm.win_h, bx = cx_ / prm.win_w;
                const int bid = (by < Nh && bx < Nw) ? by * Nw + bx : -1;
                if (bid >= 0 && blk[bid] >= 0) continue;
                double pt[3];
                float cdist = -1;
                bool ok = false;
                if (get(cy_, cx_, pt[0], pt[1], pt[2])) {
                    const double sd = pl.normal[0] * (pt[0] - pl.center[0]) + pl.normal[1] * (pt[1] - pl.center[1]) +
                                      pl.normal[2] * (pt[2] - pl.center[2]);
                    cdist = (float)std::fabs(sd);
                    ok = (double)cdist * (double)cdist < 9 * pl.mse + 1e-5;
                }
                if (ok) {
                    if (trail >= 0) {
                        const int other = planes[trail];
                        if (similarity(pl, nodes[other]) >= T.sim_refine) connect(other, planes[plid]);
                    }
                    if (cdist < dist[c]) { trail = plid; dist[c] = cdist; rfq.emplace_back(c, plid); }
                    else if (trail < 0) trail -= 1;
                } else if (trail < 0) trail -= 1;
            }
        }
    }

    out.n_queue = (int)rfq.size();
    // ---- last merge (:318-345) and relabel (:352-372) ----
    std::vector<int> final_planes;
    {
        MinMseHeap Q2;
        Q2.nodes = &nodes;
        for (size_t i = 0; i < planes.size(); ++i) if (valid_plane[i]) Q2.push(planes[i]);
        ah_cluster(nodes, Q2, ds, T, final_planes);
    }
    std::vector<int> plidmap(planes.size(), -1);
    for (size_t i = 0; i < planes.size(); ++i) {
        if (!valid_plane[i]) continue;
        const int root = ds.find(nodes[planes[i]].rid);
        for (size_t j = 0; j < final_planes.size(); ++j)
            if (nodes[final_planes[j]].rid == root) { plidmap[i] = (int)j; break; }
    }
    out.membership.assign(final_planes.size(), {});
    for (size_t i = 0; i < (size_t)w * h; ++i) {
        int32_t& pl = lab[i];
        if (pl >= 0 && plidmap[pl] >= 0) { pl = plidmap[pl]; out.membership[pl].push_back((int)i); }
    }
    for (int id : final_planes) {
        PeacPlane p;
        const Node& n = nodes[id];
        for (int k = 0; k < 3; ++k) { p.normal[k] = n.normal[k]; p.center[k] = n.center[k]; }
        p.mse = n.mse; p.curvature = n.curv; p.N = n.N; p.rid = n.rid; p.stats = n.st;
        out.planes.push_back(p);
    }
}

}  // namespace oracle
