// TEST INFRASTRUCTURE ONLY — CPU oracle restating Optimizer::LocalBundleAdjustment and the g2o block solver it drives
// (see lba.h for the file:line map).
#include "lba.h"
#include "g2o_math.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>
#include <cstdio>
#include <cstdlib>

namespace oracle {
namespace {
using namespace gm;

enum Kind { MONO, STEREO, LINE, PLANE, VER, PAR };

struct Landmark {
    bool is_plane = false;
    V3 p{0, 0, 0};
    Plane pl{{0, 0, 0, 0}};
    int col = -1;          // landmark block index in the current index mapping (-1: not active)
};
struct Pose {
    SE3 T;
    bool fixed = false;
    double fx, fy, cx, cy, bf;
    int col = -1;          // pose block index (-1: fixed or not active)
};
struct Edge {
    Kind kind;
    int dim;
    int kf, lm;
    double obs[3];
    Plane pm;
    double info[3];
    double delta;
    bool robust = true;
    int level = 0;
    double err[3] = {0, 0, 0};
};

// Plane3D::oplus, g2oAddition/Plane3D.h:84-97
void plane_oplus(Plane& p, const double v[3]) {
    const double az = v[0], el = v[1];
    const double s = std::sin(el), c = std::cos(el);
    const V3 n{c * std::cos(az), c * std::sin(az), s};
    const M3 R = plane_rotation(pn(p));
    const double d = (-p.c[3]) + v[2];
    const V3 rn = mul(R, n);
    p.c[0] = rn.x; p.c[1] = rn.y; p.c[2] = rn.z; p.c[3] = -d;
    plane_normalize(p);
}

inline void huber(double e2, double delta, double rho[3]) {      // robust_kernel_impl.cpp:78-91
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
    else { const double s = std::sqrt(e2); rho[0] = 2 * s * delta - dsqr; rho[1] = delta / s; rho[2] = -0.5 * rho[1] / e2; }
}

void compute_error(Edge& e, const Pose& P, const Landmark& L) {
    switch (e.kind) {
        case MONO: {
            const V3 p = se3_map(P.T, L.p);
            e.err[0] = e.obs[0] - (p.x / p.z * P.fx + P.cx);
            e.err[1] = e.obs[1] - (p.y / p.z * P.fy + P.cy);
            break;
        }
        case STEREO: {
            const V3 p = se3_map(P.T, L.p);
            const float invz = 1.0f / (float)p.z;                  // sic, types_six_dof_expmap.cpp:150-157
            const double r0 = p.x * invz * P.fx + P.cx, r1 = p.y * invz * P.fy + P.cy, r2 = r0 - P.bf * invz;
            e.err[0] = e.obs[0] - r0; e.err[1] = e.obs[1] - r1; e.err[2] = e.obs[2] - r2;
            break;
        }
        case LINE: {
            const V3 p = se3_map(P.T, L.p);
            const double u = p.x / p.z * P.fx + P.cx, v = p.y / p.z * P.fy + P.cy;
            e.err[0] = e.obs[0] * u + e.obs[1] * v + e.obs[2]; e.err[1] = 0; e.err[2] = 0;
            break;
        }
        case PLANE: plane_ominus(plane_transform(P.T, L.pl), e.pm, e.err); break;
        case VER: plane_ominus_ver(plane_transform(P.T, L.pl), e.pm, e.err); break;
        case PAR: plane_ominus_par(plane_transform(P.T, L.pl), e.pm, e.err); break;
    }
}
inline double chi2(const Edge& e) { double s = 0; for (int i = 0; i < e.dim; ++i) s += e.err[i] * e.info[i] * e.err[i]; return s; }

// Jacobians: A = d err / d landmark (dim x 3), B = d err / d pose (dim x 6)
void jacobians(Edge& e, const Pose& P, const Landmark& L, double A[3][3], double B[3][6]) {
    for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) A[i][j] = 0; for (int j = 0; j < 6; ++j) B[i][j] = 0; }
    if (e.kind == MONO || e.kind == STEREO || e.kind == LINE) {
        const V3 p = se3_map(P.T, L.p);
        const M3 R = quat_to_matrix(P.T.q);
        const double x = p.x, y = p.y, z = p.z;
        if (e.kind == LINE) {                                       // include/EdgeLine.h:73-114
            const double invz = 1.0 / z, invz_2 = invz * invz, lx = e.obs[0], ly = e.obs[1], fx = P.fx, fy = P.fy;
            B[0][0] = -fy * ly - fx * lx * x * y * invz_2 - fy * ly * y * y * invz_2;
            B[0][1] = fx * lx + fx * lx * x * x * invz_2 + fy * ly * x * y * invz_2;
            B[0][2] = -fx * lx * y * invz + fy * ly * x * invz;
            B[0][3] = fx * lx * invz;
            B[0][4] = fy * ly * invz;
            B[0][5] = -(fx * lx * x + fy * ly * y) * invz_2;
            const double t0 = fx * lx, t1 = fy * ly, t2 = -(fx * lx * x + fy * ly * y) * invz;
            for (int j = 0; j < 3; ++j) A[0][j] = 1. * invz * (t0 * R.m[0][j] + t1 * R.m[1][j] + t2 * R.m[2][j]);
            // rows 1, 2 of tmp are zero
            return;
        }
        const double z_2 = z * z, fx = P.fx, fy = P.fy;
        if (e.kind == MONO) {                                       // types_six_dof_expmap.cpp:103-139
            const double t02 = -x / z * fx, t12 = -y / z * fy;
            for (int j = 0; j < 3; ++j) {
                A[0][j] = -1. / z * (fx * R.m[0][j] + t02 * R.m[2][j]);
                A[1][j] = -1. / z * (fy * R.m[1][j] + t12 * R.m[2][j]);
            }
        } else {                                                    // :188-232
            for (int j = 0; j < 3; ++j) {
                A[0][j] = -fx * R.m[0][j] / z + fx * x * R.m[2][j] / z_2;
                A[1][j] = -fy * R.m[1][j] / z + fy * y * R.m[2][j] / z_2;
                A[2][j] = A[0][j] - P.bf * R.m[2][j] / z_2;
            }
        }
        B[0][0] = x * y / z_2 * fx; B[0][1] = -(1 + (x * x / z_2)) * fx; B[0][2] = y / z * fx;
        B[0][3] = -1. / z * fx; B[0][4] = 0; B[0][5] = x / z_2 * fx;
        B[1][0] = (1 + y * y / z_2) * fy; B[1][1] = -x * y / z_2 * fy; B[1][2] = -x / z * fy;
        B[1][3] = 0; B[1][4] = -1. / z * fy; B[1][5] = y / z_2 * fy;
        if (e.kind == STEREO) {
            B[2][0] = B[0][0] - P.bf * y / z_2; B[2][1] = B[0][1] + P.bf * x / z_2; B[2][2] = B[0][2];
            B[2][3] = B[0][3]; B[2][4] = 0; B[2][5] = B[0][5] - P.bf / z_2;
        }
        return;
    }
    // numeric central differences on both vertices, base_binary_edge.hpp:128-203; the error is restored afterwards.
    // g2o skips the Jacobian of a fixed vertex; it is never read in that case (constructQuadraticForm), so computing it
    // here is harmless.
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    const double keep[3] = {e.err[0], e.err[1], e.err[2]};
    for (int d = 0; d < 3; ++d) {
        double add[3] = {0, 0, 0};
        Landmark Lp = L;
        add[d] = delta;
        plane_oplus(Lp.pl, add);
        compute_error(e, P, Lp);
        const double e1[3] = {e.err[0], e.err[1], e.err[2]};
        Lp = L;
        add[d] = -delta;
        plane_oplus(Lp.pl, add);
        compute_error(e, P, Lp);
        for (int i = 0; i < e.dim; ++i) A[i][d] = scalar * (e1[i] - e.err[i]);
    }
    for (int d = 0; d < 6; ++d) {
        double add[6] = {0, 0, 0, 0, 0, 0};
        Pose Pp = P;
        add[d] = delta;
        Pp.T = se3_mul(se3_exp(add), P.T);
        compute_error(e, Pp, L);
        const double e1[3] = {e.err[0], e.err[1], e.err[2]};
        add[d] = -delta;
        Pp.T = se3_mul(se3_exp(add), P.T);
        compute_error(e, Pp, L);
        for (int i = 0; i < e.dim; ++i) B[i][d] = scalar * (e1[i] - e.err[i]);
    }
    for (int i = 0; i < 3; ++i) e.err[i] = keep[i];
}

// inverse of a 3x3 by cofactors (Eigen compute_inverse_size3)
void inv3(const double D[3][3], double I[3][3]) {
    const double c00 = D[1][1] * D[2][2] - D[1][2] * D[2][1];
    const double c10 = D[1][2] * D[2][0] - D[1][0] * D[2][2];
    const double c20 = D[1][0] * D[2][1] - D[1][1] * D[2][0];
    const double det = c00 * D[0][0] + c10 * D[0][1] + c20 * D[0][2];
    const double id = 1.0 / det;
    I[0][0] = c00 * id; I[1][0] = c10 * id; I[2][0] = c20 * id;
    I[0][1] = (D[0][2] * D[2][1] - D[0][1] * D[2][2]) * id;
    I[1][1] = (D[0][0] * D[2][2] - D[0][2] * D[2][0]) * id;
    I[2][1] = (D[0][1] * D[2][0] - D[0][0] * D[2][1]) * id;
    I[0][2] = (D[0][1] * D[1][2] - D[0][2] * D[1][1]) * id;
    I[1][2] = (D[0][2] * D[1][0] - D[0][0] * D[1][2]) * id;
    I[2][2] = (D[0][0] * D[1][1] - D[0][1] * D[1][0]) * id;
}

// dense unpivoted LDL^T of the symmetric n x n system (upper triangle of H is authoritative); false when a pivot is not positive
bool solve_ldlt(int n, const std::vector<double>& H, const std::vector<double>& b, std::vector<double>& x) {
    std::vector<double> L((size_t)n * n, 0.0), D(n), y(n);
    for (int j = 0; j < n; ++j) {
        double d = H[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        for (int i = j + 1; i < n; ++i) {
            double v = H[(size_t)j * n + i];
            for (int k = 0; k < j; ++k) v -= L[(size_t)i * n + k] * L[(size_t)j * n + k] * D[k];
            L[(size_t)i * n + j] = v / d;
        }
    }
    for (int i = 0; i < n; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[(size_t)i * n + k] * y[k]; y[i] = v; }
    for (int i = 0; i < n; ++i) y[i] /= D[i];
    // (descending k: the order a column-oriented back substitution applies the updates in)
    for (int i = n - 1; i >= 0; --i) { double v = y[i]; for (int k = n - 1; k > i; --k) v -= L[(size_t)k * n + i] * x[k]; x[i] = v; }
    return true;
}

struct Solver {
    std::vector<Pose>& poses;
    std::vector<Landmark>& lms;
    std::vector<Edge>& E;
    std::vector<int> active;          // active edge indices (level 0), creation order
    std::vector<int> pose_of_col, lm_of_col;
    int np = 0, nl = 0;               // active pose / landmark blocks
    // system
    std::vector<double> Hpp;          // (6 np)^2, upper block triangle filled, row-major dense
    std::vector<double> Hll;          // nl x 9
    std::vector<double> bp, bl;       // 6 np, 3 nl
    struct W { int pcol, lcol; double w[6][3]; };   // Hpl block contribution of one edge: B^T Omega_w A
    std::vector<W> Hpl;               // one per active edge whose two vertices are free (duplicates of a pair simply add)
    std::vector<double> xp, xl;
    double lambda = 0, ni = 2;
    int nBad = 0, trials_total = 0;
    double last_chi = 0;

    // initializeOptimization(level 0) + buildIndexMapping
    void initialize() {
        active.clear();
        for (auto& p : poses) p.col = -1;
        for (auto& l : lms) l.col = -1;
        std::vector<char> pa(poses.size(), 0), la(lms.size(), 0);
        for (size_t i = 0; i < E.size(); ++i) if (E[i].level == 0) {
            // allVerticesFixed() is never true: landmarks are never fixed
            active.push_back((int)i); pa[E[i].kf] = 1; la[E[i].lm] = 1;
        }
        pose_of_col.clear(); lm_of_col.clear();
        for (size_t i = 0; i < poses.size(); ++i) if (pa[i] && !poses[i].fixed) { poses[i].col = (int)pose_of_col.size(); pose_of_col.push_back((int)i); }
        for (size_t i = 0; i < lms.size(); ++i) if (la[i]) { lms[i].col = (int)lm_of_col.size(); lm_of_col.push_back((int)i); }
        np = (int)pose_of_col.size(); nl = (int)lm_of_col.size();
        xp.assign((size_t)6 * np, 0.0); xl.assign((size_t)3 * nl, 0.0);
    }
    void active_errors() { for (int i : active) compute_error(E[i], poses[E[i].kf], lms[E[i].lm]); }
    double robust_chi2() const {
        double chi = 0;
        for (int i : active) {
            const Edge& e = E[i];
            const double c = chi2(e);
            if (e.robust) { double rho[3]; huber(c, e.delta, rho); chi += rho[0]; } else chi += c;
        }
        return chi;
    }
    void build_system() {
        const int n = 6 * np;
        Hpp.assign((size_t)n * n, 0.0); Hll.assign((size_t)9 * nl, 0.0);
        bp.assign(n, 0.0); bl.assign((size_t)3 * nl, 0.0);
        Hpl.clear();
        for (int i : active) {
            Edge& e = E[i];
            const Pose& P = poses[e.kf];
            const Landmark& L = lms[e.lm];
            double A[3][3], B[3][6];
            jacobians(e, P, L, A, B);
            double w = 1.0;
            if (e.robust) { double rho[3]; huber(chi2(e), e.delta, rho); w = rho[1]; }
            const int lc = L.col, pc = P.col;
            // landmark (vertex 0, "from"): b += A^T (-w Omega e), Hll += A^T (w Omega) A
            for (int r = 0; r < e.dim; ++r) {
                const double oe = -(e.info[r] * e.err[r]) * w;
                for (int a = 0; a < 3; ++a) {
                    bl[3 * lc + a] += A[r][a] * oe;
                    const double wa = A[r][a] * (w * e.info[r]);
                    for (int c = 0; c < 3; ++c) Hll[9 * lc + 3 * a + c] += wa * A[r][c];
                }
            }
            if (pc < 0) continue;
            W blk; blk.pcol = pc; blk.lcol = lc;
            for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) blk.w[a][c] = 0;
            for (int r = 0; r < e.dim; ++r) {
                const double oe = -(e.info[r] * e.err[r]) * w;
                for (int a = 0; a < 6; ++a) {
                    bp[6 * pc + a] += B[r][a] * oe;
                    const double wa = B[r][a] * (w * e.info[r]);
                    for (int c = 0; c < 6; ++c) Hpp[(size_t)(6 * pc + a) * n + 6 * pc + c] += wa * B[r][c];
                    for (int c = 0; c < 3; ++c) blk.w[a][c] += wa * A[r][c];
                }
            }
            Hpl.push_back(blk);
        }
    }
    // BlockSolver::solve with the Schur complement; lambda already on the diagonals of the copies used here
    bool solve_schur() {
        const int n = 6 * np;
        std::vector<double> Hs(Hpp);                              // _Hschur = _Hpp
        for (int i = 0; i < n; ++i) Hs[(size_t)i * n + i] += lambda;
        std::vector<double> coeff(n, 0.0);
        std::vector<double> Dinv((size_t)9 * nl);
        // group the Hpl blocks by landmark (CCS columns), rows ascending
        std::vector<std::vector<int>> col(nl);
        for (size_t k = 0; k < Hpl.size(); ++k) col[Hpl[k].lcol].push_back((int)k);
        for (int l = 0; l < nl; ++l) {
            std::stable_sort(col[l].begin(), col[l].end(), [&](int a, int b) { return Hpl[a].pcol < Hpl[b].pcol; });
            double D[3][3], I[3][3];
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) D[a][c] = Hll[9 * l + 3 * a + c] + (a == c ? lambda : 0.0);
            inv3(D, I);
            for (int a = 0; a < 3; ++a) for (int c = 0; c < 3; ++c) Dinv[9 * l + 3 * a + c] = I[a][c];
            double db[3];
            for (int a = 0; a < 3; ++a) db[a] = I[a][0] * bl[3 * l] + I[a][1] * bl[3 * l + 1] + I[a][2] * bl[3 * l + 2];
            for (size_t o = 0; o < col[l].size(); ++o) {
                const W& Bi = Hpl[col[l][o]];
                double BD[6][3];
                for (int a = 0; a < 6; ++a) for (int c = 0; c < 3; ++c) BD[a][c] = Bi.w[a][0] * I[0][c] + Bi.w[a][1] * I[1][c] + Bi.w[a][2] * I[2][c];
                for (int a = 0; a < 6; ++a) coeff[6 * Bi.pcol + a] += Bi.w[a][0] * db[0] + Bi.w[a][1] * db[1] + Bi.w[a][2] * db[2];
                for (size_t q = 0; q < col[l].size(); ++q) {
                    const W& Bj = Hpl[col[l][q]];
                    if (Bj.pcol < Bi.pcol) continue;               // upper block triangle only
                    // (two edges between the same pair: g2o would have merged them into one Hpl block; the bilinear
                    // expansion visits both orders on the diagonal block, which is the same sum)
                    for (int a = 0; a < 6; ++a) for (int c = 0; c < 6; ++c)
                        Hs[(size_t)(6 * Bi.pcol + a) * n + 6 * Bj.pcol + c] -= BD[a][0] * Bj.w[c][0] + BD[a][1] * Bj.w[c][1] + BD[a][2] * Bj.w[c][2];
                }
            }
        }
        std::vector<double> bs(n);
        for (int i = 0; i < n; ++i) bs[i] = bp[i] - coeff[i];
        if (n > 0) { if (!solve_ldlt(n, Hs, bs, xp)) return false; }
        // xl = Dinv (bl - Hpl^T xp)
        std::vector<double> cl(bl);
        for (const W& Bk : Hpl)
            for (int c = 0; c < 3; ++c) {
                double s = 0;
                for (int a = 0; a < 6; ++a) s += Bk.w[a][c] * xp[6 * Bk.pcol + a];
                cl[3 * Bk.lcol + c] -= s;
            }
        for (int l = 0; l < nl; ++l)
            for (int a = 0; a < 3; ++a)
                xl[3 * l + a] = Dinv[9 * l + 3 * a] * cl[3 * l] + Dinv[9 * l + 3 * a + 1] * cl[3 * l + 1] + Dinv[9 * l + 3 * a + 2] * cl[3 * l + 2];
        return true;
    }
    void apply_update() {
        for (int c = 0; c < np; ++c) { Pose& P = poses[pose_of_col[c]]; P.T = se3_mul(se3_exp(&xp[6 * c]), P.T); }
        for (int c = 0; c < nl; ++c) {
            Landmark& L = lms[lm_of_col[c]];
            if (L.is_plane) plane_oplus(L.pl, &xl[3 * c]);
            else { L.p.x += xl[3 * c]; L.p.y += xl[3 * c + 1]; L.p.z += xl[3 * c + 2]; }
        }
    }
    int solve(int iteration) {
        active_errors();
        double currentChi = robust_chi2(), tempChi = currentChi;
        const double iniChi = currentChi;
        build_system();
        if (iteration == 0) {
            double mx = 0;
            const int n = 6 * np;
            for (int j = 0; j < n; ++j) mx = std::max(std::fabs(Hpp[(size_t)j * n + j]), mx);
            for (int l = 0; l < nl; ++l) for (int a = 0; a < 3; ++a) mx = std::max(std::fabs(Hll[9 * l + 4 * a]), mx);
            lambda = 1e-5 * mx; ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            std::vector<Pose> bp_ = poses;
            std::vector<Landmark> bl_ = lms;
            const bool ok2 = solve_schur();
            apply_update();
            active_errors();
            tempChi = robust_chi2();
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6 * np; ++j) scale += xp[j] * (lambda * xp[j] + bp[j]);
            for (int j = 0; j < 3 * nl; ++j) scale += xl[j] * (lambda * xl[j] + bl[j]);
            scale += 1e-3;
            rho /= scale;
            if (getenv("ORC_LBA_DEBUG")) fprintf(stderr, "it %d trial %d chi %.6f -> %.6f rho %.4f lambda %.4g ok %d\n", iteration, qmax, currentChi, tempChi, rho, lambda, (int)ok2);
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf; ni = 2; currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2; poses = bp_; lms = bl_;
            }
            ++qmax;
        } while (rho < 0 && qmax < 10);
        trials_total += qmax;
        last_chi = currentChi;
        if (qmax == 10 || rho == 0) return 1;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) return 1;
        return 0;
    }
    int optimize(int iterations) {
        int done = 0;
        trials_total = 0;
        if (np == 0 && nl == 0) return 0;
        bool ok = true;
        for (int i = 0; i < iterations && ok; ++i) { ok = solve(i) == 0; ++done; }
        return done;
    }
};

SE3 se3_from_float16(const float* T) {
    M3 R;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = T[i * 4 + j];
    return se3_from_Rt(R, {T[3], T[7], T[11]});
}

}  // namespace

void local_bundle_adjustment(const LbaProblem& P, LbaResult& out) {
    std::vector<Pose> poses(P.n_kf);
    for (int i = 0; i < P.n_kf; ++i) {
        poses[i].T = se3_from_float16(P.kf_Tcw + 16 * i);
        poses[i].fixed = P.kf_fixed[i] != 0;
        const float* k = P.kf_K + 5 * i;
        poses[i].fx = k[0]; poses[i].fy = k[1]; poses[i].cx = k[2]; poses[i].cy = k[3]; poses[i].bf = k[4];
    }
    const int lm_line0 = P.n_points, lm_plane0 = P.n_points + 2 * P.n_lines;
    std::vector<Landmark> lms(lm_plane0 + P.n_planes);
    for (int i = 0; i < P.n_points; ++i) lms[i].p = {P.pt_Xw[3 * i], P.pt_Xw[3 * i + 1], P.pt_Xw[3 * i + 2]};
    for (int i = 0; i < P.n_lines; ++i) {
        lms[lm_line0 + 2 * i].p = {P.line_Xw[6 * i], P.line_Xw[6 * i + 1], P.line_Xw[6 * i + 2]};
        lms[lm_line0 + 2 * i + 1].p = {P.line_Xw[6 * i + 3], P.line_Xw[6 * i + 4], P.line_Xw[6 * i + 5]};
    }
    for (int i = 0; i < P.n_planes; ++i) { lms[lm_plane0 + i].is_plane = true; lms[lm_plane0 + i].pl = plane_from_float4(P.plane_Xw + 4 * i); }

    const float thHuberMono = std::sqrt(5.991), thHuberStereo = std::sqrt(7.815);
    const double angleInfo = 3282.8 / (P.angle_info * P.angle_info), disInfo = P.dist_info * P.dist_info;
    const float deltaPlane = std::sqrt(P.plane_chi), VPdeltaPlane = std::sqrt(P.vp_chi);

    std::vector<Edge> E;
    E.reserve(P.n_pt_obs + 2 * P.n_line_obs + P.n_plane_obs[0] + P.n_plane_obs[1] + P.n_plane_obs[2]);
    for (int i = 0; i < P.n_pt_obs; ++i) {
        Edge e{};
        const float* o = P.pt_obs_uvr + 3 * i;
        const bool mono = o[2] < 0;
        e.kind = mono ? MONO : STEREO; e.dim = mono ? 2 : 3;
        e.kf = P.pt_obs_kf[i]; e.lm = P.pt_obs_pt[i];
        e.obs[0] = o[0]; e.obs[1] = o[1]; e.obs[2] = o[2];
        const double is2 = P.pt_obs_inv_sigma2[i];
        e.info[0] = e.info[1] = e.info[2] = is2;
        e.delta = mono ? thHuberMono : thHuberStereo;
        E.push_back(e);
    }
    const int e_line0 = (int)E.size();
    for (int i = 0; i < P.n_line_obs; ++i)
        for (int s = 0; s < 2; ++s) {
            Edge e{};
            e.kind = LINE; e.dim = 3; e.kf = P.line_obs_kf[i]; e.lm = lm_line0 + 2 * P.line_obs_line[i] + s;
            for (int k = 0; k < 3; ++k) { e.obs[k] = P.line_obs_l[3 * i + k]; e.info[k] = 1.0; }
            e.delta = thHuberStereo;
            E.push_back(e);
        }
    int e_plane0[3];
    for (int t = 0; t < 3; ++t) {
        e_plane0[t] = (int)E.size();
        for (int i = 0; i < P.n_plane_obs[t]; ++i) {
            Edge e{};
            e.kind = t == 0 ? PLANE : (t == 1 ? VER : PAR); e.dim = t == 0 ? 3 : 2;
            e.kf = P.plane_obs_kf[t][i]; e.lm = lm_plane0 + P.plane_obs_plane[t][i];
            e.pm = plane_from_float4(P.plane_obs_meas[t] + 4 * i);
            e.info[0] = e.info[1] = angleInfo; e.info[2] = t == 0 ? disInfo : 0.0;
            e.delta = t == 0 ? deltaPlane : VPdeltaPlane;
            E.push_back(e);
        }
    }

    Solver S{poses, lms, E};
    S.initialize();
    out.iterations[0] = S.optimize(5);
    out.trials[0] = S.trials_total; out.chi2[0] = S.last_chi; out.lambda[0] = S.lambda;

    auto depth_positive = [&](const Edge& e) { return se3_map(poses[e.kf].T, lms[e.lm].p).z > 0.0; };
    // gate with the errors left by the last LM trial (no recomputation, src/Optimizer.cc:2355-2455)
    for (int i = 0; i < e_line0; ++i) {
        Edge& e = E[i];
        if (chi2(e) > (e.kind == MONO ? 5.991 : 7.815) || !depth_positive(e)) e.level = 1;
        e.robust = false;
    }
    for (int i = 0; i < P.n_line_obs; ++i) {
        Edge& es = E[e_line0 + 2 * i]; Edge& ee = E[e_line0 + 2 * i + 1];
        if (chi2(es) > 7.815 || chi2(ee) > 7.815) { es.level = 1; ee.level = 1; }
        es.robust = false; ee.robust = false;
    }
    for (int t = 0; t < 3; ++t)
        for (int i = 0; i < P.n_plane_obs[t]; ++i) {
            Edge& e = E[e_plane0[t] + i];
            if (chi2(e) > (t == 0 ? P.plane_chi : P.vp_chi)) e.level = 1;
            e.robust = false;
        }
    S.initialize();
    out.iterations[1] = S.optimize(10);
    out.trials[1] = S.trials_total; out.chi2[1] = S.last_chi; out.lambda[1] = S.lambda;

    // erase lists (:2462-2560): every edge, including the level-1 ones whose error is stale
    out.erase_pt.assign(P.n_pt_obs, 0);
    for (int i = 0; i < P.n_pt_obs; ++i) { const Edge& e = E[i]; out.erase_pt[i] = (chi2(e) > (e.kind == MONO ? 5.991 : 7.815) || !depth_positive(e)) ? 1 : 0; }
    out.erase_line.assign(P.n_line_obs, 0);
    for (int i = 0; i < P.n_line_obs; ++i) out.erase_line[i] = (chi2(E[e_line0 + 2 * i]) > 7.815 || chi2(E[e_line0 + 2 * i + 1]) > 7.815) ? 1 : 0;
    for (int t = 0; t < 3; ++t) {
        out.erase_plane[t].assign(P.n_plane_obs[t], 0);
        for (int i = 0; i < P.n_plane_obs[t]; ++i) out.erase_plane[t][i] = chi2(E[e_plane0[t] + i]) > (t == 0 ? P.plane_chi : P.vp_chi) ? 1 : 0;
    }

    // recover (:2620-2677)
    out.kf_Tcw.assign((size_t)16 * P.n_kf, 0.f); out.kf_Tcw_d.assign((size_t)16 * P.n_kf, 0.0);
    for (int i = 0; i < P.n_kf; ++i) {
        const M3 R = quat_to_matrix(poses[i].T.q);
        double* M = &out.kf_Tcw_d[16 * i];
        for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c) M[r * 4 + c] = R.m[r][c];
        M[3] = poses[i].T.t.x; M[7] = poses[i].T.t.y; M[11] = poses[i].T.t.z; M[15] = 1.0;
        for (int k = 0; k < 16; ++k) out.kf_Tcw[16 * i + k] = (float)M[k];
    }
    out.pt_Xw.resize((size_t)3 * P.n_points); out.pt_Xw_d.resize((size_t)3 * P.n_points);
    for (int i = 0; i < P.n_points; ++i) {
        const double v[3] = {lms[i].p.x, lms[i].p.y, lms[i].p.z};
        for (int k = 0; k < 3; ++k) { out.pt_Xw_d[3 * i + k] = v[k]; out.pt_Xw[3 * i + k] = (float)v[k]; }
    }
    out.line_Xw.resize((size_t)6 * P.n_lines); out.line_Xw_d.resize((size_t)6 * P.n_lines);
    for (int i = 0; i < P.n_lines; ++i)
        for (int s = 0; s < 2; ++s) {
            const V3 p = lms[lm_line0 + 2 * i + s].p;
            const double v[3] = {p.x, p.y, p.z};
            for (int k = 0; k < 3; ++k) { out.line_Xw_d[6 * i + 3 * s + k] = v[k]; out.line_Xw[6 * i + 3 * s + k] = (double)(float)v[k]; }
        }
    out.plane_Xw.resize((size_t)4 * P.n_planes); out.plane_Xw_d.resize((size_t)4 * P.n_planes);
    for (int i = 0; i < P.n_planes; ++i)
        for (int k = 0; k < 4; ++k) { out.plane_Xw_d[4 * i + k] = lms[lm_plane0 + i].pl.c[k]; out.plane_Xw[4 * i + k] = (float)lms[lm_plane0 + i].pl.c[k]; }
}

}  // namespace oracle
