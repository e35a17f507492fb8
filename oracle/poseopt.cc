// TEST INFRASTRUCTURE ONLY — CPU oracle restating Optimizer::PoseOptimization and the g2o pieces it uses
// (see poseopt.h for the file:line map).
#include "poseopt.h"
#include "g2o_math.h"

#include <cfloat>
#include <cmath>
#include <cstring>

namespace oracle {
namespace {

using namespace gm;

// ---- edges ----
enum Kind { MONO, STEREO, LINE, PLANE, PAR, VER, MONO_T, STEREO_T, LINE_T, PLANE_T };
struct Edge {
    Kind kind;
    int dim;                 // residual dimension
    int idx;                 // index in its family (point i, line i, plane i)
    V3 Xw;                   // point edges
    double obs[3];
    double info[3];          // diagonal information
    double delta;            // Huber delta (float-rounded like the reference's const float)
    Plane pw, pm;            // plane edges: world plane, measurement
    bool robust = true;
    int level = 0;
    double err[3] = {0, 0, 0};
};

struct Cam { double fx, fy, cx, cy, bf; };

void compute_error(Edge& e, const SE3& T, const Cam& K) {
    switch (e.kind) {
        case MONO: {
            const V3 p = se3_map(T, e.Xw);
            e.err[0] = e.obs[0] - (p.x / p.z * K.fx + K.cx);
            e.err[1] = e.obs[1] - (p.y / p.z * K.fy + K.cy);
            break;
        }
        case STEREO: {
            const V3 p = se3_map(T, e.Xw);
            const float invz = 1.0f / (float)p.z;                                 // sic: float reciprocal, .cpp:300
            const double r0 = p.x * invz * K.fx + K.cx, r1 = p.y * invz * K.fy + K.cy, r2 = r0 - K.bf * invz;
            e.err[0] = e.obs[0] - r0; e.err[1] = e.obs[1] - r1; e.err[2] = e.obs[2] - r2;
            break;
        }
        case LINE: {
            const V3 p = se3_map(T, e.Xw);
            const double u = p.x / p.z * K.fx + K.cx, v = p.y / p.z * K.fy + K.cy;
            e.err[0] = e.obs[0] * u + e.obs[1] * v + e.obs[2]; e.err[1] = 0; e.err[2] = 0;
            break;
        }
        case MONO_T: {
            const V3 p = e.Xw + T.t;                                              // mapTrans: Xc + t
            e.err[0] = e.obs[0] - (p.x / p.z * K.fx + K.cx);
            e.err[1] = e.obs[1] - (p.y / p.z * K.fy + K.cy);
            break;
        }
        case STEREO_T: {
            const V3 p = e.Xw + T.t;
            const float invz = 1.0f / (float)p.z;
            const double r0 = p.x * invz * K.fx + K.cx, r1 = p.y * invz * K.fy + K.cy, r2 = r0 - K.bf * invz;
            e.err[0] = e.obs[0] - r0; e.err[1] = e.obs[1] - r1; e.err[2] = e.obs[2] - r2;
            break;
        }
        case LINE_T: {
            const V3 p = e.Xw + T.t;
            const double u = p.x / p.z * K.fx + K.cx, v = p.y / p.z * K.fy + K.cy;
            e.err[0] = e.obs[0] * u + e.obs[1] * v + e.obs[2]; e.err[1] = 0; e.err[2] = 0;
            break;
        }
        case PLANE_T: {                                                           // operator+(Isometry3D, Plane3D), Plane3D.h:201-209
            Plane r{{e.pw.c[0], e.pw.c[1], e.pw.c[2], e.pw.c[3] - (T.t.x * e.pw.c[0] + T.t.y * e.pw.c[1] + T.t.z * e.pw.c[2])}};
            if (r.c[3] < 0.0) for (int i = 0; i < 4; ++i) r.c[i] = -r.c[i];
            plane_normalize(r);
            plane_ominus(r, e.pm, e.err);
            break;
        }
        case PLANE: plane_ominus(plane_transform(T, e.pw), e.pm, e.err); break;
        case PAR: plane_ominus_par(plane_transform(T, e.pw), e.pm, e.err); break;
        case VER: plane_ominus_ver(plane_transform(T, e.pw), e.pm, e.err); break;
    }
}
inline double chi2(const Edge& e) { double s = 0; for (int i = 0; i < e.dim; ++i) s += e.err[i] * e.info[i] * e.err[i]; return s; }

void jacobian(Edge& e, const SE3& T, const Cam& K, double J[3][6]) {
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 6; ++j) J[i][j] = 0;
    if (e.kind == MONO || e.kind == STEREO || e.kind == LINE) {
        const V3 p = se3_map(T, e.Xw);
        const double x = p.x, y = p.y, invz = 1.0 / p.z, invz_2 = invz * invz;
        if (e.kind == LINE) {
            const double lx = e.obs[0], ly = e.obs[1], fx = K.fx, fy = K.fy;
            J[0][0] = -fy * ly - fx * lx * x * y * invz_2 - fy * ly * y * y * invz_2;
            J[0][1] = fx * lx + fx * lx * x * x * invz_2 + fy * ly * x * y * invz_2;
            J[0][2] = -fx * lx * y * invz + fy * ly * x * invz;
            J[0][3] = fx * lx * invz;
            J[0][4] = fy * ly * invz;
            J[0][5] = -(fx * lx * x + fy * ly * y) * invz_2;
            return;
        }
        J[0][0] = x * y * invz_2 * K.fx; J[0][1] = -(1 + (x * x * invz_2)) * K.fx; J[0][2] = y * invz * K.fx;
        J[0][3] = -invz * K.fx; J[0][4] = 0; J[0][5] = x * invz_2 * K.fx;
        J[1][0] = (1 + y * y * invz_2) * K.fy; J[1][1] = -x * y * invz_2 * K.fy; J[1][2] = -x * invz * K.fy;
        J[1][3] = 0; J[1][4] = -invz * K.fy; J[1][5] = y * invz_2 * K.fy;
        if (e.kind == STEREO) {
            J[2][0] = J[0][0] - K.bf * y * invz_2; J[2][1] = J[0][1] + K.bf * x * invz_2; J[2][2] = J[0][2];
            J[2][3] = J[0][3]; J[2][4] = 0; J[2][5] = J[0][5] - K.bf * invz_2;
        }
        return;
    }
    if (e.kind == MONO_T || e.kind == STEREO_T || e.kind == LINE_T) {
        const V3 p = e.Xw + T.t;
        const double x = p.x, y = p.y, invz = 1.0 / p.z, invz_2 = invz * invz;
        if (e.kind == LINE_T) {
            const double lx = e.obs[0], ly = e.obs[1];
            J[0][3] = K.fx * lx * invz; J[0][4] = K.fy * ly * invz; J[0][5] = -(K.fx * lx * x + K.fy * ly * y) * invz_2;
            return;
        }
        J[0][3] = -invz * K.fx; J[0][5] = x * invz_2 * K.fx;
        J[1][4] = -invz * K.fy; J[1][5] = y * invz_2 * K.fy;
        if (e.kind == STEREO_T) { J[2][3] = J[0][3]; J[2][5] = J[0][5] - K.bf * invz_2; }
        return;
    }
    // numeric central differences, delta = 1e-9 (base_unary_edge.hpp:94-116); _error is restored afterwards
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    double keep[3] = {e.err[0], e.err[1], e.err[2]};
    for (int d = 0; d < 6; ++d) {
        double add[6] = {0, 0, 0, 0, 0, 0};
        add[d] = delta;
        compute_error(e, se3_mul(se3_exp(add), T), K);
        const double e1[3] = {e.err[0], e.err[1], e.err[2]};
        add[d] = -delta;
        compute_error(e, se3_mul(se3_exp(add), T), K);
        for (int i = 0; i < e.dim; ++i) J[i][d] = scalar * (e1[i] - e.err[i]);
    }
    for (int i = 0; i < 3; ++i) e.err[i] = keep[i];
    if (e.kind == PLANE_T) for (int i = 0; i < 3; ++i) J[i][0] = J[i][1] = J[i][2] = 0;   // EdgePlane.h:292-308
}

inline void huber(double e2, double delta, double rho[3]) {
    const double dsqr = delta * delta;
    if (e2 <= dsqr) { rho[0] = e2; rho[1] = 1.; rho[2] = 0.; }
    else { const double s = std::sqrt(e2); rho[0] = 2 * s * delta - dsqr; rho[1] = delta / s; rho[2] = -0.5 * rho[1] / e2; }
}

// unpivoted LDL^T of a symmetric 6x6; false when a pivot is not positive (LDLT::isPositive)
bool solve6(const double H[6][6], const double b[6], double x[6]) {
    double L[6][6] = {{0}}, D[6];
    for (int j = 0; j < 6; ++j) {
        double d = H[j][j];
        for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k] * D[k];
        if (!(d > 0)) return false;
        D[j] = d;
        L[j][j] = 1;
        for (int i = j + 1; i < 6; ++i) {
            double v = H[i][j];
            for (int k = 0; k < j; ++k) v -= L[i][k] * L[j][k] * D[k];
            L[i][j] = v / d;
        }
    }
    double y[6];
    for (int i = 0; i < 6; ++i) { double v = b[i]; for (int k = 0; k < i; ++k) v -= L[i][k] * y[k]; y[i] = v; }
    for (int i = 0; i < 6; ++i) y[i] /= D[i];
    for (int i = 5; i >= 0; --i) { double v = y[i]; for (int k = i + 1; k < 6; ++k) v -= L[k][i] * x[k]; x[i] = v; }
    return true;
}

struct Lm {
    std::vector<Edge>& E;
    const Cam& K;
    SE3 T;
    double lambda = 0, ni = 2;
    int nBad = 0;
    double H[6][6], b[6], x[6] = {0, 0, 0, 0, 0, 0};
    int last_trials = 0;
    double last_chi = 0;

    void active_errors() { for (Edge& e : E) if (e.level == 0) compute_error(e, T, K); }
    double robust_chi2() const {
        double chi = 0;
        for (const Edge& e : E) if (e.level == 0) {
            const double c = chi2(e);
            if (e.robust) { double rho[3]; huber(c, e.delta, rho); chi += rho[0]; } else chi += c;
        }
        return chi;
    }
    void build_system() {
        for (int i = 0; i < 6; ++i) { b[i] = 0; for (int j = 0; j < 6; ++j) H[i][j] = 0; }
        for (Edge& e : E) if (e.level == 0) {
            double J[3][6];
            jacobian(e, T, K, J);
            double w = 1.0;
            if (e.robust) { double rho[3]; huber(chi2(e), e.delta, rho); w = rho[1]; }
            for (int r = 0; r < e.dim; ++r) {
                const double oe = e.info[r] * e.err[r];
                for (int i = 0; i < 6; ++i) {
                    b[i] -= w * J[r][i] * oe;
                    const double wi = w * e.info[r] * J[r][i];
                    for (int j = 0; j < 6; ++j) H[i][j] += wi * J[r][j];
                }
            }
        }
    }
    // returns 0 OK, 1 Terminate
    int solve(int iteration) {
        active_errors();
        double currentChi = robust_chi2(), tempChi = currentChi;
        const double iniChi = currentChi;
        build_system();
        if (iteration == 0) {
            double mx = 0;
            for (int j = 0; j < 6; ++j) mx = std::max(std::fabs(H[j][j]), mx);
            lambda = 1e-5 * mx; ni = 2; nBad = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
            const SE3 backup = T;
            double Hl[6][6];
            for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) Hl[i][j] = H[i][j] + (i == j ? lambda : 0.0);
            const bool ok2 = solve6(Hl, b, x);
            // on failure g2o leaves x untouched (the previous solution, zero at the very start) and still applies it
            T = se3_mul(se3_exp(x), T);
            active_errors();
            tempChi = robust_chi2();
            if (!ok2) tempChi = DBL_MAX;
            rho = currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6; ++j) scale += x[j] * (lambda * x[j] + b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && std::isfinite(tempChi)) {
                double alpha = 1. - std::pow((2 * rho - 1), 3);
                alpha = std::min(alpha, 2. / 3.);
                const double sf = std::max(1. / 3., alpha);
                lambda *= sf; ni = 2; currentChi = tempChi;
            } else {
                lambda *= ni; ni *= 2; T = backup;
            }
            ++qmax;
        } while (rho < 0 && qmax < 10);
        last_trials += qmax;
        last_chi = currentChi;
        if (qmax == 10 || rho == 0) return 1;
        if ((iniChi - currentChi) * 1e3 < iniChi) ++nBad; else nBad = 0;
        if (nBad >= 3) return 1;
        return 0;
    }
    int optimize(int iterations) {
        int done = 0;
        last_trials = 0;
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

void pose_optimization(const PoseProblem& P, const float* Tcw_in, PoseResult& out) {
    const Cam K{P.fx, P.fy, P.cx, P.cy, P.bf};
    std::vector<Edge> E;
    const float deltaMono = std::sqrt(5.991), deltaStereo = std::sqrt(7.815);
    double angleInfo = 3282.8 / (P.angle_info * P.angle_info), disInfo = P.dist_info * P.dist_info;
    double parInfo = 3282.8 / (P.par_info * P.par_info), verInfo = 3282.8 / (P.ver_info * P.ver_info);
    const float deltaPlane = std::sqrt(P.plane_chi), VPdeltaPlane = std::sqrt(P.vp_chi);
    out.outlier_pt.assign(P.n_points, 0); out.outlier_line.assign(P.n_lines, 0); out.outlier_plane.assign(P.n_planes, 0);
    out.outlier_par.assign(P.n_par, 0); out.outlier_ver.assign(P.n_ver, 0);
    int nInitial = 0;
    for (int i = 0; i < P.n_points; ++i) {
        Edge e;
        const bool mono = P.obs[3 * i + 2] < 0;
        e.kind = mono ? MONO : STEREO; e.dim = mono ? 2 : 3; e.idx = i;
        e.Xw = {P.Xw[3 * i], P.Xw[3 * i + 1], P.Xw[3 * i + 2]};
        for (int k = 0; k < 3; ++k) { e.obs[k] = P.obs[3 * i + k]; e.info[k] = P.inv_sigma2[i]; }
        e.delta = mono ? deltaMono : deltaStereo;
        E.push_back(e);
        ++nInitial;
    }
    for (int i = 0; i < P.n_lines; ++i)
        for (int s = 0; s < 2; ++s) {
            Edge e;
            e.kind = LINE; e.dim = 3; e.idx = i;
            e.Xw = {P.line_Xw[6 * i + 3 * s], P.line_Xw[6 * i + 3 * s + 1], P.line_Xw[6 * i + 3 * s + 2]};
            for (int k = 0; k < 3; ++k) { e.obs[k] = P.line_obs[3 * i + k]; e.info[k] = 1.0; }
            e.delta = deltaStereo;
            E.push_back(e);
            if (s == 0) ++nInitial;
        }
    auto add_plane = [&](Kind kd, int n, const float* meas, const float* map, double i0, double i1, double i2, double delta) {
        for (int i = 0; i < n; ++i) {
            Edge e;
            e.kind = kd; e.dim = (kd == PLANE) ? 3 : 2; e.idx = i;
            e.pm = plane_from_float4(meas + 4 * i); e.pw = plane_from_float4(map + 4 * i);
            e.info[0] = i0; e.info[1] = i1; e.info[2] = i2; e.delta = delta;
            e.obs[0] = e.obs[1] = e.obs[2] = 0; e.Xw = {0, 0, 0};
            E.push_back(e);
            ++nInitial;
        }
    };
    add_plane(PLANE, P.n_planes, P.plane_meas, P.plane_map, angleInfo, angleInfo, disInfo, deltaPlane);
    add_plane(PAR, P.n_par, P.par_meas, P.par_map, parInfo, parInfo, 0, VPdeltaPlane);
    add_plane(VER, P.n_ver, P.ver_meas, P.ver_map, verInfo, verInfo, 0, VPdeltaPlane);

    const SE3 T0 = se3_from_float16(Tcw_in);
    auto write_pose = [&](const SE3& T) {
        const M3 R = quat_to_matrix(T.q);
        const double tt[3] = {T.t.x, T.t.y, T.t.z};
        for (int i = 0; i < 16; ++i) out.Tcw_d[i] = (i == 15) ? 1.0 : 0.0;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out.Tcw_d[i * 4 + j] = R.m[i][j]; out.Tcw_d[i * 4 + 3] = tt[i]; }
        for (int i = 0; i < 16; ++i) out.Tcw[i] = (float)out.Tcw_d[i];
    };
    out.n_rounds = 0;
    if (nInitial < 3) { out.n_inliers = 0; for (int i = 0; i < 16; ++i) { out.Tcw[i] = Tcw_in[i]; out.Tcw_d[i] = Tcw_in[i]; } return; }   // return 0, mTcw untouched
    // the reference evaluates computeError() once on every plane edge while building the graph (:896, :935, :975)
    for (Edge& e : E) if (e.kind == PLANE || e.kind == PAR || e.kind == VER) compute_error(e, T0, K);

    const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
    Lm lm{E, K, T0};
    int nBad = 0;
    for (int it = 0; it < 4; ++it) {
        lm.T = T0;
        const int iters = lm.optimize(10);
        nBad = 0;
        for (size_t k = 0; k < E.size(); ++k) {
            Edge& e = E[k];
            if (e.kind == MONO || e.kind == STEREO) {
                if (out.outlier_pt[e.idx]) compute_error(e, lm.T, K);
                const float c = (float)chi2(e);
                if (c > (e.kind == MONO ? chi2Mono : chi2Stereo)) { out.outlier_pt[e.idx] = 1; e.level = 1; ++nBad; }
                else { out.outlier_pt[e.idx] = 0; e.level = 0; }
            } else if (e.kind == LINE) {
                Edge& e2 = E[k + 1];                       // start / end edges are adjacent
                compute_error(e, lm.T, K); compute_error(e2, lm.T, K);
                const float cs = (float)(e.err[0] * e.err[0]), ce = (float)(e2.err[0] * e2.err[0]);
                if (cs > 2 * chi2Mono || ce > 2 * chi2Mono) { out.outlier_line[e.idx] = 1; e.level = e2.level = 1; ++nBad; }
                else { out.outlier_line[e.idx] = 0; e.level = e2.level = 0; }
                if (it == 2) e2.robust = false;
                ++k;
            } else {
                std::vector<uint8_t>& flags = e.kind == PLANE ? out.outlier_plane : (e.kind == PAR ? out.outlier_par : out.outlier_ver);
                if (flags[e.idx]) compute_error(e, lm.T, K);
                const float c = (float)chi2(e);
                const double th = e.kind == PLANE ? P.plane_chi : P.vp_chi;
                if (c > th) { flags[e.idx] = 1; e.level = 1; ++nBad; } else { flags[e.idx] = 0; e.level = 0; }
            }
            if (it == 2) e.robust = false;
        }
        out.rounds[it] = {iters, lm.last_trials, lm.last_chi, lm.lambda, nBad};
        out.n_rounds = it + 1;
        if (E.size() < 10) break;
    }
    write_pose(lm.T);
    out.n_inliers = nInitial - nBad;
}


void translation_optimization(const PoseProblem& P, const float* Tcw_in, PoseResult& out) {
    const Cam K{P.fx, P.fy, P.cx, P.cy, P.bf};
    std::vector<Edge> E;
    const float deltaMono = std::sqrt(5.991), deltaStereo = std::sqrt(7.815);
    const double angleInfo = 3282.8 / (P.angle_info * P.angle_info), disInfo = P.dist_info * P.dist_info;
    const float deltaPlane = std::sqrt(P.plane_chi);
    out.outlier_pt.assign(P.n_points, 0); out.outlier_line.assign(P.n_lines, 0); out.outlier_plane.assign(P.n_planes, 0);
    out.outlier_par.assign(P.n_par, 0); out.outlier_ver.assign(P.n_ver, 0);
    // R_cw * X in float matrices (cv::Mat CV_32F product: double accumulation, float result)
    auto rot_f = [&](double x, double y, double z, int row) -> float {
        return (float)((double)Tcw_in[row * 4 + 0] * (double)(float)x + (double)Tcw_in[row * 4 + 1] * (double)(float)y + (double)Tcw_in[row * 4 + 2] * (double)(float)z);
    };
    int nInitial = 0;
    for (int i = 0; i < P.n_points; ++i) {
        Edge e;
        const bool mono = P.obs[3 * i + 2] < 0;
        e.kind = mono ? MONO_T : STEREO_T; e.dim = mono ? 2 : 3; e.idx = i;
        e.Xw = {rot_f(P.Xw[3 * i], P.Xw[3 * i + 1], P.Xw[3 * i + 2], 0), rot_f(P.Xw[3 * i], P.Xw[3 * i + 1], P.Xw[3 * i + 2], 1),
                rot_f(P.Xw[3 * i], P.Xw[3 * i + 1], P.Xw[3 * i + 2], 2)};
        for (int k = 0; k < 3; ++k) { e.obs[k] = P.obs[3 * i + k]; e.info[k] = P.inv_sigma2[i]; }
        e.delta = mono ? deltaMono : deltaStereo;
        E.push_back(e);
        ++nInitial;
    }
    for (int i = 0; i < P.n_lines; ++i)
        for (int s = 0; s < 2; ++s) {
            Edge e;
            e.kind = LINE_T; e.dim = 3; e.idx = i;
            const double* X = P.line_Xw + 6 * i + 3 * s;      // Converter::toCvVec: double -> float before the product
            e.Xw = {rot_f(X[0], X[1], X[2], 0), rot_f(X[0], X[1], X[2], 1), rot_f(X[0], X[1], X[2], 2)};
            for (int k = 0; k < 3; ++k) { e.obs[k] = P.line_obs[3 * i + k]; e.info[k] = 1.0; }
            e.delta = deltaStereo;
            E.push_back(e);
        }
    const SE3 T0 = se3_from_float16(Tcw_in);
    auto write_pose = [&](const SE3& T) {
        const M3 R = quat_to_matrix(T.q);
        const double tt[3] = {T.t.x, T.t.y, T.t.z};
        for (int i = 0; i < 16; ++i) out.Tcw_d[i] = (i == 15) ? 1.0 : 0.0;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) out.Tcw_d[i * 4 + j] = R.m[i][j]; out.Tcw_d[i * 4 + 3] = tt[i]; }
        for (int i = 0; i < 16; ++i) out.Tcw[i] = (float)out.Tcw_d[i];
    };
    out.n_rounds = 0;
    if (nInitial < 3) { out.n_inliers = 0; for (int i = 0; i < 16; ++i) { out.Tcw[i] = Tcw_in[i]; out.Tcw_d[i] = Tcw_in[i]; } return; }   // return 0, mTcw untouched
    for (int i = 0; i < P.n_planes; ++i) {
        Edge e;
        e.kind = PLANE_T; e.dim = 3; e.idx = i;
        e.pm = plane_from_float4(P.plane_meas + 4 * i);
        Plane pw = plane_from_float4(P.plane_map + 4 * i);
        // Xw.rotateNormal(Converter::toMatrix3d(R_cw)): the float rotation widened to double, no renormalisation
        const double n[3] = {pw.c[0], pw.c[1], pw.c[2]};
        for (int r = 0; r < 3; ++r) pw.c[r] = (double)Tcw_in[r * 4 + 0] * n[0] + (double)Tcw_in[r * 4 + 1] * n[1] + (double)Tcw_in[r * 4 + 2] * n[2];
        e.pw = pw;
        e.info[0] = angleInfo; e.info[1] = angleInfo; e.info[2] = disInfo; e.delta = deltaPlane;
        e.obs[0] = e.obs[1] = e.obs[2] = 0; e.Xw = {0, 0, 0};
        compute_error(e, T0, K);                           // :3300-3305
        E.push_back(e);
    }
    const float chi2Mono = 5.991f, chi2Stereo = 7.815f;
    Lm lm{E, K, T0};
    int nBad = 0;
    for (int it = 0; it < 4; ++it) {
        lm.T = T0;
        const int iters = lm.optimize(10);
        nBad = 0;
        for (size_t k = 0; k < E.size(); ++k) {
            Edge& e = E[k];
            if (e.kind == MONO_T || e.kind == STEREO_T) {
                if (out.outlier_pt[e.idx]) compute_error(e, lm.T, K);
                const float c = (float)chi2(e);
                if (c > (e.kind == MONO_T ? chi2Mono : chi2Stereo)) { out.outlier_pt[e.idx] = 1; e.level = 1; ++nBad; }
                else { out.outlier_pt[e.idx] = 0; e.level = 0; }
            } else if (e.kind == LINE_T) {
                Edge& e2 = E[k + 1];
                if (out.outlier_line[e.idx]) { compute_error(e, lm.T, K); compute_error(e2, lm.T, K); }
                const float cs = (float)(e.err[0] * e.err[0]), ce = (float)(e2.err[0] * e2.err[0]);
                if (cs > 2 * chi2Mono || ce > 2 * chi2Mono) { out.outlier_line[e.idx] = 1; e.level = e2.level = 1; }   // nLineBad, not nBad
                else { out.outlier_line[e.idx] = 0; e.level = e2.level = 0; }
                if (it == 2) e2.robust = false;
                ++k;
            } else {
                if (out.outlier_plane[e.idx]) compute_error(e, lm.T, K);
                const float c = (float)chi2(e);
                if (c > P.plane_chi) { out.outlier_plane[e.idx] = 1; e.level = 1; ++nBad; } else { out.outlier_plane[e.idx] = 0; e.level = 0; }
            }
            if (it == 2) e.robust = false;
        }
        out.rounds[it] = {iters, lm.last_trials, lm.last_chi, lm.lambda, nBad};
        out.n_rounds = it + 1;
        if (E.size() < 10) break;
    }
    write_pose(lm.T);
    out.n_inliers = nInitial - nBad;
}

}  // namespace oracle
