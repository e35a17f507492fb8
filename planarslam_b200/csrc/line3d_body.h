// Per-frame 3-D line fit - Frame::isLineGood (src/Frame.cc:189-267) with compPt3dCov / extract3dline_mahdist / verify3dLine /
// mah_dist3d_pt_line / computeLine3d_svd (src/LineExtractor.cpp:1157-1470) and random_unique (include/LSDextractor.h:239-251).
//
// One thread runs one frame: the lines of a frame draw from ONE rand() stream in order and the number of draws a line makes
// depends on its data, so the lines of a frame are a sequential chain; frames are independent.  Everything is plain IEEE double
// arithmetic (+, -, *, /, sqrt, floor) in a fixed order, written once for host and device: nvcc (--fmad=false) and g++
// (-ffp-contract=off) produce the same bits, which is how the CPU suite checks this file against the oracle without a GPU
// (tests/test_line3d_host.py).  Differences to the reference, all at rounding level: cv::SVD is OpenCV's Jacobi algorithm with
// hypot(p, b) evaluated as sqrt(p*p + b*b) (no libm on the device; operands are far from over/underflow here); rand() is glibc's
// TYPE_3 generator restated (bit-exact, tested against libc), one stream per frame (seed, skip).
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define L3D_HD __host__ __device__
#else
#define L3D_HD
#endif

#define L3D_MAX_PTS 51

struct L3dKeyLine {                      // cv::line_descriptor::KeyLine, 68 bytes (= pslam_keyline)
    float angle; int32_t class_id; int32_t octave; float pt_x, pt_y; float response; float size;
    float startPointX, startPointY, endPointX, endPointY, sPointInOctaveX, sPointInOctaveY, ePointInOctaveX, ePointInOctaveY;
    float lineLength; int32_t numOfPixels;
};

struct L3dPoint { double pos[3]; double DU[9]; };

struct L3dRand { int32_t state[31]; int f, b; int32_t drawn; };
#define L3D_JUMP_ABOVE 4096        // draws to discard above which l3d_srand jumps ahead instead of stepping

L3D_HD inline int32_t l3d_rand(L3dRand& g) {                 // glibc random_r, TYPE_3
    const uint32_t val = (uint32_t)g.state[g.f] + (uint32_t)g.state[g.b];
    g.state[g.f] = (int32_t)val;
    if (++g.f >= 31) g.f = 0;
    if (++g.b >= 31) g.b = 0;
    ++g.drawn;
    return (int32_t)(val >> 1);
}
L3D_HD inline void l3d_srand(L3dRand& g, uint32_t seed, int skip) {       // glibc srandom_r + `skip` discarded draws
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    g.state[0] = word;
    for (int i = 1; i < 31; ++i) {
        const long long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        g.state[i] = word;
    }
    g.f = 3; g.b = 0; g.drawn = 0;
    const long long D = 310LL + (skip > 0 ? skip : 0);            // glibc discards 310 draws after seeding
    if (D <= L3D_JUMP_ABOVE) {
        for (int i = 0; i < (int)D; ++i) (void)l3d_rand(g);
    } else {
        // Jump ahead in O(log D): draw k adds state[k % 31] to state[(k + 3) % 31], so with t_0..t_30 = (s_3 .. s_30, s_0, s_1, s_2) every later value obeys
        // t_{k+31} = t_k + t_{k+28} (mod 2^32): a linear recurrence with characteristic polynomial x^31 = x^28 + 1.  With p = x^D mod that polynomial,
        // t_{D+j} = sum_i p_i t_{i+j}; after D draws the value t_{k+31} sits in state[(3 + k) % 31] for k = D - 31 .. D - 1.  (A replay driver passes the number
        // of draws its sequence has made so far: the loop above would cost that many steps in the one thread that owns the frame.)
        uint32_t t[61];
        for (int i = 0; i < 28; ++i) t[i] = (uint32_t)g.state[i + 3];
        for (int i = 0; i < 3; ++i) t[28 + i] = (uint32_t)g.state[i];
        for (int i = 31; i < 61; ++i) t[i] = t[i - 31] + t[i - 3];
        uint32_t p[31], q[61];
        for (int i = 0; i < 31; ++i) p[i] = 0;
        p[0] = 1;
        int top = 62;
        while (top > 0 && !((D >> (top - 1)) & 1)) --top;
        for (int bit = top - 1; bit >= 0; --bit) {
            for (int i = 0; i < 61; ++i) q[i] = 0;               // q = p * p
            for (int i = 0; i < 31; ++i) { const uint32_t a = p[i]; if (a) for (int j = 0; j < 31; ++j) q[i + j] += a * p[j]; }
            for (int d = 60; d >= 31; --d) { const uint32_t c = q[d]; q[d - 3] += c; q[d - 31] += c; }        // x^d = x^(d-3) + x^(d-31)
            if ((D >> bit) & 1) {                                 // ... * x: the coefficient pushed to x^31 comes back as x^28 + 1
                const uint32_t c = q[30];
                for (int i = 30; i > 0; --i) q[i] = q[i - 1];
                q[0] = c; q[28] += c;
            }
            for (int i = 0; i < 31; ++i) p[i] = q[i];
        }
        for (int j = 0; j < 31; ++j) {
            uint32_t v = 0;
            for (int i = 0; i < 31; ++i) v += p[i] * t[i + j];
            const long long k = D - 31 + j;                       // t_{D+j} = t_{k+31}
            g.state[(int)((3 + k) % 31)] = (int32_t)v;
        }
        g.f = (int)((3 + D) % 31); g.b = (int)(D % 31);
    }
    g.drawn = 0;
}

// cv::SVD of an M x 3 matrix (M >= 3) by OpenCV's one-sided Jacobi: At holds the three columns as rows of length M (stride ld) and is
// overwritten by U^T (rows normalised); W: singular values, descending; Vt: 3 x 3.
L3D_HD inline void l3d_jacobi3(double* At, int M, int ld, double* Wout, double* Vt) {
    double W[3];
    const double eps = 2.220446049250313e-16 * 10, minval = 2.2250738585072014e-308;
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) { const double t = At[i * ld + k]; sd += t * t; }
        W[i] = sd;
        for (int k = 0; k < 3; ++k) Vt[3 * i + k] = (i == k) ? 1.0 : 0.0;
    }
    const int max_iter = M > 30 ? M : 30;
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < 2; ++i)
            for (int j = i + 1; j < 3; ++j) {
                double* Ai = At + i * ld;
                double* Aj = At + j * ld;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < M; ++k) p += Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = sqrt(p * p + beta * beta);
                double c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = sqrt(delta / gamma);
                    c = p / (gamma * s * 2);
                } else {
                    c = sqrt((gamma + beta) / (gamma * 2));
                    s = p / (gamma * c * 2);
                }
                a = b = 0;
                for (int k = 0; k < M; ++k) {
                    const double t0 = c * Ai[k] + s * Aj[k];
                    const double t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += t0 * t0; b += t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                double* Vi = Vt + 3 * i;
                double* Vj = Vt + 3 * j;
                for (int k = 0; k < 3; ++k) {
                    const double t0 = c * Vi[k] + s * Vj[k];
                    const double t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) { const double t = At[i * ld + k]; sd += t * t; }
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < 2; ++i) {
        int j = i;
        for (int k = i + 1; k < 3; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            double t = W[i]; W[i] = W[j]; W[j] = t;
            for (int k = 0; k < M; ++k) { t = At[i * ld + k]; At[i * ld + k] = At[j * ld + k]; At[j * ld + k] = t; }
            for (int k = 0; k < 3; ++k) { t = Vt[3 * i + k]; Vt[3 * i + k] = Vt[3 * j + k]; Vt[3 * j + k] = t; }
        }
    }
    for (int i = 0; i < 3; ++i) {
        Wout[i] = W[i];
        const double s = W[i] > minval ? 1 / W[i] : 0.;
        for (int k = 0; k < M; ++k) At[i * ld + k] *= s;
    }
}

L3D_HD inline void l3d_mat3_mul(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
            c[3 * i + j] = s;
        }
}

// compPt3dCov (src/LineExtractor.cpp:1196-1248): cov = J diag(1, 1, s(z)^2) J^T, SVD, DU = diag(1 / sqrt(w)) U^T
L3D_HD inline void l3d_point_cov(L3dPoint& rp, double f) {
    const double x = rp.pos[0], y = rp.pos[1], z = rp.pos[2];
    const double J0[9] = {z / f, 0, x / z, 0, z / f, y / z, 0, 0, 1};
    const double sd = 0.00273 * z * z + 0.00074 * z + -0.00058;                  // depthStdDev :1182-1195
    const double cg[9] = {1, 0, 0, 0, 1, 0, 0, 0, sd * sd};
    double J0t[9], t[9], cov[9], At[9], w[3], vt[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J0t[3 * i + j] = J0[3 * j + i];
    l3d_mat3_mul(J0, cg, t);
    l3d_mat3_mul(t, J0t, cov);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) At[3 * i + k] = cov[3 * k + i];
    l3d_jacobi3(At, 3, 3, w, vt);                                                // At = U^T
    for (int r = 0; r < 3; ++r) {
        const double inv = 1 / sqrt(w[r]);
        for (int c = 0; c < 3; ++c) rp.DU[3 * r + c] = inv * At[3 * r + c];
    }
}

// mah_dist3d_pt_line (src/LineExtractor.cpp:1418-1470)
L3D_HD inline double l3d_mah_dist(const L3dPoint& pt, const double* q1, const double* q2) {
    const double xa = q1[0], ya = q1[1], za = q1[2], xb = q2[0], yb = q2[1], zb = q2[2];
    const double c1 = pt.DU[0], c2 = pt.DU[1], c3 = pt.DU[2], c4 = pt.DU[3], c5 = pt.DU[4], c6 = pt.DU[5], c7 = pt.DU[6], c8 = pt.DU[7], c9 = pt.DU[8];
    const double x1 = pt.pos[0], x2 = pt.pos[1], x3 = pt.pos[2];
    const double a1 = c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za), b1 = c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb);
    const double a2 = c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za), b2 = c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb);
    const double a3 = c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za), b3 = c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb);
    const double term1 = a1 * b2 - a2 * b1, term2 = a1 * b3 - a3 * b1, term3 = a2 * b3 - a3 * b2;
    const double term4 = (c1 * (x1 - xa) - c1 * (x1 - xb) + c2 * (x2 - ya) - c2 * (x2 - yb) + c3 * (x3 - za) - c3 * (x3 - zb));
    const double term5 = (c4 * (x1 - xa) - c4 * (x1 - xb) + c5 * (x2 - ya) - c5 * (x2 - yb) + c6 * (x3 - za) - c6 * (x3 - zb));
    const double term6 = (c7 * (x1 - xa) - c7 * (x1 - xb) + c8 * (x2 - ya) - c8 * (x2 - yb) + c9 * (x3 - za) - c9 * (x3 - zb));
    return sqrt((term1 * term1 + term2 * term2 + term3 * term3) / (term4 * term4 + term5 * term5 + term6 * term6));
}

L3D_HD inline double l3d_dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

// projectPt3d2Ln3d (src/LineExtractor.cpp:278-286)
L3D_HD inline void l3d_project(const double* P, const double* mid, const double* drct, double* out) {
    double B[3], AB[3], AP[3];
    for (int c = 0; c < 3; ++c) { B[c] = mid[c] + drct[c]; AB[c] = B[c] - mid[c]; AP[c] = P[c] - mid[c]; }
    const double s = l3d_dot3(AB, AP) / l3d_dot3(AB, AB);
    for (int c = 0; c < 3; ++c) out[c] = mid[c] + AB[c] * s;
}

// verify3dLine (src/LineExtractor.cpp:1361-1415) on the points whose bit is set in `set`
L3D_HD inline bool l3d_verify(const L3dPoint* pts, int n, uint64_t set, const double* A, const double* B) {
    double BA[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
    double minv = 100, maxv = -100;
    int idx1 = -1, idx2 = -1, first = -1;
    for (int i = 0; i < n; ++i) {
        if (!((set >> i) & 1)) continue;
        if (first < 0) first = i;
        const double d[3] = {pts[i].pos[0] - A[0], pts[i].pos[1] - A[1], pts[i].pos[2] - A[2]};
        const double v = l3d_dot3(d, BA);
        if (v < minv) { minv = v; idx1 = i; }
        if (v > maxv) { maxv = v; idx2 = i; }
    }
    if (idx1 < 0) idx1 = first;                  // the reference starts both indices at element 0 of the subset
    if (idx2 < 0) idx2 = first;
    const double mid[3] = {(A[0] + B[0]) * 0.5, (A[1] + B[1]) * 0.5, (A[2] + B[2]) * 0.5};
    double C[3], D[3];
    l3d_project(pts[idx1].pos, mid, BA, C);
    l3d_project(pts[idx2].pos, mid, BA, D);
    const double DC[3] = {D[0] - C[0], D[1] - C[1], D[2] - C[2]};
    const double cd = sqrt(DC[0] * DC[0] + DC[1] * DC[1] + DC[2] * DC[2]);
    if (cd < 1e-10) return false;
    unsigned cells = 0;
    for (int i = 0; i < n; ++i) {
        if (!((set >> i) & 1)) continue;
        const double d[3] = {pts[i].pos[0] - C[0], pts[i].pos[1] - C[1], pts[i].pos[2] - C[2]};
        const double lambda = fabs(l3d_dot3(d, DC) / cd / cd);
        if (lambda >= 1) cells |= 1u << 9;
        else cells |= 1u << (unsigned)floor(lambda * 10);
    }
    int sum = 0;
    for (int i = 0; i < 10; ++i) sum += (cells >> i) & 1;
    return (double)sum / 10 > 0.7;
}

L3D_HD inline int l3d_popc64(uint64_t v) { int c = 0; while (v) { v &= v - 1; ++c; } return c; }

struct L3dCam { int w, h; float fx, fy, cx, cy, invfx, invfy, depth_factor; };

struct L3dLineOut {
    uint8_t valid; float depth; double A[3], B[3], director[3]; int32_t n_points, n_inliers; uint64_t inliers;
};

// extract3dline_mahdist (src/LineExtractor.cpp:1265-1359).  At: scratch 3 x L3D_MAX_PTS doubles.
L3D_HD inline void l3d_extract(const L3dPoint* pts, int n, L3dRand& rng, double* At, L3dLineOut& R) {
    const int pairs = (int)(n * (n - 1) * 0.5);
    const int maxIterNo = pairs < 10 ? pairs : 10;
    const double distThresh = 1.5;
    int indexes[L3D_MAX_PTS];
    for (int i = 0; i < n; ++i) indexes[i] = i;
    uint64_t maxSet = 0;
    int maxCount = 0, bestA = 0, bestB = 0;
    for (int iter = 0; iter < maxIterNo; ++iter) {
        {   // random_unique(begin, end, 2): partial Fisher-Yates
            const int r0 = (int)((uint64_t)l3d_rand(rng) % (uint64_t)n);
            int t = indexes[0]; indexes[0] = indexes[r0]; indexes[r0] = t;
            const int r1 = 1 + (int)((uint64_t)l3d_rand(rng) % (uint64_t)(n - 1));
            t = indexes[1]; indexes[1] = indexes[r1]; indexes[r1] = t;
        }
        const double* A = pts[indexes[0]].pos;
        const double* B = pts[indexes[1]].pos;
        const double dAB[3] = {B[0] - A[0], B[1] - A[1], B[2] - A[2]};
        if (sqrt(dAB[0] * dAB[0] + dAB[1] * dAB[1] + dAB[2] * dAB[2]) < 1e-10) continue;
        uint64_t set = 0;
        int count = 0;
        for (int i = 0; i < n; ++i)
            if (l3d_mah_dist(pts[i], A, B) < distThresh) { set |= (uint64_t)1 << i; ++count; }
        if (count > maxCount && l3d_verify(pts, n, set, A, B)) { maxSet = set; maxCount = count; bestA = indexes[0]; bestB = indexes[1]; }
        if ((double)maxCount > n * 0.6) break;
    }
    double rA[3] = {0, 0, 0}, rB[3] = {0, 0, 0};
    if (maxCount >= 2) {
        double m[3], d[3];
        for (int c = 0; c < 3; ++c) { m[c] = (pts[bestA].pos[c] + pts[bestB].pos[c]) * 0.5; d[c] = pts[bestB].pos[c] - pts[bestA].pos[c]; }
        while (true) {
            // computeLine3d_svd (:1157-1179): mean of the inliers, SVD of the centred n x 3 matrix, direction = first row of V^T
            double mean[3] = {0, 0, 0}, w[3], vt[9];
            for (int i = 0; i < n; ++i)
                if ((maxSet >> i) & 1) for (int c = 0; c < 3; ++c) mean[c] = mean[c] + pts[i].pos[c];
            const double inv = 1.0 / maxCount;
            for (int c = 0; c < 3; ++c) mean[c] = mean[c] * inv;
            int q = 0;
            for (int i = 0; i < n; ++i)
                if ((maxSet >> i) & 1) { for (int c = 0; c < 3; ++c) At[c * L3D_MAX_PTS + q] = pts[i].pos[c] - mean[c]; ++q; }
            l3d_jacobi3(At, maxCount, L3D_MAX_PTS, w, vt);             // maxCount >= 8: a verified set occupies more than 7 of 10 cells
            const double e2[3] = {mean[0] + vt[0], mean[1] + vt[1], mean[2] + vt[2]};
            uint64_t set = 0;
            int count = 0;
            for (int i = 0; i < n; ++i)
                if (l3d_mah_dist(pts[i], mean, e2) < distThresh) { set |= (uint64_t)1 << i; ++count; }
            if (count > maxCount) { maxSet = set; maxCount = count; for (int c = 0; c < 3; ++c) { m[c] = mean[c]; d[c] = vt[c]; } }
            else break;
        }
        double minv = 100, maxv = -100;
        int e1 = -1, e2i = -1, first = -1;
        for (int i = 0; i < n; ++i) {
            if (!((maxSet >> i) & 1)) continue;
            if (first < 0) first = i;
            const double dd[3] = {pts[i].pos[0] - m[0], pts[i].pos[1] - m[1], pts[i].pos[2] - m[2]};
            const double dp = l3d_dot3(dd, d);
            if (dp < minv) { minv = dp; e1 = i; }
            if (dp > maxv) { maxv = dp; e2i = i; }
        }
        if (e1 < 0) e1 = first;
        if (e2i < 0) e2i = first;
        for (int c = 0; c < 3; ++c) { rA[c] = pts[e1].pos[c]; rB[c] = pts[e2i].pos[c]; }
    }
    const double ab[3] = {rA[0] - rB[0], rA[1] - rB[1], rA[2] - rB[2]};
    const double nn = sqrt(l3d_dot3(ab, ab));
    for (int c = 0; c < 3; ++c) { R.director[c] = ab[c] / nn; R.A[c] = rA[c]; R.B[c] = rB[c]; }
    R.inliers = maxSet;
    R.n_inliers = maxCount;
}

// Frame::isLineGood for one line.  depth16: raw depth [h][w] of the frame; metres = (float)raw * depth_factor (the reference's
// imDepth.convertTo(CV_32F, mDepthMapFactor)).  pts / At: per-thread scratch (L3D_MAX_PTS points, 3 * L3D_MAX_PTS doubles).
L3D_HD inline void l3d_line(const L3dKeyLine& k, const uint16_t* depth16, const L3dCam& cam, L3dRand& rng, L3dPoint* pts, double* At, L3dLineOut& R) {
    R.valid = 0; R.depth = -1.0f; R.n_points = 0; R.n_inliers = 0; R.inliers = 0;
    for (int c = 0; c < 3; ++c) { R.A[c] = 0; R.B[c] = 0; R.director[c] = 0; }
    const float ddx = k.startPointX - k.endPointX, ddy = k.startPointY - k.endPointY;       // cv::norm(Point2f): float difference, double norm
    const double len = sqrt((double)ddx * ddx + (double)ddy * ddy);
    const int ilen = (int)len;
    const double numSmp = (double)(ilen < 50 ? ilen : 50);
    int n = 0;
    if (numSmp >= 1)                 // numSmp == 0 divides 0 / 0 in the reference (undefined look-up); no line that short leaves the detector
        for (int j = 0; j <= numSmp; ++j) {
            // Point2f * double -> Point2f (product in double, rounded to float); Point2f + Point2f in float; then -> Point2d
            const double t1 = 1 - j / numSmp, t2 = j / numSmp;
            const float px = (float)(k.startPointX * t1) + (float)(k.endPointX * t2);
            const float py = (float)(k.startPointY * t1) + (float)(k.endPointY * t2);
            const double ptx = px, pty = py;
            if (ptx < 0 || pty < 0 || ptx >= cam.w || pty >= cam.h) continue;
            int row, col;
            if (floor(ptx) == ptx && floor(pty) == pty) { col = (int)(ptx - 1); if (col < 0) col = 0; row = (int)(pty - 1); if (row < 0) row = 0; }
            else { col = (int)ptx; row = (int)pty; }
            const float dv = (float)depth16[(size_t)row * cam.w + col] * cam.depth_factor;
            if ((double)dv <= 0.01) continue;
            L3dPoint& p = pts[n++];
            p.pos[2] = dv;
            p.pos[0] = (double)((float)col - cam.cx) * p.pos[2] * (double)cam.invfx;
            p.pos[1] = (double)((float)row - cam.cy) * p.pos[2] * (double)cam.invfy;
        }
    R.n_points = n;
    if (n < 10) return;
    for (int j = 0; j < n; ++j) l3d_point_cov(pts[j], (double)cam.fx);
    l3d_extract(pts, n, rng, At, R);
    const double ab[3] = {R.A[0] - R.B[0], R.A[1] - R.B[1], R.A[2] - R.B[2]};
    if ((double)R.n_inliers / len > 0.4 && sqrt(l3d_dot3(ab, ab)) > 0.02) {
        R.valid = 1;
        // Mat::at<float>(float row, float col): the KeyLine coordinates are truncated to int
        const float de = (float)depth16[(size_t)(int)k.endPointY * cam.w + (int)k.endPointX] * cam.depth_factor;
        const float ds = (float)depth16[(size_t)(int)k.startPointY * cam.w + (int)k.startPointX] * cam.depth_factor;
        R.depth = ds < de ? ds : de;                 // std::min(de, ds)
    } else {
        for (int c = 0; c < 3; ++c) { R.A[c] = 0; R.B[c] = 0; }
    }
}
