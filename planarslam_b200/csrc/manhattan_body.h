// Manhattan-frame tracking step - Tracking::TrackManhattanFrame (src/Tracking.cc:963-1137) with ProjectSN2Conic (:888-961),
// ProjectSN2MF (:763-886) and MeanShift (:1139-1157).
//
// One thread runs one frame (six passes over <= ~8 600 surface normals + <= 40 line directions, sums in the reference's order),
// written once for host and device so that the CPU suite can check it against the oracle (tests/test_manhattan_host.py).
// Kept from the reference: R_cm aliases R_cm_update (:970) - the columns written for axis 1 are read when the cones of axes 2 and 3
// are rebuilt in the second pass, and with fewer than two directions the partially updated matrix is returned without the SVD.
// sin(0.2018), sin(0.1018), sin(0.2518) are spelled as the doubles glibc returns; asin / exp / tanf come from the platform libm
// (CUDA's on the device: last-ulp differences, absorbed by the float results).
#pragma once
#include <float.h>
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MH_HD __host__ __device__
#else
#define MH_HD
#endif

#define MH_SIN_2018 0.2004331278383219
#define MH_SIN_1018 0.10162426144709341
#define MH_SIN_2518 0.24914759987751361

struct MhResult {                      // = pslam_manhattan_result
    float R[9];
    float density[3];
    int32_t found[3], n_cone[3], n_selected[3], min_num, svd_applied;
};

// row i of the matrix the reference passes as R_mc_new = column c_i of R_cm, c = ((a+3)%3, (a+4)%3, (a+5)%3)
MH_HD inline void mh_axis_rotation(const float* R, int a, float* T) {
    const int c[3] = {(a + 3) % 3, (a + 4) % 3, (a + 5) % 3};
    for (int i = 0; i < 3; ++i)
        for (int r = 0; r < 3; ++r) T[3 * i + r] = R[3 * r + c[i]];
}
MH_HD inline void mh_rotate_normal(const float* T, const float* v, float* q) {          // float products, float sums
    for (int r = 0; r < 3; ++r) q[r] = T[3 * r] * v[0] + T[3 * r + 1] * v[1] + T[3 * r + 2] * v[2];
}
MH_HD inline void mh_rotate_dir(const float* T, const double* v, float* q) {            // float * double, double sums, stored as float
    for (int r = 0; r < 3; ++r) q[r] = (float)(T[3 * r] * v[0] + T[3 * r + 1] * v[1] + T[3 * r + 2] * v[2]);
}
MH_HD inline double mh_lambda(const float* q) { return (double)sqrtf(q[0] * q[0] + q[1] * q[1]); }

struct MhShift { double nx, ny, den; int count; };

// the body of the ProjectSN2MF loop for one rotated vector; returns true when it is inside the 0.2518 cone
MH_HD inline bool mh_consider(const float* q, MhShift& S) {
    const double lambda = mh_lambda(q);
    if (!(lambda < MH_SIN_2518)) return false;
    const double tan_alfa = lambda / fabsf(q[2]);
    const double alfa = asin(lambda);
    const double mx = alfa / tan_alfa * q[0] / q[2], my = alfa / tan_alfa * q[1] / q[2];
    if (!(mx != mx) && !(my != my)) {
        const double nr = sqrt(mx * mx + my * my);
        const double k = exp(-20 * nr * nr);
        S.nx += k * mx; S.ny += k * my; S.den += k;
        ++S.count;
    }
    return true;
}

// cv::SVD of a 3x3 float matrix (OpenCV's Jacobi, float flavour: eps = 2 FLT_EPSILON): A row-major in, U (3x3) and Vt out
MH_HD inline void mh_svd3f(const float* A, float* U, float* Vt) {
    float At[9];
    double W[3];
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < 3; ++k) { At[3 * i + k] = A[3 * k + i]; sd += (double)At[3 * i + k] * At[3 * i + k]; }
        W[i] = sd;
        for (int k = 0; k < 3; ++k) Vt[3 * i + k] = (i == k) ? 1.f : 0.f;
    }
    const float eps = FLT_EPSILON * 2;
    for (int iter = 0; iter < 30; ++iter) {
        bool changed = false;
        for (int i = 0; i < 2; ++i)
            for (int j = i + 1; j < 3; ++j) {
                float* Ai = At + 3 * i;
                float* Aj = At + 3 * j;
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < 3; ++k) p += (double)Ai[k] * Aj[k];
                if (fabs(p) <= eps * sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = sqrt(p * p + beta * beta);
                float c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = (float)sqrt(delta / gamma);
                    c = (float)(p / (gamma * s * 2));
                } else {
                    c = (float)sqrt((gamma + beta) / (gamma * 2));
                    s = (float)(p / (gamma * c * 2));
                }
                a = b = 0;
                for (int k = 0; k < 3; ++k) {
                    const float t0 = c * Ai[k] + s * Aj[k];
                    const float t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += (double)t0 * t0; b += (double)t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                float* Vi = Vt + 3 * i;
                float* Vj = Vt + 3 * j;
                for (int k = 0; k < 3; ++k) {
                    const float t0 = c * Vi[k] + s * Vj[k];
                    const float t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < 3; ++i) {
        double sd = 0;
        for (int k = 0; k < 3; ++k) sd += (double)At[3 * i + k] * At[3 * i + k];
        W[i] = sqrt(sd);
    }
    for (int i = 0; i < 2; ++i) {
        int j = i;
        for (int k = i + 1; k < 3; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            const double tw = W[i]; W[i] = W[j]; W[j] = tw;
            for (int k = 0; k < 3; ++k) { float t = At[3 * i + k]; At[3 * i + k] = At[3 * j + k]; At[3 * j + k] = t; t = Vt[3 * i + k]; Vt[3 * i + k] = Vt[3 * j + k]; Vt[3 * j + k] = t; }
        }
    }
    for (int i = 0; i < 3; ++i) {
        const float s = (float)(W[i] > (double)FLT_MIN ? 1 / W[i] : 0.);
        for (int k = 0; k < 3; ++k) U[3 * k + i] = At[3 * i + k] * s;
    }
}

// normals [n][3] float, dirs [m][3] double.  nmask [n] / dmask [m] are written: bit a-1 = appended to vSurfaceNormal{x,y,z} /
// vVanishingLine{x,y,z} by the second pass, bit 3+a = inside the first-pass cone of axis a (ProjectSN2Conic's SNVector / Linesvector).
MH_HD inline void mh_track(const float* R_last, const float* normals, int n, const double* dirs, int m, MhResult& res, uint8_t* nmask, uint8_t* dmask) {
    float R[9];
    for (int i = 0; i < 9; ++i) { R[i] = R_last[i]; res.R[i] = 0; }
    for (int a = 0; a < 3; ++a) { res.density[a] = 0; res.found[a] = 0; res.n_cone[a] = 0; res.n_selected[a] = 0; }
    res.svd_applied = 0;
    for (int i = 0; i < n; ++i) nmask[i] = 0;
    for (int i = 0; i < m; ++i) dmask[i] = 0;
    for (int a = 1; a < 4; ++a) {                      // ProjectSN2Conic
        float T[9], q[3];
        mh_axis_rotation(R, a, T);
        int cnt = 0;
        for (int i = 0; i < n; ++i) {
            mh_rotate_normal(T, normals + 3 * i, q);
            if (mh_lambda(q) < MH_SIN_2018) { nmask[i] |= (uint8_t)(8 << a); ++cnt; }
        }
        for (int i = 0; i < m; ++i) {
            mh_rotate_dir(T, dirs + 3 * i, q);
            if (mh_lambda(q) < MH_SIN_1018) dmask[i] |= (uint8_t)(8 << a);
        }
        res.n_cone[a - 1] = cnt;
    }
    int minNum = n / 20;
    {
        int a = res.n_cone[0], b = res.n_cone[1], c = res.n_cone[2], t;
        if (a > b) { t = a; a = b; b = t; }
        if (b > c) { t = b; b = c; c = t; }
        if (a > b) { t = a; a = b; b = t; }
        if (b < minNum) minNum = (b + a) / 2;
    }
    res.min_num = minNum;
    int nfound = 0;
    for (int a = 1; a < 4; ++a) {                      // ProjectSN2MF; T is rebuilt from the partly updated R
        float T[9], q[3];
        mh_axis_rotation(R, a, T);
        MhShift S;
        S.nx = 0; S.ny = 0; S.den = 0; S.count = 0;
        for (int i = 0; i < n; ++i) {
            if (!(nmask[i] & (8 << a))) continue;
            mh_rotate_normal(T, normals + 3 * i, q);
            if (mh_consider(q, S)) nmask[i] |= (uint8_t)(1 << (a - 1));
        }
        for (int i = 0; i < m; ++i) {
            if (!(dmask[i] & (8 << a))) continue;
            mh_rotate_dir(T, dirs + 3 * i, q);
            if (mh_consider(q, S)) dmask[i] |= (uint8_t)(1 << (a - 1));
        }
        res.n_selected[a - 1] = S.count;
        if (S.count > minNum) {
            const double sx = S.nx / S.den, sy = S.ny / S.den;
            const float density = (float)(S.den / S.count);
            const float alfa = (float)sqrt(sx * sx + sy * sy);
            const float tr = tanf(alfa) / alfa;
            const float t1[3] = {(float)(tr * sx), (float)(tr * sy), 1.0f};
            float rec[3];
            double nn = 0;
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += (double)T[3 * k + r] * t1[k];
                rec[r] = (float)s;
            }
            for (int r = 0; r < 3; ++r) nn += (double)rec[r] * rec[r];
            const float inv = (float)(1.0 / sqrt(nn));
            for (int r = 0; r < 3; ++r) rec[r] = rec[r] * inv;
            const double sum = (double)rec[0] + (double)rec[1] + (double)rec[2];
            if (sum != 0) {
                ++nfound;
                res.found[a - 1] = 1;
                res.density[a - 1] = density;
                for (int r = 0; r < 3; ++r) R[3 * r + a - 1] = rec[r];
            }
        }
    }
    if (nfound < 2) {
        for (int i = 0; i < 9; ++i) res.R[i] = R[i];
        return;
    }
    if (nfound == 2) {
        int ca, cb, target;
        bool swap;                                     // cross(b, a) instead of cross(a, b)
        if (res.found[0] && res.found[1]) { ca = 0; cb = 1; target = 2; swap = false; }
        else if (res.found[1] && res.found[2]) { ca = 1; cb = 2; target = 0; swap = true; }      // v1 = v3.cross(v2)
        else { ca = 0; cb = 2; target = 1; swap = false; }                                       // v2 = v1.cross(v3)
        float u[3], w[3], x[3];
        for (int r = 0; r < 3; ++r) { u[r] = R[3 * r + (swap ? cb : ca)]; w[r] = R[3 * r + (swap ? ca : cb)]; }
        x[0] = u[1] * w[2] - u[2] * w[1]; x[1] = u[2] * w[0] - u[0] * w[2]; x[2] = u[0] * w[1] - u[1] * w[0];
        for (int r = 0; r < 3; ++r) R[3 * r + target] = x[r];
        const double det = (double)R[0] * ((double)R[4] * R[8] - (double)R[5] * R[7]) - (double)R[1] * ((double)R[3] * R[8] - (double)R[5] * R[6]) +
                           (double)R[2] * ((double)R[3] * R[7] - (double)R[4] * R[6]);
        if (fabs(det + 1) < 0.5)
            for (int r = 0; r < 3; ++r) R[3 * r + target] = -x[r];
    }
    float U[9], Vt[9];
    mh_svd3f(R, U, Vt);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)U[3 * r + k] * Vt[3 * k + c];
            res.R[3 * r + c] = (float)s;
        }
    res.svd_applied = 1;
}
