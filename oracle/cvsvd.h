// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// cv::SVD::compute (flags 0 / MODIFY_A, i.e. no FULL_UV) as OpenCV's own one-sided Jacobi implements it (core lapack.cpp
// JacobiSVDImpl_ + _SVDcompute; NOT in /root/reference - the reference calls it at src/LineExtractor.cpp:1175,1229 and
// src/Tracking.cc:1102).  Restated from the published algorithm: Hestenes rotations over the columns of A (rows of A^T), sweeps until
// no pair rotates (at most max(m, 30)), singular values = column norms sorted descending, U = normalised columns, V^T = accumulated
// rotations.  PINNING: OpenCV builds that find LAPACK (the in-container cv2 4.13 does: OpenBLAS) route double SVDs to dgesdd
// instead, so this cannot be pinned bit-for-bit here: tests/test_oracle_line3d.py checks it against cv2.SVDecomp to 1e-12 in w and
// up to the sign of each singular-vector pair.  Zero singular values (OpenCV fills the vectors from a fixed-seed RNG) are not
// restated: the vector is left zero.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

namespace oracle {

// A: m x n row-major.  w: min(m,n); u: m x min(m,n) row-major; vt: min(m,n) x n row-major.
template <class T>
void cv_svd(const T* A, int m, int n, T* w, T* u, T* vt) {
    const bool at = m < n;
    int M = m, N = n;
    std::vector<T> At;                 // N x M: row i = column i of the matrix being orthogonalised
    if (at) { std::swap(M, N); At.assign(A, A + (size_t)m * n); }
    else { At.resize((size_t)N * M); for (int i = 0; i < N; ++i) for (int k = 0; k < M; ++k) At[(size_t)i * M + k] = A[(size_t)k * n + i]; }
    std::vector<double> W(N);
    std::vector<T> Vt((size_t)N * N, T(0));
    const T eps = std::is_same<T, float>::value ? T(FLT_EPSILON * 2) : T(DBL_EPSILON * 10);
    const double minval = std::is_same<T, float>::value ? FLT_MIN : DBL_MIN;
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) { const T t = At[(size_t)i * M + k]; sd += (double)t * t; }
        W[i] = sd;
        Vt[(size_t)i * N + i] = 1;
    }
    const int max_iter = std::max(M, 30);
    for (int iter = 0; iter < max_iter; ++iter) {
        bool changed = false;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) {
                T* Ai = &At[(size_t)i * M];
                T* Aj = &At[(size_t)j * M];
                double a = W[i], p = 0, b = W[j];
                for (int k = 0; k < M; ++k) p += (double)Ai[k] * Aj[k];
                if (std::abs(p) <= eps * std::sqrt(a * b)) continue;
                p *= 2;
                const double beta = a - b, gamma = hypot(p, beta);
                T c, s;
                if (beta < 0) {
                    const double delta = (gamma - beta) * 0.5;
                    s = (T)std::sqrt(delta / gamma);
                    c = (T)(p / (gamma * s * 2));
                } else {
                    c = (T)std::sqrt((gamma + beta) / (gamma * 2));
                    s = (T)(p / (gamma * c * 2));
                }
                a = b = 0;
                for (int k = 0; k < M; ++k) {
                    const T t0 = c * Ai[k] + s * Aj[k];
                    const T t1 = -s * Ai[k] + c * Aj[k];
                    Ai[k] = t0; Aj[k] = t1;
                    a += (double)t0 * t0; b += (double)t1 * t1;
                }
                W[i] = a; W[j] = b;
                changed = true;
                T* Vi = &Vt[(size_t)i * N];
                T* Vj = &Vt[(size_t)j * N];
                for (int k = 0; k < N; ++k) {
                    const T t0 = c * Vi[k] + s * Vj[k];
                    const T t1 = -s * Vi[k] + c * Vj[k];
                    Vi[k] = t0; Vj[k] = t1;
                }
            }
        if (!changed) break;
    }
    for (int i = 0; i < N; ++i) {
        double sd = 0;
        for (int k = 0; k < M; ++k) { const T t = At[(size_t)i * M + k]; sd += (double)t * t; }
        W[i] = std::sqrt(sd);
    }
    for (int i = 0; i < N - 1; ++i) {
        int j = i;
        for (int k = i + 1; k < N; ++k)
            if (W[j] < W[k]) j = k;
        if (i != j) {
            std::swap(W[i], W[j]);
            for (int k = 0; k < M; ++k) std::swap(At[(size_t)i * M + k], At[(size_t)j * M + k]);
            for (int k = 0; k < N; ++k) std::swap(Vt[(size_t)i * N + k], Vt[(size_t)j * N + k]);
        }
    }
    for (int i = 0; i < N; ++i) {
        w[i] = (T)W[i];
        const T s = (T)(W[i] > minval ? 1 / W[i] : 0.);
        for (int k = 0; k < M; ++k) At[(size_t)i * M + k] *= s;
    }
    // At: N x M = U^T of the (possibly transposed) problem; Vt: N x N
    if (!at) {
        for (int r = 0; r < m; ++r) for (int c2 = 0; c2 < N; ++c2) u[(size_t)r * N + c2] = At[(size_t)c2 * M + r];
        for (int i = 0; i < N * N; ++i) vt[i] = Vt[i];
    } else {
        // the problem solved was A^T = U' W V'^T  ->  A = V' W U'^T: u = V' (m x m... here N == m), vt = U'^T (N x n)
        for (int r = 0; r < N; ++r) for (int c2 = 0; c2 < N; ++c2) u[(size_t)r * N + c2] = Vt[(size_t)c2 * N + r];
        for (int i = 0; i < N; ++i) for (int k = 0; k < M; ++k) vt[(size_t)i * M + k] = At[(size_t)i * M + k];
    }
}

}  // namespace oracle
