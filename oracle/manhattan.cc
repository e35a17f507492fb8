// TEST INFRASTRUCTURE ONLY — see manhattan.h
#include "manhattan.h"

#include <cmath>
#include <cstring>

#include "cvsvd.h"

namespace oracle {
namespace {

struct Mat3f {
    float v[9];
    float& at(int r, int c) { return v[3 * r + c]; }
    float at(int r, int c) const { return v[3 * r + c]; }
};

// the permuted, transposed rotation the reference passes as R_mc_new: row i = column c_i of R_cm
Mat3f axis_rotation(const Mat3f& R_cm, int a) {
    const int c[3] = {(a + 3) % 3, (a + 4) % 3, (a + 5) % 3};
    Mat3f R_mc, T;
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) R_mc.at(r, k) = R_cm.at(r, c[k]);
    for (int r = 0; r < 3; ++r) for (int k = 0; k < 3; ++k) T.at(r, k) = R_mc.at(k, r);
    return T;
}

inline void rotate_normal(const Mat3f& R, const float* nrm, float* out) {          // float * float, float sums
    for (int r = 0; r < 3; ++r) out[r] = R.at(r, 0) * nrm[0] + R.at(r, 1) * nrm[1] + R.at(r, 2) * nrm[2];
}
inline void rotate_dir(const Mat3f& R, const double* d, float* out) {              // float * double, double sums, stored to float
    for (int r = 0; r < 3; ++r) out[r] = (float)(R.at(r, 0) * d[0] + R.at(r, 1) * d[1] + R.at(r, 2) * d[2]);
}
inline double cone_lambda(const float* n) { return std::sqrt(n[0] * n[0] + n[1] * n[1]); }       // std::sqrt(float)

struct AxisSets { std::vector<int> normals, lines; };

}  // namespace

void track_manhattan_frame(const float* R_last, const float* normals, int n, const double* dirs, int m, ManhattanResult& res, uint8_t* normal_mask,
                           uint8_t* dir_mask) {
    Mat3f R;                                       // R_cm and R_cm_update share their data until the SVD
    std::memcpy(R.v, R_last, sizeof(R.v));
    std::memset(&res, 0, sizeof(res));
    if (normal_mask) std::memset(normal_mask, 0, (size_t)n);
    if (dir_mask) std::memset(dir_mask, 0, (size_t)m);
    AxisSets sets[3];
    for (int a = 1; a < 4; ++a) {                  // ProjectSN2Conic
        const Mat3f T = axis_rotation(R, a);
        for (int i = 0; i < n; ++i) {
            float q[3];
            rotate_normal(T, normals + 3 * i, q);
            if (cone_lambda(q) < std::sin(0.2018)) sets[a - 1].normals.push_back(i);
        }
        for (int i = 0; i < m; ++i) {
            float q[3];
            rotate_dir(T, dirs + 3 * i, q);
            if (cone_lambda(q) < std::sin(0.1018)) sets[a - 1].lines.push_back(i);
        }
        res.n_cone[a - 1] = (int)sets[a - 1].normals.size();
    }
    int minNumOfSN = n / 20;
    {
        int a = res.n_cone[0], b = res.n_cone[1], c = res.n_cone[2], t;
        if (a > b) t = a, a = b, b = t;
        if (b > c) t = b, b = c, c = t;
        if (a > b) t = a, a = b, b = t;
        if (b < minNumOfSN) minNumOfSN = (b + a) / 2;
    }
    res.min_num = minNumOfSN;
    int numDirectionFound = 0;
    for (int a = 1; a < 4; ++a) {                  // ProjectSN2MF
        const Mat3f T = axis_rotation(R, a);       // reads the columns already replaced for the earlier axes
        double nom_x = 0, nom_y = 0, den = 0;
        int count = 0;
        auto consider = [&](const float* q) -> bool {
            const double lambda = cone_lambda(q);
            if (!(lambda < std::sin(0.2518))) return false;
            const double tan_alfa = lambda / std::abs(q[2]);
            const double alfa = std::asin(lambda);
            const double mx = alfa / tan_alfa * q[0] / q[2], my = alfa / tan_alfa * q[1] / q[2];
            if (!std::isnan(mx) && !std::isnan(my)) {
                // MeanShift (:1139-1157): k = exp(-20 |m|^2)
                const double nr = std::sqrt(mx * mx + my * my);
                const double k = std::exp(-20 * nr * nr);
                nom_x += k * mx; nom_y += k * my; den += k;
                ++count;
            }
            return true;
        };
        for (int i : sets[a - 1].normals) {
            float q[3];
            rotate_normal(T, normals + 3 * i, q);
            if (consider(q) && normal_mask) normal_mask[i] |= (uint8_t)(1 << (a - 1));
        }
        for (int i : sets[a - 1].lines) {
            float q[3];
            rotate_dir(T, dirs + 3 * i, q);
            if (consider(q) && dir_mask) dir_mask[i] |= (uint8_t)(1 << (a - 1));
        }
        res.n_selected[a - 1] = count;
        if (count > minNumOfSN) {
            const double sx = nom_x / den, sy = nom_y / den;
            const float density = (float)(den / count);
            const float alfa = (float)std::sqrt(sx * sx + sy * sy);
            const float ma_x = (float)(std::tan(alfa) / alfa * sx);          // std::tan(float) / float -> float, * double -> double, -> float
            const float ma_y = (float)(std::tan(alfa) / alfa * sy);
            const float t1[3] = {ma_x, ma_y, 1.0f};
            // R_cm_Rec = R_mc_new^T * temp1 (float gemm, double accumulators), then / norm (scaled by (float)(1 / norm))
            float rec[3];
            for (int r = 0; r < 3; ++r) {
                double s = 0;
                for (int k = 0; k < 3; ++k) s += (double)T.at(k, r) * t1[k];
                rec[r] = (float)s;
            }
            double nn = 0;
            for (int r = 0; r < 3; ++r) nn += (double)rec[r] * rec[r];
            const float inv = (float)(1.0 / std::sqrt(nn));
            for (int r = 0; r < 3; ++r) rec[r] = rec[r] * inv;
            const double sum = (double)rec[0] + (double)rec[1] + (double)rec[2];
            if (sum != 0) {
                ++numDirectionFound;
                res.found[a - 1] = 1;
                res.density[a - 1] = density;
                for (int r = 0; r < 3; ++r) R.at(r, a - 1) = rec[r];
            }
        }
    }
    if (numDirectionFound < 2) {                   // "R_cm_update = R_cm": same buffer; returned as is
        std::memcpy(res.R, R.v, sizeof(R.v));
        return;
    }
    if (numDirectionFound == 2) {
        auto col = [&](int c, float* v) { for (int r = 0; r < 3; ++r) v[r] = R.at(r, c); };
        auto cross = [](const float* a, const float* b, float* o) {
            o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0];
        };
        auto det = [&]() {
            return (double)(R.at(0, 0) * (R.at(1, 1) * R.at(2, 2) - R.at(1, 2) * R.at(2, 1)) - R.at(0, 1) * (R.at(1, 0) * R.at(2, 2) - R.at(1, 2) * R.at(2, 0)) +
                            R.at(0, 2) * (R.at(1, 0) * R.at(2, 1) - R.at(1, 1) * R.at(2, 0)));
        };
        float u[3], w[3], x[3];
        int target;
        if (res.found[0] && res.found[1]) { col(0, u); col(1, w); cross(u, w, x); target = 2; }
        else if (res.found[1] && res.found[2]) { col(1, u); col(2, w); cross(w, u, x); target = 0; }       // v1 = v3.cross(v2)
        else { col(0, u); col(2, w); cross(u, w, x); target = 1; }                                         // v2 = v1.cross(v3)
        for (int r = 0; r < 3; ++r) R.at(r, target) = x[r];
        if (std::abs(det() + 1) < 0.5)
            for (int r = 0; r < 3; ++r) R.at(r, target) = -x[r];
    }
    float W[3], U[9], VT[9];
    cv_svd<float>(R.v, 3, 3, W, U, VT);
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += (double)U[3 * r + k] * VT[3 * k + c];
            res.R[3 * r + c] = (float)s;
        }
    res.svd_applied = 1;
}

}  // namespace oracle
