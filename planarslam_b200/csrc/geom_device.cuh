// Double-precision SE(3) / quaternion / Plane3D device arithmetic shared by the pose-only and local-BA kernels; same
// formulas (and the same operation order) as Eigen / g2o use on the CPU:
//   SE3Quat  Thirdparty/g2o/g2o/types/se3quat.h (exp :227-260, operator* :98-104, normalizeRotation :286)
//   Plane3D  g2oAddition/Plane3D.h (normalize :175-180, rotation :76-82, azimuth / elevation :46-62)
#pragma once
#include <cuda_runtime.h>

namespace pslam {

// ---------------- small linear algebra (same formulas as Eigen / g2o use) ----------------
struct dV3 { double x, y, z; };
__device__ __forceinline__ dV3 dv(double x, double y, double z) { dV3 r; r.x = x; r.y = y; r.z = z; return r; }
__device__ __forceinline__ dV3 operator+(dV3 a, dV3 b) { return dv(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ dV3 operator*(double s, dV3 a) { return dv(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ double ddot(dV3 a, dV3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ dV3 dcross(dV3 a, dV3 b) { return dv(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
struct dM3 { double m[3][3]; };
__device__ __forceinline__ dV3 mmul(const dM3& A, dV3 v) {
    return dv(A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
              A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z);
}
__device__ __forceinline__ dM3 mmul(const dM3& A, const dM3& B) {
    dM3 C;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
struct dQuat { double x, y, z, w; };
struct dSE3 { dQuat q; dV3 t; };
__device__ __forceinline__ dQuat qnorm_pos(dQuat q) {
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    q.x /= n; q.y /= n; q.z /= n; q.w /= n;
    return q;
}
__device__ __forceinline__ dV3 qrot(dQuat q, dV3 v) {
    const dV3 u = dv(q.x, q.y, q.z);
    dV3 uv = dcross(u, v);
    uv = uv + uv;
    return v + q.w * uv + dcross(u, uv);
}
__device__ __forceinline__ dQuat qmul(dQuat a, dQuat b) {
    dQuat r;
    r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y; r.y = a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z;
    r.z = a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x; r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
    return r;
}
__device__ __forceinline__ dQuat quat_from_matrix(const dM3& R) {
    dQuat q;
    double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
    if (t > 0) {
        t = sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R.m[2][1] - R.m[1][2]) * t; q.y = (R.m[0][2] - R.m[2][0]) * t; q.z = (R.m[1][0] - R.m[0][1]) * t;
    } else {
        int i = 0;
        if (R.m[1][1] > R.m[0][0]) i = 1;
        if (R.m[2][2] > R.m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
        double c[3];
        c[i] = 0.5 * t;
        t = 0.5 / t;
        q.w = (R.m[k][j] - R.m[j][k]) * t;
        c[j] = (R.m[j][i] + R.m[i][j]) * t;
        c[k] = (R.m[k][i] + R.m[i][k]) * t;
        q.x = c[0]; q.y = c[1]; q.z = c[2];
    }
    return q;
}
__device__ __forceinline__ dM3 quat_to_matrix(dQuat q) {
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    dM3 R;
    R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz; R.m[0][2] = txz + twy;
    R.m[1][0] = txy + twz; R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
    R.m[2][0] = txz - twy; R.m[2][1] = tyz + twx; R.m[2][2] = 1 - (txx + tyy);
    return R;
}
static __device__ __noinline__ dSE3 se3_mul(const dSE3& a, const dSE3& b) {
    dSE3 r;
    r.t = a.t + qrot(a.q, b.t);
    r.q = qnorm_pos(qmul(a.q, b.q));
    return r;
}
static __device__ __noinline__ dSE3 se3_exp(const double u[6]) {
    const dV3 om = dv(u[0], u[1], u[2]), up = dv(u[3], u[4], u[5]);
    const double theta = sqrt(ddot(om, om));
    dM3 O;
    O.m[0][0] = 0; O.m[0][1] = -om.z; O.m[0][2] = om.y; O.m[1][0] = om.z; O.m[1][1] = 0; O.m[1][2] = -om.x;
    O.m[2][0] = -om.y; O.m[2][1] = om.x; O.m[2][2] = 0;
    const dM3 O2 = mmul(O, O);
    dM3 R, V;
    double a = 1.0, b = 1.0, c = 1.0;
    const bool small = theta < 0.00001;
    if (!small) { a = sin(theta) / theta; b = (1 - cos(theta)) / (theta * theta); c = (theta - sin(theta)) / pow(theta, 3.0); }
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const double id = (i == j) ? 1.0 : 0.0;
            R.m[i][j] = id + a * O.m[i][j] + b * O2.m[i][j];
            V.m[i][j] = small ? R.m[i][j] : id + b * O.m[i][j] + c * O2.m[i][j];
        }
    dSE3 T;
    T.q = qnorm_pos(quat_from_matrix(R));
    T.t = mmul(V, up);
    return T;
}

// ---------------- Plane3D ----------------
__device__ __forceinline__ void plane_normalize(double p[4]) {
    const double s = 1. / sqrt(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
#pragma unroll
    for (int i = 0; i < 4; ++i) p[i] = p[i] * s;
    if (p[3] < 0.0) {
#pragma unroll
        for (int i = 0; i < 4; ++i) p[i] = -p[i];
    }
}
__device__ __forceinline__ double azimuth(dV3 v) { return atan2(v.y, v.x); }
__device__ __forceinline__ double elevation(dV3 v) { return atan2(v.z, sqrt(v.x * v.x + v.y * v.y)); }
__device__ __forceinline__ dM3 plane_rotation_T(dV3 v) {        // transpose of Rz(azimuth) * Ry(-elevation)
    const double az = azimuth(v), el = -elevation(v);
    const double ca = cos(az), sa = sin(az), ce = cos(el), se = sin(el);
    dM3 Rz, Ry;
    Rz.m[0][0] = ca; Rz.m[0][1] = -sa; Rz.m[0][2] = 0; Rz.m[1][0] = sa; Rz.m[1][1] = ca; Rz.m[1][2] = 0; Rz.m[2][0] = 0; Rz.m[2][1] = 0; Rz.m[2][2] = 1;
    Ry.m[0][0] = ce; Ry.m[0][1] = 0; Ry.m[0][2] = se; Ry.m[1][0] = 0; Ry.m[1][1] = 1; Ry.m[1][2] = 0; Ry.m[2][0] = -se; Ry.m[2][1] = 0; Ry.m[2][2] = ce;
    const dM3 R = mmul(Rz, Ry);
    dM3 T;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) T.m[i][j] = R.m[j][i];
    return T;
}

}  // namespace pslam
