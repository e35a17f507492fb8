// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Shared double-precision SE3 / quaternion / Plane3D arithmetic used by the pose-only and local-BA restatements:
//   SE3Quat      Thirdparty/g2o/g2o/types/se3quat.h (exp :227-260, operator* :98-104, map :217, normalizeRotation :286)
//   Plane3D      g2oAddition/Plane3D.h (normalize :175-180, operator* :186-199, rotation :76-82, oplus :84-97,
//                ominus :127-134, ominus_ver :136-153, ominus_par :155-173) ; Converter::toPlane3D src/Converter.cc:171-180
// Eigen's quaternion<->matrix conversions and AngleAxis products are restated from Eigen's published algorithms.
#pragma once
#include <cmath>

namespace oracle {
namespace gm {

struct V3 { double x, y, z; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline double norm(V3 a) { return std::sqrt(dot(a, a)); }

struct M3 { double m[3][3]; };
inline V3 mul(const M3& A, V3 v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline M3 mul(const M3& A, const M3& B) {
    M3 C;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) C.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return C;
}
inline M3 transpose(const M3& A) { M3 T; for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.m[i][j] = A.m[j][i]; return T; }

struct Quat { double x, y, z, w; };
inline Quat qmul(Quat a, Quat b) {
    return {a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y, a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z,
            a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x, a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z};
}
inline Quat qnormalized_pos(Quat q) {   // SE3Quat::normalizeRotation
    if (q.w < 0) { q.x = -q.x; q.y = -q.y; q.z = -q.z; q.w = -q.w; }
    const double n = std::sqrt(q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w);
    return {q.x / n, q.y / n, q.z / n, q.w / n};
}
inline V3 qrot(Quat q, V3 v) {          // Eigen QuaternionBase::_transformVector
    const V3 u{q.x, q.y, q.z};
    V3 uv = cross(u, v);
    uv = uv + uv;
    return v + q.w * uv + cross(u, uv);
}
inline Quat quat_from_matrix(const M3& R) {    // Eigen quaternionbase_assign_impl<.., 3, 3>
    Quat q;
    double t = R.m[0][0] + R.m[1][1] + R.m[2][2];
    if (t > 0) {
        t = std::sqrt(t + 1.0);
        q.w = 0.5 * t;
        t = 0.5 / t;
        q.x = (R.m[2][1] - R.m[1][2]) * t; q.y = (R.m[0][2] - R.m[2][0]) * t; q.z = (R.m[1][0] - R.m[0][1]) * t;
    } else {
        int i = 0;
        if (R.m[1][1] > R.m[0][0]) i = 1;
        if (R.m[2][2] > R.m[i][i]) i = 2;
        const int j = (i + 1) % 3, k = (j + 1) % 3;
        t = std::sqrt(R.m[i][i] - R.m[j][j] - R.m[k][k] + 1.0);
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
inline M3 quat_to_matrix(Quat q) {             // Eigen QuaternionBase::toRotationMatrix
    const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
    const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x;
    const double tyy = ty * q.y, tyz = tz * q.y, tzz = tz * q.z;
    M3 R;
    R.m[0][0] = 1 - (tyy + tzz); R.m[0][1] = txy - twz; R.m[0][2] = txz + twy;
    R.m[1][0] = txy + twz; R.m[1][1] = 1 - (txx + tzz); R.m[1][2] = tyz - twx;
    R.m[2][0] = txz - twy; R.m[2][1] = tyz + twx; R.m[2][2] = 1 - (txx + tyy);
    return R;
}

struct SE3 { Quat q; V3 t; };
inline V3 se3_map(const SE3& T, V3 p) { return qrot(T.q, p) + T.t; }
inline SE3 se3_from_Rt(const M3& R, V3 t) { return {qnormalized_pos(quat_from_matrix(R)), t}; }
inline SE3 se3_mul(const SE3& a, const SE3& b) {   // SE3Quat::operator*
    SE3 r;
    r.t = a.t + qrot(a.q, b.t);
    r.q = qnormalized_pos(qmul(a.q, b.q));
    return r;
}
inline SE3 se3_exp(const double u[6]) {            // SE3Quat::exp, update = [omega, upsilon]
    const V3 om{u[0], u[1], u[2]}, up{u[3], u[4], u[5]};
    const double theta = norm(om);
    M3 O = {{{0, -om.z, om.y}, {om.z, 0, -om.x}, {-om.y, om.x, 0}}};
    M3 O2 = mul(O, O), R, V;
    if (theta < 0.00001) {
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R.m[i][j] = (i == j ? 1.0 : 0.0) + O.m[i][j] + O2.m[i][j];
        V = R;
    } else {
        const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta), c = (theta - std::sin(theta)) / std::pow(theta, 3);
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) {
            R.m[i][j] = (i == j ? 1.0 : 0.0) + a * O.m[i][j] + b * O2.m[i][j];
            V.m[i][j] = (i == j ? 1.0 : 0.0) + b * O.m[i][j] + c * O2.m[i][j];
        }
    }
    return se3_from_Rt(R, mul(V, up));
}

// ---- Plane3D ----
struct Plane { double c[4]; };
inline void plane_normalize(Plane& p) {
    const double n = std::sqrt(p.c[0] * p.c[0] + p.c[1] * p.c[1] + p.c[2] * p.c[2]);
    const double s = 1. / n;
    for (int i = 0; i < 4; ++i) p.c[i] = p.c[i] * s;
    if (p.c[3] < 0.0) for (int i = 0; i < 4; ++i) p.c[i] = -p.c[i];
}
inline Plane plane_from_float4(const float* v) {   // Converter::toPlane3D
    Plane p{{v[0], v[1], v[2], v[3]}};
    if (v[3] < 0.0f) for (int i = 0; i < 4; ++i) p.c[i] = -p.c[i];
    plane_normalize(p);
    return p;
}
inline V3 pn(const Plane& p) { return {p.c[0], p.c[1], p.c[2]}; }
inline double azimuth(V3 v) { return std::atan2(v.y, v.x); }
inline double elevation(V3 v) { return std::atan2(v.z, std::sqrt(v.x * v.x + v.y * v.y)); }
inline M3 plane_rotation(V3 v) {                    // Rz(azimuth) * Ry(-elevation)
    const double a = azimuth(v), e = -elevation(v);
    const double ca = std::cos(a), sa = std::sin(a), ce = std::cos(e), se = std::sin(e);
    M3 Rz = {{{ca, -sa, 0}, {sa, ca, 0}, {0, 0, 1}}}, Ry = {{{ce, 0, se}, {0, 1, 0}, {-se, 0, ce}}};
    return mul(Rz, Ry);
}
inline Plane plane_transform(const SE3& T, const Plane& p) {   // operator*(Isometry3D, Plane3D)
    const M3 R = quat_to_matrix(T.q);
    const V3 n = mul(R, pn(p));
    Plane r{{n.x, n.y, n.z, p.c[3] - dot(T.t, n)}};
    if (r.c[3] < 0.0) for (int i = 0; i < 4; ++i) r.c[i] = -r.c[i];
    plane_normalize(r);
    return r;
}
inline void plane_ominus(const Plane& self, const Plane& meas, double e[3]) {
    const M3 R = transpose(plane_rotation(pn(self)));
    const V3 n = mul(R, pn(meas));
    e[0] = azimuth(n); e[1] = elevation(n); e[2] = (-self.c[3]) - (-meas.c[3]);
}
inline void plane_ominus_par(const Plane& self, const Plane& meas, double e[2]) {
    V3 nor = pn(self);
    if (dot(pn(meas), nor) < 0) nor = -1.0 * nor;
    const M3 R = transpose(plane_rotation(nor));
    const V3 n = mul(R, pn(meas));
    e[0] = azimuth(n); e[1] = elevation(n);
}
inline void plane_ominus_ver(const Plane& self, const Plane& meas, double e[2]) {
    const V3 v = cross(pn(self), pn(meas));
    const V3 ax = (1.0 / norm(v)) * v;
    // AngleAxis(pi/2, ax) * normal  (Rodrigues)
    const double ang = M_PI / 2, c = std::cos(ang), s = std::sin(ang);
    const V3 nrm = pn(self);
    const V3 b = c * nrm + s * cross(ax, nrm) + ((1 - c) * dot(ax, nrm)) * ax;
    const M3 R = transpose(plane_rotation(b));
    const V3 n = mul(R, pn(meas));
    e[0] = azimuth(n); e[1] = elevation(n);
}

}  // namespace gm
}  // namespace oracle
