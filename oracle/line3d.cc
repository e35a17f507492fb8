// TEST INFRASTRUCTURE ONLY — see line3d.h
#include "line3d.h"

#include <algorithm>
#include <cmath>

#include "cvsvd.h"

namespace oracle {

void GlibcRand::srand(uint32_t seed) {
    // glibc __srandom_r, TYPE_3 (deg 31, sep 3): state[0] = seed (0 -> 1), state[i] = 16807 * state[i-1] % 2147483647 (Schrage),
    // then 310 outputs are discarded
    if (seed == 0) seed = 1;
    state[0] = (int32_t)seed;
    int32_t word = (int32_t)seed;                // int32_t in glibc: seeds >= 2^31 start negative
    for (int i = 1; i < 31; ++i) {
        const long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        state[i] = word;
    }
    f = 3; b = 0;
    drawn = 0;
    for (int i = 0; i < 310; ++i) (void)rand();
    drawn = 0;
}

int32_t GlibcRand::rand() {
    const uint32_t val = (uint32_t)state[f] + (uint32_t)state[b];
    state[f] = (int32_t)val;
    const int32_t result = (int32_t)(val >> 1);
    if (++f >= 31) f = 0;
    if (++b >= 31) b = 0;
    ++drawn;
    return result;
}

namespace {

struct P3 { double x, y, z; };
inline P3 operator+(const P3& a, const P3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline P3 operator-(const P3& a, const P3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline P3 operator*(const P3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline double dot(const P3& a, const P3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline double norm(const P3& a) { return std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z); }

struct RandomPoint3d { P3 pos; double DU[9]; };

const double kEps = 1e-10;

double depth_std_dev(double d) { return 0.00273 * d * d + 0.00074 * d + -0.00058; }

void mat3_mul(const double* a, const double* b, double* c) {
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double s = 0;
            for (int k = 0; k < 3; ++k) s += a[3 * i + k] * b[3 * k + j];
            c[3 * i + j] = s;
        }
}

RandomPoint3d comp_pt3d_cov(const P3& pt, double f) {
    RandomPoint3d rp;
    rp.pos = pt;
    const double J0[9] = {pt.z / f, 0, pt.x / pt.z, 0, pt.z / f, pt.y / pt.z, 0, 0, 1};
    const double sd = depth_std_dev(pt.z);
    const double cg[9] = {1, 0, 0, 0, 1, 0, 0, 0, sd * sd};
    double J0t[9], t[9], cov[9];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) J0t[3 * i + j] = J0[3 * j + i];
    mat3_mul(J0, cg, t);
    mat3_mul(t, J0t, cov);
    double w[3], u[9], vt[9];
    cv_svd<double>(cov, 3, 3, w, u, vt);
    const double ws[3] = {std::sqrt(w[0]), std::sqrt(w[1]), std::sqrt(w[2])};
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) rp.DU[3 * r + c] = (1 / ws[r]) * u[3 * c + r];        // D * U^T
    return rp;
}

double mah_dist3d_pt_line(const RandomPoint3d& pt, const P3& q1, const P3& q2) {
    const double xa = q1.x, ya = q1.y, za = q1.z, xb = q2.x, yb = q2.y, zb = q2.z;
    const double c1 = pt.DU[0], c2 = pt.DU[1], c3 = pt.DU[2], c4 = pt.DU[3], c5 = pt.DU[4], c6 = pt.DU[5], c7 = pt.DU[6], c8 = pt.DU[7], c9 = pt.DU[8];
    const double x1 = pt.pos.x, x2 = pt.pos.y, x3 = pt.pos.z;
    const double term1 = ((c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za)) * (c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb)) -
                          (c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za)) * (c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb)));
    const double term2 = ((c1 * (x1 - xa) + c2 * (x2 - ya) + c3 * (x3 - za)) * (c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb)) -
                          (c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za)) * (c1 * (x1 - xb) + c2 * (x2 - yb) + c3 * (x3 - zb)));
    const double term3 = ((c4 * (x1 - xa) + c5 * (x2 - ya) + c6 * (x3 - za)) * (c7 * (x1 - xb) + c8 * (x2 - yb) + c9 * (x3 - zb)) -
                          (c7 * (x1 - xa) + c8 * (x2 - ya) + c9 * (x3 - za)) * (c4 * (x1 - xb) + c5 * (x2 - yb) + c6 * (x3 - zb)));
    const double term4 = (c1 * (x1 - xa) - c1 * (x1 - xb) + c2 * (x2 - ya) - c2 * (x2 - yb) + c3 * (x3 - za) - c3 * (x3 - zb));
    const double term5 = (c4 * (x1 - xa) - c4 * (x1 - xb) + c5 * (x2 - ya) - c5 * (x2 - yb) + c6 * (x3 - za) - c6 * (x3 - zb));
    const double term6 = (c7 * (x1 - xa) - c7 * (x1 - xb) + c8 * (x2 - ya) - c8 * (x2 - yb) + c9 * (x3 - za) - c9 * (x3 - zb));
    return std::sqrt((term1 * term1 + term2 * term2 + term3 * term3) / (term4 * term4 + term5 * term5 + term6 * term6));
}

P3 project_pt3d_2_ln3d(const P3& P, const P3& mid, const P3& drct) {
    const P3 A = mid, B = mid + drct, AB = B - A, AP = P - A;
    return A + AB * (dot(AB, AP) / dot(AB, AB));
}

bool verify3d_line(const std::vector<RandomPoint3d>& pts, const P3& A, const P3& B) {
    const int nCells = 10;
    int cells[10] = {0};
    const double ratio = 0.7;
    const int nPts = (int)pts.size();
    double minv = 100, maxv = -100;
    int idx1 = 0, idx2 = 0;
    for (int i = 0; i < nPts; ++i) {
        const double v = dot(pts[i].pos - A, B - A);
        if (v < minv) { minv = v; idx1 = i; }
        if (v > maxv) { maxv = v; idx2 = i; }
    }
    const P3 C = project_pt3d_2_ln3d(pts[idx1].pos, (A + B) * 0.5, B - A);
    const P3 D = project_pt3d_2_ln3d(pts[idx2].pos, (A + B) * 0.5, B - A);
    const double cd = norm(D - C);
    if (cd < kEps) return false;
    for (int i = 0; i < nPts; ++i) {
        const double lambda = std::abs(dot(pts[i].pos - C, D - C) / cd / cd);
        if (lambda >= 1) cells[nCells - 1] += 1;
        else cells[(unsigned)std::floor(lambda * 10)] += 1;
    }
    double sum = 0;
    for (int i = 0; i < nCells; ++i)
        if (cells[i] > 0) sum = sum + 1;
    return sum / nCells > ratio;
}

void compute_line3d_svd(const std::vector<RandomPoint3d>& pts, const std::vector<int>& idx, P3& mean, P3& drct) {
    const int n = (int)idx.size();
    mean = {0, 0, 0};
    for (int i = 0; i < n; ++i) mean = mean + pts[idx[i]].pos;
    mean = mean * (1.0 / n);
    std::vector<double> Pt((size_t)n * 3);           // P.t(): n x 3
    for (int i = 0; i < n; ++i) {
        Pt[3 * i] = pts[idx[i]].pos.x - mean.x; Pt[3 * i + 1] = pts[idx[i]].pos.y - mean.y; Pt[3 * i + 2] = pts[idx[i]].pos.z - mean.z;
    }
    const int k = std::min(n, 3);
    std::vector<double> w(k), u((size_t)n * k), vt((size_t)k * 3);
    cv_svd<double>(Pt.data(), n, 3, w.data(), u.data(), vt.data());
    drct = {vt[0], vt[1], vt[2]};
}

struct RandomLine3d { std::vector<int> pts; P3 A{0, 0, 0}, B{0, 0, 0}, director{0, 0, 0}; };

RandomLine3d extract3dline_mahdist(const std::vector<RandomPoint3d>& pts, GlibcRand& rng) {
    const int n = (int)pts.size();
    const int maxIterNo = std::min(10, int(n * (n - 1) * 0.5));
    const double distThresh = 1.5;
    std::vector<int> indexes(n);
    for (int i = 0; i < n; ++i) indexes[i] = i;
    std::vector<int> maxInlierSet;
    int bestA = 0, bestB = 0;
    for (int iter = 0; iter < maxIterNo; ++iter) {
        std::vector<int> inlierSet;
        {   // random_unique(begin, end, 2)
            size_t left = (size_t)n;
            int* begin = indexes.data();
            for (int q = 0; q < 2; ++q) {
                std::swap(*begin, *(begin + (size_t)rng.rand() % left));
                ++begin; --left;
            }
        }
        const RandomPoint3d& A = pts[indexes[0]];
        const RandomPoint3d& B = pts[indexes[1]];
        if (norm(B.pos - A.pos) < kEps) continue;
        for (int i = 0; i < n; ++i)
            if (mah_dist3d_pt_line(pts[i], A.pos, B.pos) < distThresh) inlierSet.push_back(i);
        if (inlierSet.size() > maxInlierSet.size()) {
            std::vector<RandomPoint3d> inlierPts(inlierSet.size());
            for (size_t ii = 0; ii < inlierSet.size(); ++ii) inlierPts[ii] = pts[inlierSet[ii]];
            if (verify3d_line(inlierPts, A.pos, B.pos)) { maxInlierSet = inlierSet; bestA = indexes[0]; bestB = indexes[1]; }
        }
        if (maxInlierSet.size() > n * 0.6) break;
    }
    RandomLine3d rl;
    if (maxInlierSet.size() >= 2) {
        P3 m = (pts[bestA].pos + pts[bestB].pos) * 0.5, d = pts[bestB].pos - pts[bestA].pos;
        while (true) {
            std::vector<int> tmpInlierSet;
            P3 tmp_m, tmp_d;
            compute_line3d_svd(pts, maxInlierSet, tmp_m, tmp_d);
            for (int i = 0; i < n; ++i)
                if (mah_dist3d_pt_line(pts[i], tmp_m, tmp_m + tmp_d) < distThresh) tmpInlierSet.push_back(i);
            if (tmpInlierSet.size() > maxInlierSet.size()) { maxInlierSet = tmpInlierSet; m = tmp_m; d = tmp_d; }
            else break;
        }
        double minv = 100, maxv = -100;
        int idx_end1 = 0, idx_end2 = 0;
        for (size_t i = 0; i < maxInlierSet.size(); ++i) {
            const double dproduct = dot(pts[maxInlierSet[i]].pos - m, d);
            if (dproduct < minv) { minv = dproduct; idx_end1 = (int)i; }
            if (dproduct > maxv) { maxv = dproduct; idx_end2 = (int)i; }
        }
        rl.A = pts[maxInlierSet[idx_end1]].pos;
        rl.B = pts[maxInlierSet[idx_end2]].pos;
    }
    const P3 ab = rl.A - rl.B;
    const double nn = std::sqrt(dot(ab, ab));
    rl.director = {ab.x / nn, ab.y / nn, ab.z / nn};
    rl.pts = maxInlierSet;
    return rl;
}

}  // namespace

void lines3d_frame(const KeyLine* kl, int n_lines, const float* depth, const Line3dCam& cam, GlibcRand& rng, Line3dResult* out) {
    for (int i = 0; i < n_lines; ++i) {
        Line3dResult& R = out[i];
        R = Line3dResult();
        const KeyLine& k = kl[i];
        // cv::norm(Point2f): the difference is taken in float, the norm in double
        const float ddx = k.startPointX - k.endPointX, ddy = k.startPointY - k.endPointY;
        const double len = std::sqrt((double)ddx * ddx + (double)ddy * ddy);
        const double numSmp = (double)std::min((int)len, 50);
        std::vector<P3> pts3d;
        if (numSmp >= 1)                   // numSmp == 0 divides 0 / 0 in the reference (undefined look-up); no line that short leaves the detector
            for (int j = 0; j <= numSmp; ++j) {
                // Point2f * double -> Point2f (product in double, rounded to float), Point2f + Point2f in float, then -> Point2d
                const double t1 = 1 - j / numSmp, t2 = j / numSmp;
                const float px = (float)(k.startPointX * t1) + (float)(k.endPointX * t2);
                const float py = (float)(k.startPointY * t1) + (float)(k.endPointY * t2);
                const double ptx = px, pty = py;
                if (ptx < 0 || pty < 0 || ptx >= cam.w || pty >= cam.h) continue;
                int row, col;
                if (std::floor(ptx) == ptx && std::floor(pty) == pty) { col = std::max(int(ptx - 1), 0); row = std::max(int(pty - 1), 0); }
                else { col = int(ptx); row = int(pty); }
                const float dv = depth[(size_t)row * cam.w + col];
                if (dv <= 0.01) continue;
                P3 p;
                p.z = dv;
                p.x = (col - cam.cx) * p.z * cam.invfx;          // (int - float) in float, then double products
                p.y = (row - cam.cy) * p.z * cam.invfy;
                pts3d.push_back(p);
            }
        R.n_points = (int)pts3d.size();
        if (pts3d.size() < 10.0) continue;
        std::vector<RandomPoint3d> rnd;
        rnd.reserve(pts3d.size());
        for (const P3& p : pts3d) rnd.push_back(comp_pt3d_cov(p, (double)cam.fx));
        const RandomLine3d tl = extract3dline_mahdist(rnd, rng);
        R.n_inliers = (int)tl.pts.size();
        for (int q : tl.pts) R.inliers |= (uint64_t)1 << q;
        for (int c = 0; c < 3; ++c) R.director[c] = (&tl.director.x)[c];
        const P3 ab = tl.A - tl.B;
        if (tl.pts.size() / len > 0.4 && norm(ab) > 0.02) {
            R.valid = 1;
            // Mat::at<float>(float row, float col): the KeyLine coordinates are truncated to int
            R.depth = std::min(depth[(size_t)(int)k.endPointY * cam.w + (int)k.endPointX], depth[(size_t)(int)k.startPointY * cam.w + (int)k.startPointX]);
            R.A[0] = tl.A.x; R.A[1] = tl.A.y; R.A[2] = tl.A.z; R.B[0] = tl.B.x; R.B[1] = tl.B.y; R.B[2] = tl.B.z;
        }
    }
}

}  // namespace oracle
