// TEST INFRASTRUCTURE ONLY - see planepost.h.
#include "planepost.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstring>
#include <limits>

namespace oracle {

PclRng::PclRng(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
}
uint32_t PclRng::next_u32() {
    if (idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            const uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
            mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11; y ^= (y << 7) & 0x9d2c5680u; y ^= (y << 15) & 0xefc60000u; y ^= y >> 18;
    return y;
}

namespace {

struct P3 { float x, y, z; };

// cloud.vertices[pix] of PlaneDetection::readDepthImage (double), narrowed to float like Frame::ComputePlanes does (:657-661)
P3 vertex(const uint16_t* depth, int w, const PlanePostParams& K, int pix) {
    const int i = pix / w, j = pix - i * w;
    const double z = (double)depth[pix] * (double)K.scale;
    const double x = ((double)j - (double)K.cx) * z / (double)K.fx, y = ((double)i - (double)K.cy) * z / (double)K.fy;
    return P3{(float)x, (float)y, (float)z};
}

// pcl::VoxelGrid::applyFilter, leaf 0.1, min_points_per_voxel 0.  Centroid: PCL adds the floats of a voxel in whatever order std::sort (unstable, keyed on the
// voxel index only) leaves them - an unspecified order, so no bit pattern of PCL's sum can be targeted.  Here the sum is ORDER-FREE: every coordinate is
// quantised to 2^-20 m (llrint of the exact product), summed in 64-bit integers, divided in double and narrowed to float (within 1e-6 m of any float
// summation order); the CUDA path accumulates the same integers with atomics.
void voxel_grid(const std::vector<P3>& in, std::vector<P3>& out) {
    out.clear();
    if (in.empty()) return;
    const float inv = 1.0f / 0.1f;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (const P3& p : in) {
        mn[0] = std::min(mn[0], p.x); mn[1] = std::min(mn[1], p.y); mn[2] = std::min(mn[2], p.z);
        mx[0] = std::max(mx[0], p.x); mx[1] = std::max(mx[1], p.y); mx[2] = std::max(mx[2], p.z);
    }
    int min_b[3], max_b[3], div_b[3];
    for (int k = 0; k < 3; ++k) { min_b[k] = (int)std::floor(mn[k] * inv); max_b[k] = (int)std::floor(mx[k] * inv); div_b[k] = max_b[k] - min_b[k] + 1; }
    const int mul[3] = {1, div_b[0], div_b[0] * div_b[1]};
    std::vector<std::pair<unsigned, int>> iv(in.size());
    for (size_t q = 0; q < in.size(); ++q) {
        const int i0 = (int)(std::floor(in[q].x * inv) - (float)min_b[0]), i1 = (int)(std::floor(in[q].y * inv) - (float)min_b[1]),
                  i2 = (int)(std::floor(in[q].z * inv) - (float)min_b[2]);
        iv[q] = {(unsigned)(i0 * mul[0] + i1 * mul[1] + i2 * mul[2]), (int)q};
    }
    std::stable_sort(iv.begin(), iv.end(), [](const std::pair<unsigned, int>& a, const std::pair<unsigned, int>& b) { return a.first < b.first; });
    for (size_t a = 0; a < iv.size();) {
        size_t b = a;
        long long sx = 0, sy = 0, sz = 0;
        while (b < iv.size() && iv[b].first == iv[a].first) {
            const P3& p = in[iv[b].second];
            sx += std::llrint((double)p.x * 1048576.0); sy += std::llrint((double)p.y * 1048576.0); sz += std::llrint((double)p.z * 1048576.0);
            ++b;
        }
        const double n = (double)(b - a) * 1048576.0;
        out.push_back(P3{(float)((double)sx / n), (float)((double)sy / n), (float)((double)sz / n)});
        a = b;
    }
}

float dot4(const float c[4], const P3& p) { return ((c[0] * p.x + c[1] * p.y) + c[2] * p.z) + c[3]; }      // Eigen's 4-float dot of (a, b, c, d) . (x, y, z, 1)

// pcl::eigen33 (smallest eigen pair of a symmetric 3x3, closed form) in float like PCL's Matrix3f instantiation
void pcl_eigen33(const float m_in[3][3], float& eigenvalue, float v[3]) {
    float scale = 0;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(m_in[i][j]));
    if (scale <= std::numeric_limits<float>::min()) scale = 1.0f;
    float m[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m[i][j] = m_in[i][j] / scale;
    float roots[3];
    auto roots2 = [&](float b, float c) {          // computeRoots2: x^2 - b x + c = 0 plus the root 0
        roots[0] = 0.f;
        float d = b * b - 4.0f * c;
        if (d < 0.0f) d = 0.0f;
        const float sd = std::sqrt(d);
        roots[2] = 0.5f * (b + sd);
        roots[1] = 0.5f * (b - sd);
    };
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] - m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] - m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (std::fabs(c0) < std::numeric_limits<float>::epsilon()) roots2(c2, c1);
    else {
        const float s_inv3 = (float)(1.0 / 3.0), s_sqrt3 = std::sqrt(3.0f);
        const float c2_over_3 = c2 * s_inv3;
        float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
        if (a_over_3 > 0.0f) a_over_3 = 0.0f;
        const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
        float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
        if (q > 0.0f) q = 0.0f;
        const float rho = std::sqrt(-a_over_3);
        const float theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
        const float cos_theta = std::cos(theta), sin_theta = std::sin(theta);
        roots[0] = c2_over_3 + 2.0f * rho * cos_theta;
        roots[1] = c2_over_3 - rho * (cos_theta + s_sqrt3 * sin_theta);
        roots[2] = c2_over_3 - rho * (cos_theta - s_sqrt3 * sin_theta);
        if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]);
        if (roots[1] >= roots[2]) { std::swap(roots[1], roots[2]); if (roots[0] >= roots[1]) std::swap(roots[0], roots[1]); }
        if (roots[0] <= 0) roots2(c2, c1);
    }
    eigenvalue = roots[0] * scale;
    for (int i = 0; i < 3; ++i) m[i][i] -= roots[0];
    auto cross = [](const float a[3], const float b[3], float o[3]) { o[0] = a[1] * b[2] - a[2] * b[1]; o[1] = a[2] * b[0] - a[0] * b[2]; o[2] = a[0] * b[1] - a[1] * b[0]; };
    float v1[3], v2[3], v3[3];
    cross(m[0], m[1], v1); cross(m[0], m[2], v2); cross(m[1], m[2], v3);
    const float l1 = v1[0] * v1[0] + v1[1] * v1[1] + v1[2] * v1[2], l2 = v2[0] * v2[0] + v2[1] * v2[1] + v2[2] * v2[2], l3 = v3[0] * v3[0] + v3[1] * v3[1] + v3[2] * v3[2];
    const float* best = (l1 >= l2 && l1 >= l3) ? v1 : (l2 >= l1 && l2 >= l3) ? v2 : v3;
    const float len = std::sqrt((l1 >= l2 && l1 >= l3) ? l1 : (l2 >= l1 && l2 >= l3) ? l2 : l3);
    for (int k = 0; k < 3; ++k) v[k] = best[k] / len;
}

// pcl::SACSegmentation::segment: RANSAC (max 50 iterations, probability 0.99) + optimizeModelCoefficients; returns false when no model was found
bool sac_plane(const std::vector<P3>& pts, double th, float coef[4], int& n_inliers, int& n_iter) {
    const int N = (int)pts.size();
    n_inliers = 0; n_iter = 0;
    if (N < 3) return false;
    PclRng rng;
    std::vector<int> shuffled(N);
    for (int i = 0; i < N; ++i) shuffled[i] = i;
    int best_count = -INT_MAX, iterations = 0;
    double k = 1.0;
    const double log_probability = std::log(1.0 - 0.99), one_over = 1.0 / (double)N;
    float best[4] = {0, 0, 0, 0};
    bool have = false;
    unsigned skipped = 0;
    const unsigned max_skip = 50 * 10;
    while (iterations < k && skipped < max_skip) {
        int s[3];
        bool good = false;
        for (unsigned it = 0; it < 1000 && !good; ++it) {            // getSamples: drawIndexSample until isSampleGood
            for (int i = 0; i < 3; ++i) std::swap(shuffled[i], shuffled[i + (rng.rnd() % (N - i))]);
            for (int i = 0; i < 3; ++i) s[i] = shuffled[i];
            const P3 &p0 = pts[s[0]], &p1 = pts[s[1]], &p2 = pts[s[2]];
            const float d0 = (p1.x - p0.x) / (p2.x - p0.x), d1 = (p1.y - p0.y) / (p2.y - p0.y), d2 = (p1.z - p0.z) / (p2.z - p0.z);
            good = (d0 != d1) || (d2 != d1);
        }
        if (!good) break;
        const P3 &p0 = pts[s[0]], &p1 = pts[s[1]], &p2 = pts[s[2]];
        const float a[3] = {p1.x - p0.x, p1.y - p0.y, p1.z - p0.z}, b[3] = {p2.x - p0.x, p2.y - p0.y, p2.z - p0.z};
        const float e0 = a[0] / b[0], e1 = a[1] / b[1], e2 = a[2] / b[2];
        if (e0 == e1 && e2 == e1) { ++skipped; continue; }
        float c[4] = {a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0], 0.f};
        const float nrm = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
        for (int q = 0; q < 3; ++q) c[q] /= nrm;
        c[3] = -1 * (c[0] * p0.x + c[1] * p0.y + c[2] * p0.z);
        int cnt = 0;
        for (const P3& p : pts) if (std::fabs((double)dot4(c, p)) < th) ++cnt;
        if (cnt > best_count) {
            best_count = cnt; std::memcpy(best, c, sizeof best); have = true;
            const double w = (double)best_count * one_over;
            double p_no = 1.0 - std::pow(w, 3.0);
            p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
            p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
            k = log_probability / std::log(p_no);
        }
        ++iterations;
        if (iterations > 50) break;
    }
    n_iter = iterations;
    if (!have) return false;
    std::vector<int> inl;
    for (int i = 0; i < N; ++i) if (std::fabs((double)dot4(best, pts[i])) < th) inl.push_back(i);
    if (inl.empty()) return false;
    float opt[4];
    std::memcpy(opt, best, sizeof opt);
    if (inl.size() > 3) {          // optimizeModelCoefficients: single-pass float moments (computeMeanAndCovarianceMatrix) + eigen33
        float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        for (int i : inl) {
            const P3& p = pts[i];
            acc[0] += p.x * p.x; acc[1] += p.x * p.y; acc[2] += p.x * p.z; acc[3] += p.y * p.y; acc[4] += p.y * p.z; acc[5] += p.z * p.z; acc[6] += p.x; acc[7] += p.y; acc[8] += p.z;
        }
        for (float& v : acc) v /= (float)inl.size();
        float cov[3][3];
        cov[0][0] = acc[0] - acc[6] * acc[6]; cov[0][1] = acc[1] - acc[6] * acc[7]; cov[0][2] = acc[2] - acc[6] * acc[8];
        cov[1][1] = acc[3] - acc[7] * acc[7]; cov[1][2] = acc[4] - acc[7] * acc[8]; cov[2][2] = acc[5] - acc[8] * acc[8];
        cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
        float ev, vec[3];
        pcl_eigen33(cov, ev, vec);
        opt[0] = vec[0]; opt[1] = vec[1]; opt[2] = vec[2];
        opt[3] = -1 * (opt[0] * acc[6] + opt[1] * acc[7] + opt[2] * acc[8]);
    }
    int cnt = 0;
    for (int i = 0; i < N; ++i) if (std::fabs((double)dot4(opt, pts[i])) < th) ++cnt;       // selectWithinDistance with the refined model
    if (cnt == 0) return false;
    n_inliers = cnt;
    std::memcpy(coef, opt, sizeof opt);
    return true;
}

}  // namespace

void compute_planes_post(const uint16_t* depth, int w, int /*h*/, const PlanePostParams& prm, const PeacResult& peac, std::vector<PostPlane>& out) {
    out.clear();
    for (size_t i = 0; i < peac.planes.size(); ++i) {
        std::vector<P3> cloud;
        cloud.reserve(peac.membership[i].size());
        for (int pix : peac.membership[i]) cloud.push_back(vertex(depth, w, prm, pix));
        const PeacPlane& e = peac.planes[i];
        const float d = (float)-(e.normal[0] * e.center[0] + e.normal[1] * e.center[1] + e.normal[2] * e.center[2]);
        float coef[4] = {(float)e.normal[0], (float)e.normal[1], (float)e.normal[2], d};
        std::vector<P3> coarse;
        voxel_grid(cloud, coarse);
        bool valid = true;                                           // MaxPointDistanceFromPlane: every voxel centroid within the threshold
        for (const P3& p : coarse) if ((double)std::abs(dot4(coef, p)) > prm.dist_th) { valid = false; break; }
        if (!valid) continue;
        float nc[4];
        int n_inl = 0, n_it = 0;
        if (!sac_plane(coarse, prm.dist_th, nc, n_inl, n_it)) continue;
        const float old_d = coef[3], new_d = nc[3];
        std::memcpy(coef, nc, sizeof nc);
        if ((new_d < 0 && old_d > 0) || (new_d > 0 && old_d < 0)) for (float& v : coef) v = -v;
        PostPlane P;
        P.src = (int)i; std::memcpy(P.coef, coef, sizeof coef); P.n_inliers = n_inl; P.n_iterations = n_it;
        for (const P3& p : coarse) { P.points.push_back(p.x); P.points.push_back(p.y); P.points.push_back(p.z); }
        out.push_back(std::move(P));
    }
}

void map_plane_update(const std::vector<std::vector<float>>& clouds, const std::vector<const double*>& T, std::vector<float>& out_points) {
    std::vector<P3> combined, coarse;
    for (size_t c = 0; c < clouds.size(); ++c) {
        const double* t = T[c];
        for (size_t i = 0; i + 2 < clouds[c].size(); i += 3) {
            const double x = clouds[c][i], y = clouds[c][i + 1], z = clouds[c][i + 2];
            P3 p;                       // pcl::transformPointCloud<PointT, double>: rows of the matrix times (x, y, z, 1) in double, stored as float
            p.x = (float)(((t[0] * x + t[1] * y) + t[2] * z) + t[3]);
            p.y = (float)(((t[4] * x + t[5] * y) + t[6] * z) + t[7]);
            p.z = (float)(((t[8] * x + t[9] * y) + t[10] * z) + t[11]);
            combined.push_back(p);
        }
    }
    voxel_grid(combined, coarse);
    out_points.clear();
    for (const P3& p : coarse) { out_points.push_back(p.x); out_points.push_back(p.y); out_points.push_back(p.z); }
}

void surface_normals(const uint16_t* depth, int w, int h, const PlanePostParams& prm, std::vector<SurfaceNormal>& out) {
    out.clear();
    const int W = (int)std::ceil(w / 3.0), H = (int)std::ceil(h / 3.0);
    std::vector<P3> pts((size_t)W * H);
    for (int m = 0, r = 0; m < h; m += 3, ++r)
        for (int n = 0, c = 0; n < w; n += 3, ++c) {
            const float d = (float)depth[(size_t)m * w + n] * prm.scale;               // imDepth (CV_32F) = raw * mDepthMapFactor
            P3 p;
            p.z = d; p.x = ((float)n - prm.cx) * p.z / prm.fx; p.y = ((float)m - prm.cy) * p.z / prm.fy;
            pts[(size_t)r * W + c] = p;
        }
    const size_t NP = pts.size();
    // depth-change map -> chamfer distance map (IntegralImageNormalEstimation::computeFeature)
    std::vector<unsigned char> change(NP, 255);
    const float factor = 0.05f;
    for (int ri = 0; ri < H - 1; ++ri)
        for (int ci = 0; ci < W - 1; ++ci) {
            const size_t idx = (size_t)ri * W + ci;
            const float dep = pts[idx].z, depR = pts[idx + 1].z, depD = pts[idx + W].z;
            const float tol = factor * (std::fabs(dep) + 1.0f) * 2.0f;
            if (std::fabs(dep - depR) > tol || !std::isfinite(dep) || !std::isfinite(depR)) { change[idx] = 0; change[idx + 1] = 0; }
            if (std::fabs(dep - depD) > tol || !std::isfinite(dep) || !std::isfinite(depD)) { change[idx] = 0; change[idx + W] = 0; }
        }
    std::vector<float> dist(NP + 2, 0.f);          // flat array with one guard element on either side: PCL reads previous_row[ci + 1] / next_row[ci - 1] across row ends
    float* dm = dist.data() + 1;
    for (size_t i = 0; i < NP; ++i) dm[i] = change[i] == 0 ? 0.0f : (float)(W + H);
    dist[0] = dist[NP + 1] = (float)(W + H);
    for (int ri = 1; ri < H; ++ri) {
        float* prev = dm + (size_t)(ri - 1) * W; float* cur = dm + (size_t)ri * W;
        for (int ci = 1; ci < W; ++ci) {
            const float upLeft = prev[ci - 1] + 1.4f, up = prev[ci] + 1.0f, upRight = prev[ci + 1] + 1.4f, left = cur[ci - 1] + 1.0f, center = cur[ci];
            const float mv = std::min(std::min(upLeft, up), std::min(left, upRight));
            if (mv < center) cur[ci] = mv;
        }
    }
    for (int ri = H - 2; ri >= 0; --ri) {
        float* next = dm + (size_t)(ri + 1) * W; float* cur = dm + (size_t)ri * W;
        for (int ci = W - 2; ci >= 0; --ci) {
            const float lowerLeft = next[ci - 1] + 1.4f, lower = next[ci] + 1.0f, lowerRight = next[ci + 1] + 1.4f, right = cur[ci + 1] + 1.0f, center = cur[ci];
            const float mv = std::min(std::min(lowerLeft, lower), std::min(right, lowerRight));
            if (mv < center) cur[ci] = mv;
        }
    }
    // 3-D gradients (central differences) and their integral images in double (IntegralImage2D<float, 3>)
    std::vector<float> dxm(NP * 3, 0.f), dym(NP * 3, 0.f);
    for (int ri = 1; ri < H - 1; ++ri)
        for (int ci = 1; ci < W - 1; ++ci) {
            const size_t idx = (size_t)ri * W + ci;
            const P3 &rg = pts[idx + 1], &lf = pts[idx - 1], &dn = pts[idx + W], &up = pts[idx - W];
            dxm[idx * 3] = rg.x - lf.x; dxm[idx * 3 + 1] = rg.y - lf.y; dxm[idx * 3 + 2] = rg.z - lf.z;
            dym[idx * 3] = dn.x - up.x; dym[idx * 3 + 1] = dn.y - up.y; dym[idx * 3 + 2] = dn.z - up.z;
        }
    auto integral = [&](const std::vector<float>& src, std::vector<double>& I) {
        I.assign((size_t)(W + 1) * (H + 1) * 3, 0.0);
        for (int r = 0; r < H; ++r)
            for (int c = 0; c < W; ++c)
                for (int k = 0; k < 3; ++k) {
                    double v = I[((size_t)r * (W + 1) + c + 1) * 3 + k] + I[((size_t)(r + 1) * (W + 1) + c) * 3 + k] - I[((size_t)r * (W + 1) + c) * 3 + k];
                    const float e = src[((size_t)r * W + c) * 3 + k];
                    if (std::isfinite(src[((size_t)r * W + c) * 3]) && std::isfinite(src[((size_t)r * W + c) * 3 + 1]) && std::isfinite(src[((size_t)r * W + c) * 3 + 2])) v += (double)e;
                    I[((size_t)(r + 1) * (W + 1) + c + 1) * 3 + k] = v;
                }
    };
    std::vector<double> IX, IY;
    integral(dxm, IX); integral(dym, IY);
    auto rect_sum = [&](const std::vector<double>& I, int sx, int sy, int rw, int rh, double o[3]) {
        const size_t ul = (size_t)sy * (W + 1) + sx, ur = ul + rw, ll = (size_t)(sy + rh) * (W + 1) + sx, lr = ll + rw;
        for (int k = 0; k < 3; ++k) o[k] = I[lr * 3 + k] + I[ul * 3 + k] - I[ur * 3 + k] - I[ll * 3 + k];
    };
    const float nan = std::numeric_limits<float>::quiet_NaN();
    std::vector<float> nrm(NP * 3, nan);
    const int border = 10;
    const int bottom = H > border ? H - border : 0, right = W > border ? W - border : 0;
    for (int ri = border; ri < bottom; ++ri)
        for (int ci = border; ci < right; ++ci) {
            const size_t idx = (size_t)ri * W + ci;
            if (!std::isfinite(pts[idx].z)) continue;
            const float smoothing = std::min(dm[idx], 10.0f);
            if (!(smoothing > 2.0f)) continue;
            const int rw = (int)smoothing, rw2 = rw / 2;
            double gx[3], gy[3];
            rect_sum(IX, ci - rw2, ri - rw2, rw, rw, gx);
            rect_sum(IY, ci - rw2, ri - rw2, rw, rw, gy);
            double nv[3] = {gy[1] * gx[2] - gy[2] * gx[1], gy[2] * gx[0] - gy[0] * gx[2], gy[0] * gx[1] - gy[1] * gx[0]};
            const double len = nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2];
            if (len == 0.0) continue;
            const double s = std::sqrt(len);
            float nx = (float)(nv[0] / s), ny = (float)(nv[1] / s), nz = (float)(nv[2] / s);
            const P3& p = pts[idx];                                      // flipNormalTowardsViewpoint, view point (0, 0, 0)
            const float vx = 0.f - p.x, vy = 0.f - p.y, vz = 0.f - p.z;
            const float cos_theta = (vx * nx + vy * ny) + vz * nz;
            if (cos_theta < 0) { nx *= -1; ny *= -1; nz *= -1; }
            nrm[idx * 3] = nx; nrm[idx * 3 + 1] = ny; nrm[idx * 3 + 2] = nz;
        }
    for (int m = 0; m < H; ++m) {
        if (m % 2 == 0) continue;
        for (int n = 0; n < W; ++n) {
            if (n % 2 == 0) continue;
            const size_t idx = (size_t)m * W + n;
            SurfaceNormal s;
            s.normal[0] = nrm[idx * 3]; s.normal[1] = nrm[idx * 3 + 1]; s.normal[2] = nrm[idx * 3 + 2];
            s.cam[0] = pts[idx].x; s.cam[1] = pts[idx].y; s.cam[2] = pts[idx].z;
            s.frame_xy[0] = (float)(n * 3); s.frame_xy[1] = (float)(m * 3);
            out.push_back(s);
        }
    }
}

}  // namespace oracle
