// TEST INFRASTRUCTURE ONLY — CPU oracle restating cv::LineSegmentDetector (OpenCV imgproc lsd.cpp), the KeyLine glue of
// opencv_contrib's LSDDetector and LineSegment::ExtractLineSegment (see lsd.h for the map and the pinning status).
#include "lsd.h"
#include <cstdio>
#include <cstdlib>
#include "detmath.h"

#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstring>

namespace oracle {
namespace {

const double kPi = 3.14159265358979323846;
const double kNotDef = -1024.0;
const double kDegToRads = kPi / 180;
const double k3_2Pi = (3 * kPi) / 2;
const double k2Pi = 2 * kPi;
const double kLn10 = 2.30258509299404568402;

// cv::GaussianBlur(u8, Size(7,7), 0.75): bit-exact 8.8 fixed-point path, taps from getGaussianKernel's fixed-point variant
const int kGauss7S075[7] = {0, 4, 56, 136, 56, 4, 0};

void gaussian_blur_7x7_s075_u8(const Img8& s, uint8_t* dst) {
    const int w = s.w, h = s.h;
    std::vector<uint16_t> tmp((size_t)w * h);
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            unsigned acc = 0;
            for (int k = -3; k <= 3; ++k) acc += kGauss7S075[k + 3] * s.at(y, reflect101(x + k, w));
            tmp[(size_t)y * w + x] = (uint16_t)acc;
        }
    for (int y = 0; y < h; ++y)
        for (int x = 0; x < w; ++x) {
            uint32_t acc = 0;
            for (int k = -3; k <= 3; ++k) acc += (uint32_t)kGauss7S075[k + 3] * tmp[(size_t)reflect101(y + k, h) * w + x];
            dst[(size_t)y * w + x] = (uint8_t)std::min<uint32_t>(255u, (acc + 32768u) >> 16);
        }
}

// cv::resize(u8, INTER_LINEAR_EXACT): 8.8 fixed-point coefficients, horizontal pass in 8.8, vertical in 16.16, round half up
// cv::resize(src, dst, Size(), fx, fy, INTER_LINEAR_EXACT): the destination size is cvRound(size * f) but the sampling step stays 1 / f
// (it is NOT recomputed as src / dst; the two differ whenever size * f is not an integer - checked against cv2 on 641 x 479)
void resize_linear_exact_u8(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, double inv_fx, double inv_fy) {
    auto coef = [](int dn, int sn, double scale, std::vector<int>& idx, std::vector<int>& a) {
        idx.resize(dn); a.resize(dn);
        for (int d = 0; d < dn; ++d) {
            const double f = (d + 0.5) * scale - 0.5;
            int i = (int)std::floor(f);
            double fr = f - i;
            if (i < 0) { i = 0; fr = 0; }
            if (i >= sn - 1) { i = sn - 1; fr = 0; }
            idx[d] = i; a[d] = (int)std::floor(fr * 256 + 0.5);
        }
    };
    std::vector<int> ix, ax, iy, ay;
    coef(dw, sw, inv_fx, ix, ax); coef(dh, sh, inv_fy, iy, ay);
    std::vector<uint32_t> hl((size_t)sh * dw);
    for (int y = 0; y < sh; ++y)
        for (int x = 0; x < dw; ++x) {
            const int x0 = ix[x], x1 = std::min(x0 + 1, sw - 1);
            hl[(size_t)y * dw + x] = (uint32_t)(256 - ax[x]) * src[(size_t)y * sw + x0] + (uint32_t)ax[x] * src[(size_t)y * sw + x1];
        }
    for (int y = 0; y < dh; ++y) {
        const int y0 = iy[y], y1 = std::min(y0 + 1, sh - 1);
        for (int x = 0; x < dw; ++x) {
            const uint32_t v = (uint32_t)(256 - ay[y]) * hl[(size_t)y0 * dw + x] + (uint32_t)ay[y] * hl[(size_t)y1 * dw + x];
            dst[(size_t)y * dw + x] = (uint8_t)((v + 32768u) >> 16);
        }
    }
}

struct RegionPoint { int x, y; double angle, modgrad; };
struct Rect { double x1, y1, x2, y2, width, x, y, theta, dx, dy, prec, p; };
struct NormPoint { int x, y, norm; };

inline double dist_sq(double x1, double y1, double x2, double y2) { return (x2 - x1) * (x2 - x1) + (y2 - y1) * (y2 - y1); }
inline double dist(double x1, double y1, double x2, double y2) { return std::sqrt(dist_sq(x1, y1, x2, y2)); }
inline double angle_diff_signed(double a, double b) {
    double diff = a - b;
    while (diff <= -kPi) diff += k2Pi;
    while (diff > kPi) diff -= k2Pi;
    return diff;
}
inline double angle_diff(double a, double b) { return std::fabs(angle_diff_signed(a, b)); }
inline bool double_equal(double a, double b) {
    if (a == b) return true;
    const double abs_diff = std::fabs(a - b), aa = std::fabs(a), bb = std::fabs(b);
    double abs_max = (aa > bb) ? aa : bb;
    if (abs_max < DBL_MIN) abs_max = DBL_MIN;
    return (abs_diff / abs_max) <= (100.0 * DBL_EPSILON);
}
inline double log_gamma_windschitl(double x) {
    return 0.918938533204673 + (x - 0.5) * std::log(x) - x + 0.5 * x * std::log(x * std::sinh(1 / x) + 1 / (810.0 * std::pow(x, 6.0)));
}
inline double log_gamma_lanczos(double x) {
    static const double q[7] = {75122.6331530, 80916.6278952, 36308.2951477, 8687.24529705, 1168.92649479, 83.8676043424, 2.50662827511};
    double a = (x + 0.5) * std::log(x + 5.5) - (x + 5.5);
    double b = 0;
    for (int n = 0; n < 7; ++n) {
        a -= std::log(x + double(n));
        b += q[n] * std::pow(x, double(n));
    }
    return a + std::log(b);
}
inline double log_gamma(double x) { return x > 15.0 ? log_gamma_windschitl(x) : log_gamma_lanczos(x); }

// Row spans (y, first column, last column; clamped to the W x H image, empty rows skipped) of cv2 4.13's rect_nfa enumeration -
// see the comment at Lsd::rect_count_cv4.
template <class F>
void rect_rows_cv4(double x1, double y1, double x2, double y2, double width, double dx, double dy, int W, int H, F&& emit) {
    const double half = 0.5 * width, dyhw = dy * half, dxhw = half * dx;
    const double cx[4] = {x1 - dyhw, x2 - dyhw, x2 + dyhw, x1 + dyhw};
    const double cy[4] = {y1 + dxhw, y2 + dxhw, y2 - dxhw, y1 - dxhw};
    int off = 0;
    for (int i = 1; i < 4; ++i)
        if (cy[i] < cy[off] || (cy[i] == cy[off] && cx[off] > cx[i])) off = i;
    double vx[4], vy[4];
    for (int q = 0; q < 4; ++q) { vx[q] = cx[(off + q) & 3]; vy[q] = cy[(off + q) & 3]; }
    const int y0 = (int)std::ceil(vy[0]), c1 = (int)std::ceil(vy[1]), c2 = (int)std::ceil(vy[2]), c3 = (int)std::ceil(vy[3]);
    const double s01 = c1 != y0 ? (vx[1] - vx[0]) / (vy[1] - vy[0]) : 0.0;
    const double s12 = c2 != c1 ? (vx[2] - vx[1]) / (vy[2] - vy[1]) : 0.0;
    const double s03 = c3 != y0 ? (vx[3] - vx[0]) / (vy[3] - vy[0]) : 0.0;
    const double s32 = c3 != c2 ? (vx[2] - vx[3]) / (vy[2] - vy[3]) : 0.0;
    for (int y = y0; y <= c2; ++y) {
        if (y < 0 || y >= H) continue;
        const double yd = (double)y;
        const double left = y > c1 ? (yd - vy[1]) * s12 + vx[1] : (yd - vy[0]) * s01 + vx[0];
        const double right = y >= c3 ? (yd - vy[3]) * s32 + vx[3] : (yd - vy[0]) * s03 + vx[0];
        int xa = (int)std::ceil(left), xb = (int)right;       // x86 conversions: out-of-range -> INT_MIN, like the reference binary
        if (xa < 0) xa = 0;
        if (xb > W - 1) xb = W - 1;
        if (xb >= xa) emit(y, xa, xb);
    }
}

struct Lsd {
    int W = 0, H = 0;
    std::vector<uint8_t> scaled;
    std::vector<double> angles, modgrad;
    std::vector<uint8_t> used;
    std::vector<NormPoint> ordered;
    double LOG_NT = 0;
    const double LOG_EPS = 0, DENSITY_TH = 0.7;

    bool is_aligned(int x, int y, double theta, double prec) const {
        if (x < 0 || y < 0 || x >= W || y >= H) return false;
        const double a = angles[(size_t)y * W + x];
        if (a == kNotDef) return false;
        double n_theta = theta - a;
        if (n_theta < 0) n_theta = -n_theta;
        if (n_theta > k3_2Pi) {
            n_theta -= k2Pi;
            if (n_theta < 0) n_theta = -n_theta;
        }
        return n_theta <= prec;
    }

    void ll_angle(double threshold, int n_bins) {
        angles.assign((size_t)W * H, kNotDef);
        modgrad.assign((size_t)W * H, 0.0);
        double max_grad = -1;
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) {
                const size_t addr = (size_t)y * W + x;
                const int DA = scaled[addr + W + 1] - scaled[addr];
                const int BC = scaled[addr + 1] - scaled[addr + W];
                const int gx = DA + BC, gy = DA - BC;
                const double norm = std::sqrt((gx * gx + gy * gy) / 4.0);
                modgrad[addr] = norm;
                if (norm <= threshold) angles[addr] = kNotDef;
                else {
                    angles[addr] = fast_atan2_deg(float(gx), float(-gy)) * kDegToRads;
                    if (norm > max_grad) max_grad = norm;
                }
            }
        const double bin_coef = (max_grad > 0) ? double(n_bins - 1) / max_grad : 0;
        ordered.clear();
        ordered.reserve((size_t)(W - 1) * (H - 1));
        for (int y = 0; y < H - 1; ++y)
            for (int x = 0; x < W - 1; ++x) ordered.push_back({x, y, int(modgrad[(size_t)y * W + x] * bin_coef)});
        // equal bins keep their row-major order (verified against cv2: the segment order of REFINE_NONE depends on it)
        std::stable_sort(ordered.begin(), ordered.end(), [](const NormPoint& a, const NormPoint& b) { return a.norm > b.norm; });
    }

    void region_grow(int sx, int sy, std::vector<RegionPoint>& reg, double& reg_angle, double prec) {
        reg.clear();
        reg_angle = angles[(size_t)sy * W + sx];
        reg.push_back({sx, sy, reg_angle, modgrad[(size_t)sy * W + sx]});
        double sn, cs;
        det_sincos(reg_angle, sn, cs);             // std::cos / std::sin in OpenCV; see detmath.h
        if (libm_sincos) { cs = std::cos(reg_angle); sn = std::sin(reg_angle); }
        float sumdx = float(cs);
        float sumdy = float(sn);
        used[(size_t)sy * W + sx] = 1;
        for (size_t i = 0; i < reg.size(); ++i) {
            const int px = reg[i].x, py = reg[i].y;
            const int xx_min = std::max(px - 1, 0), xx_max = std::min(px + 1, W - 1);
            const int yy_min = std::max(py - 1, 0), yy_max = std::min(py + 1, H - 1);
            for (int yy = yy_min; yy <= yy_max; ++yy)
                for (int xx = xx_min; xx <= xx_max; ++xx) {
                    uint8_t& is_used = used[(size_t)yy * W + xx];
                    if (is_used != 1 && is_aligned(xx, yy, reg_angle, prec)) {
                        const double angle = angles[(size_t)yy * W + xx];
                        is_used = 1;
                        reg.push_back({xx, yy, angle, modgrad[(size_t)yy * W + xx]});
                        sumdx += std::cos(float(angle));
                        sumdy += std::sin(float(angle));
                        reg_angle = fast_atan2_deg(sumdy, sumdx) * kDegToRads;
                    }
                }
        }
    }

    double get_theta(const std::vector<RegionPoint>& reg, double x, double y, double reg_angle, double prec) const {
        double Ixx = 0.0, Iyy = 0.0, Ixy = 0.0;
        for (size_t i = 0; i < reg.size(); ++i) {
            const double regx = reg[i].x, regy = reg[i].y, weight = reg[i].modgrad;
            const double dx = regx - x, dy = regy - y;
            Ixx += dy * dy * weight;
            Iyy += dx * dx * weight;
            Ixy -= dx * dy * weight;
        }
        const double lambda = 0.5 * (Ixx + Iyy - std::sqrt((Ixx - Iyy) * (Ixx - Iyy) + 4.0 * Ixy * Ixy));
        double theta = (std::fabs(Ixx) > std::fabs(Iyy)) ? double(fast_atan2_deg(float(lambda - Ixx), float(Ixy)))
                                                         : double(fast_atan2_deg(float(Ixy), float(lambda - Iyy)));
        theta *= kDegToRads;
        if (angle_diff(theta, reg_angle) > prec) theta += kPi;
        return theta;
    }

    void region2rect(const std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec) const {
        double x = 0, y = 0, sum = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            const double weight = reg[i].modgrad;
            x += double(reg[i].x) * weight;
            y += double(reg[i].y) * weight;
            sum += weight;
        }
        x /= sum;
        y /= sum;
        const double theta = get_theta(reg, x, y, reg_angle, prec);
        double dx, dy;
        det_sincos(theta, dy, dx);                 // std::cos / std::sin in OpenCV; see detmath.h
        if (libm_sincos) { dx = std::cos(theta); dy = std::sin(theta); }
        double l_min = 0, l_max = 0, w_min = 0, w_max = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            const double regdx = double(reg[i].x) - x, regdy = double(reg[i].y) - y;
            const double l = regdx * dx + regdy * dy;
            const double w = -regdx * dy + regdy * dx;
            if (l > l_max) l_max = l; else if (l < l_min) l_min = l;
            if (w > w_max) w_max = w; else if (w < w_min) w_min = w;
        }
        rec.x1 = x + l_min * dx; rec.y1 = y + l_min * dy;
        rec.x2 = x + l_max * dx; rec.y2 = y + l_max * dy;
        rec.width = w_max - w_min;
        rec.x = x; rec.y = y; rec.theta = theta; rec.dx = dx; rec.dy = dy; rec.prec = prec; rec.p = p;
        if (rec.width < 1.0) rec.width = 1.0;
    }

    bool reduce_region_radius(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density, double density_th) {
        const double xc = double(reg[0].x), yc = double(reg[0].y);
        const double radSq1 = dist_sq(xc, yc, rec.x1, rec.y1), radSq2 = dist_sq(xc, yc, rec.x2, rec.y2);
        double radSq = radSq1 > radSq2 ? radSq1 : radSq2;
        while (density < density_th) {
            radSq *= 0.75 * 0.75;
            for (size_t i = 0; i < reg.size(); ++i) {
                if (dist_sq(xc, yc, double(reg[i].x), double(reg[i].y)) > radSq) {
                    used[(size_t)reg[i].y * W + reg[i].x] = 0;
                    std::swap(reg[i], reg[reg.size() - 1]);
                    reg.pop_back();
                    --i;
                }
            }
            if (reg.size() < 2) return false;
            region2rect(reg, reg_angle, prec, p, rec);
            density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        }
        return true;
    }

    bool refine(std::vector<RegionPoint>& reg, double reg_angle, double prec, double p, Rect& rec, double density_th) {
        double density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density >= density_th) return true;
        const double xc = double(reg[0].x), yc = double(reg[0].y), ang_c = reg[0].angle;
        double sum = 0, s_sum = 0;
        int n = 0;
        for (size_t i = 0; i < reg.size(); ++i) {
            used[(size_t)reg[i].y * W + reg[i].x] = 0;
            if (dist(xc, yc, reg[i].x, reg[i].y) < rec.width) {
                const double ang_d = angle_diff_signed(reg[i].angle, ang_c);
                sum += ang_d;
                s_sum += ang_d * ang_d;
                ++n;
            }
        }
        const double mean_angle = sum / double(n);
        const double tau = 2.0 * std::sqrt((s_sum - 2.0 * mean_angle * sum) / double(n) + mean_angle * mean_angle);
        const int sx = reg[0].x, sy = reg[0].y;
        region_grow(sx, sy, reg, reg_angle, tau);
        if (reg.size() < 2) return false;
        region2rect(reg, reg_angle, prec, p, rec);
        density = double(reg.size()) / (dist(rec.x1, rec.y1, rec.x2, rec.y2) * rec.width);
        if (density < density_th) return reduce_region_radius(reg, reg_angle, prec, p, rec, density, density_th);
        return true;
    }

    double nfa(int n, int k, double p) const {
        if (n == 0 || k == 0) return -LOG_NT;
        if (n == k) return -LOG_NT - double(n) * std::log10(p);
        const double p_term = p / (1 - p);
        const double log1term = log_gamma(double(n) + 1) - log_gamma(double(k) + 1) - log_gamma(double(n - k) + 1) + double(k) * std::log(p) +
                                double(n - k) * std::log(1.0 - p);
        double term = std::exp(log1term);
        if (double_equal(term, 0)) {
            if (k > n * p) return -log1term / kLn10 - LOG_NT;
            return -LOG_NT;
        }
        double bin_tail = term;
        const double tolerance = 0.1;
        for (int i = k + 1; i <= n; ++i) {
            const double bin_term = double(n - i + 1) / double(i);
            const double mult_term = bin_term * p_term;
            term *= mult_term;
            bin_tail += term;
            if (bin_term < 1) {
                const double err = term * ((1 - std::pow(mult_term, double(n - i + 1))) / (1 - mult_term) - 1);
                if (err < tolerance * std::fabs(-std::log10(bin_tail) - LOG_NT) * bin_tail) break;
            }
        }
        return -std::log10(bin_tail) - LOG_NT;
    }


    // original LSD rectangle iterator (ri_ini / ri_inc / ri_end): integer points inside the rectangle, column by column
    void rect_iter_count(const Rect& r, int& n, int& k) const {
        double vx[4], vy[4], rx[4], ry[4];
        vx[0] = r.x1 - r.dy * r.width / 2.0; vy[0] = r.y1 + r.dx * r.width / 2.0;
        vx[1] = r.x2 - r.dy * r.width / 2.0; vy[1] = r.y2 + r.dx * r.width / 2.0;
        vx[2] = r.x2 + r.dy * r.width / 2.0; vy[2] = r.y2 - r.dx * r.width / 2.0;
        vx[3] = r.x1 + r.dy * r.width / 2.0; vy[3] = r.y1 - r.dx * r.width / 2.0;
        int offset;
        if (r.x1 < r.x2 && r.y1 <= r.y2) offset = 0;
        else if (r.x1 >= r.x2 && r.y1 < r.y2) offset = 1;
        else if (r.x1 > r.x2 && r.y1 >= r.y2) offset = 2;
        else offset = 3;
        for (int q = 0; q < 4; ++q) { rx[q] = vx[(offset + q) % 4]; ry[q] = vy[(offset + q) % 4]; }
        auto inter_low = [](double x, double x1, double y1, double x2, double y2) {
            if (double_equal(x1, x2) && y1 < y2) return y1;
            if (double_equal(x1, x2) && y1 > y2) return y2;
            return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
        };
        auto inter_hi = [](double x, double x1, double y1, double x2, double y2) {
            if (double_equal(x1, x2) && y1 < y2) return y2;
            if (double_equal(x1, x2) && y1 > y2) return y1;
            return y1 + (x - x1) * (y2 - y1) / (x2 - x1);
        };
        n = 0; k = 0;
        for (int x = (int)std::ceil(rx[0]); (double)x <= rx[2]; ++x) {
            const double ys = (double)x < rx[3] ? inter_low(x, rx[0], ry[0], rx[3], ry[3]) : inter_low(x, rx[3], ry[3], rx[2], ry[2]);
            const double ye = (double)x < rx[1] ? inter_hi(x, rx[0], ry[0], rx[1], ry[1]) : inter_hi(x, rx[1], ry[1], rx[2], ry[2]);
            for (int y = (int)std::ceil(ys); (double)y <= ye; ++y) {
                if (x >= 0 && y >= 0 && x < W && y < H) { ++n; if (is_aligned(x, y, r.theta, r.prec)) ++k; }
            }
        }
    }

    // cv2 4.13's rect_nfa enumeration (OpenCV >= 4.5.x imgproc lsd.cpp; no source in this image - recovered from the behaviour of the
    // in-container cv2 build and pinned by tests/test_oracle_lsd.py on whole detections, log-NFA included): the four corners as
    // doubles, rotated so that the first is the top one (smallest y, ties -> smallest x); rows from ceil(top.y) to ceil(bottom.y)
    // INCLUSIVE; per row the left bound follows top -> v1 -> bottom (switching after row ceil(v1.y)), the right bound follows
    // top -> v3 -> bottom (switching AT row ceil(v3.y)); an edge whose two ends round up to the same row has slope 0; columns from
    // ceil(left) to trunc(right).  Points outside the image are not counted.
    void rect_count_cv4(const Rect& r, int& n, int& k) const {
        n = 0; k = 0;
        rect_rows_cv4(r.x1, r.y1, r.x2, r.y2, r.width, r.dx, r.dy, W, H, [&](int y, int xa, int xb) {
            for (int x = xa; x <= xb; ++x) {
                ++n;
                if (is_aligned(x, y, r.theta, r.prec)) ++k;
            }
        });
    }

    // Which enumeration rect_nfa uses.  0: the published LSD rectangle iterator (what the CUDA path implements this round);
    // 1: cv2 4.13's (above).  NONE / STD never call rect_nfa.  With 1 the oracle agrees with cv2 4.13 LSD_REFINE_ADV bit for bit;
    // with 0 the accepted set differs for short, barely meaningful segments while the 40 longest - all the reference consumes
    // (src/LSDextractor.cpp:18-26) - agree (tests/test_oracle_lsd.py).  The reference's own pinned OpenCV 3.4.1 predates the
    // 4.x rewrite and enumerates with integer-truncated corners; no build of it is available here, so that variant is unpinned
    // (DESIGN.md §5.7).
    int rect_enum = 0;                 // bit 0 of the rect_enum argument
    // bit 1 of the rect_enum argument: rectangle axes from the host libm (what OpenCV calls) instead of detmath.h.  The two differ by 1 ulp now and
    // then, which moves a scan-line bound across an integer for about one rectangle in a few thousand; with the libm axes the
    // oracle reproduces cv2 4.13 LSD_REFINE_ADV in every field.  The CUDA path has no libm, hence detmath.h as the default.
    bool libm_sincos = false;
    bool debug = std::getenv("ORC_LSD_DEBUG") != nullptr;
    double rect_nfa(const Rect& rec) const {
        int total_pts = 0, alg_pts = 0;
        if (rect_enum == 1) rect_count_cv4(rec, total_pts, alg_pts);
        else rect_iter_count(rec, total_pts, alg_pts);
        if (debug) std::fprintf(stderr, "rect_nfa %.17g %.17g %.17g %.17g w %.17g th %.17g dx %.17g dy %.17g prec %.17g p %.17g n %d k %d nfa %.17g\n", rec.x1, rec.y1,
                                rec.x2, rec.y2, rec.width, rec.theta, rec.dx, rec.dy, rec.prec, rec.p, total_pts, alg_pts, nfa(total_pts, alg_pts, rec.p));
        return nfa(total_pts, alg_pts, rec.p);
    }

    double rect_improve(Rect& rec) const {
        const double delta = 0.5, delta_2 = delta / 2.0;
        double log_nfa = rect_nfa(rec);
        if (log_nfa > LOG_EPS) return log_nfa;
        Rect r = rec;
        for (int n = 0; n < 5; ++n) {
            r.p /= 2;
            r.prec = r.p * kPi;
            const double v = rect_nfa(r);
            if (v > log_nfa) { log_nfa = v; rec = r; }
        }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.width -= delta;
                const double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 += -r.dy * delta_2; r.y1 += r.dx * delta_2;
                r.x2 += -r.dy * delta_2; r.y2 += r.dx * delta_2;
                r.width -= delta;
                const double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.x1 -= -r.dy * delta_2; r.y1 -= r.dx * delta_2;
                r.x2 -= -r.dy * delta_2; r.y2 -= r.dx * delta_2;
                r.width -= delta;
                const double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        if (log_nfa > LOG_EPS) return log_nfa;
        r = rec;
        for (unsigned n = 0; n < 5; ++n)
            if ((r.width - delta) >= 0.5) {
                r.p /= 2;
                r.prec = r.p * kPi;
                const double v = rect_nfa(r);
                if (v > log_nfa) { rec = r; log_nfa = v; }
            }
        return log_nfa;
    }
};

}  // namespace

int lsd_cv4_spans(const double* rect, int W, int H, int32_t* rows, int cap) {
    int m = 0;
    rect_rows_cv4(rect[0], rect[1], rect[2], rect[3], rect[4], rect[5], rect[6], W, H, [&](int y, int xa, int xb) {
        if (m < cap) { rows[3 * m] = y; rows[3 * m + 1] = xa; rows[3 * m + 2] = xb; }
        ++m;
    });
    return m;
}

void lsd_detect_stages(const Img8& img, int refine, std::vector<LsdSegment>& out, LsdStages& st, int rect_enum) {
    out.clear();
    const double SCALE = 0.8, SIGMA_SCALE = 0.6, QUANT = 2.0, ANG_TH = 22.5;
    const int N_BINS = 1024;
    const double prec = kPi * ANG_TH / 180, p = ANG_TH / 180, rho = QUANT / std::sin(prec);
    (void)SIGMA_SCALE;
    Lsd L;
    L.rect_enum = rect_enum & 1;
    L.libm_sincos = (rect_enum & 2) != 0;
    // sigma = 0.6 / 0.8 = 0.75, h = ceil(0.75 * sqrt(2 * 3 * ln 10)) = 3 -> 7x7 kernel
    st.blurred.resize((size_t)img.w * img.h);
    gaussian_blur_7x7_s075_u8(img, st.blurred.data());
    L.W = cv_round(img.w * SCALE); L.H = cv_round(img.h * SCALE);
    L.scaled.resize((size_t)L.W * L.H);
    resize_linear_exact_u8(st.blurred.data(), img.w, img.h, L.scaled.data(), L.W, L.H, 1.0 / SCALE, 1.0 / SCALE);
    L.ll_angle(rho, N_BINS);
    L.LOG_NT = 5 * (std::log10(double(L.W)) + std::log10(double(L.H))) / 2 + std::log10(11.0);
    const size_t min_reg_size = size_t(-L.LOG_NT / std::log10(p));
    L.used.assign((size_t)L.W * L.H, 0);
    st.w = L.W; st.h = L.H;
    st.region_id.assign((size_t)L.W * L.H, -1);
    std::vector<RegionPoint> reg;
    for (size_t i = 0; i < L.ordered.size(); ++i) {
        const int px = L.ordered[i].x, py = L.ordered[i].y;
        const size_t a = (size_t)py * L.W + px;
        if (L.used[a] != 0 || L.angles[a] == kNotDef) continue;
        double reg_angle;
        L.region_grow(px, py, reg, reg_angle, prec);
        if (reg.size() < min_reg_size) continue;
        Rect rec;
        L.region2rect(reg, reg_angle, prec, p, rec);
        double log_nfa = -1;
        if (refine > 0) {
            if (!L.refine(reg, reg_angle, prec, p, rec, L.DENSITY_TH)) continue;
            if (refine >= 2) {
                log_nfa = L.rect_improve(rec);
                if (log_nfa <= L.LOG_EPS) continue;
            }
        }
        rec.x1 += 0.5; rec.y1 += 0.5; rec.x2 += 0.5; rec.y2 += 0.5;
        rec.x1 /= SCALE; rec.y1 /= SCALE; rec.x2 /= SCALE; rec.y2 /= SCALE; rec.width /= SCALE;
        for (const RegionPoint& q : reg) st.region_id[(size_t)q.y * L.W + q.x] = (int32_t)out.size();
        out.push_back({float(rec.x1), float(rec.y1), float(rec.x2), float(rec.y2), rec.width, rec.p, log_nfa});
    }
    st.scaled = L.scaled; st.modgrad = L.modgrad; st.angles = L.angles;
    st.order.resize(L.ordered.size());
    for (size_t i = 0; i < L.ordered.size(); ++i) st.order[i] = L.ordered[i].y * L.W + L.ordered[i].x;
}

void lsd_detect(const Img8& img, int refine, std::vector<LsdSegment>& out, int rect_enum) {
    LsdStages st;
    lsd_detect_stages(img, refine, out, st, rect_enum);
}

void extract_line_segments(const Img8& img, int max_lines, std::vector<KeyLine>& kls, std::vector<double>& lf, int rect_enum) {
    std::vector<LsdSegment> segs;
    lsd_detect(img, 2, segs, rect_enum);
    kls.clear();
    int class_counter = -1;
    for (const LsdSegment& s : segs) {
        float e[4] = {s.x1, s.y1, s.x2, s.y2};
        // checkLineExtremes (LSDDetector.cpp): clamp into the image
        if (e[0] < 0) e[0] = 0; if (e[0] >= img.w) e[0] = (float)img.w - 1.0f;
        if (e[2] < 0) e[2] = 0; if (e[2] >= img.w) e[2] = (float)img.w - 1.0f;
        if (e[1] < 0) e[1] = 0; if (e[1] >= img.h) e[1] = (float)img.h - 1.0f;
        if (e[3] < 0) e[3] = 0; if (e[3] >= img.h) e[3] = (float)img.h - 1.0f;
        KeyLine k;
        k.startPointX = e[0]; k.startPointY = e[1]; k.endPointX = e[2]; k.endPointY = e[3];       // octaveScale = 1
        k.sPointInOctaveX = e[0]; k.sPointInOctaveY = e[1]; k.ePointInOctaveX = e[2]; k.ePointInOctaveY = e[3];
        k.lineLength = (float)std::sqrt(std::pow(e[0] - e[2], 2) + std::pow(e[1] - e[3], 2));
        // cv::LineIterator(img, Point(cvRound ...), connectivity 8).count for end points inside the image
        const int ax = cv_round(e[0]), ay = cv_round(e[1]), bx = cv_round(e[2]), by = cv_round(e[3]);
        k.numOfPixels = std::max(std::abs(bx - ax), std::abs(by - ay)) + 1;
        k.angle = (float)std::atan2((k.endPointY - k.startPointY), (k.endPointX - k.startPointX));
        k.class_id = ++class_counter;
        k.octave = 0;
        k.size = (k.endPointX - k.startPointX) * (k.endPointY - k.startPointY);
        k.response = k.lineLength / std::max(img.w, img.h);
        k.pt_x = (k.endPointX + k.startPointX) / 2; k.pt_y = (k.endPointY + k.startPointY) / 2;
        kls.push_back(k);
    }
    // src/LSDextractor.cpp:18-26: sort by response (descending), keep the first max_lines, renumber class_id
    if ((int)kls.size() > max_lines) {
        std::sort(kls.begin(), kls.end(), [](const KeyLine& a, const KeyLine& b) { return a.response > b.response; });
        kls.resize(max_lines);
        for (int i = 0; i < max_lines; ++i) kls[i].class_id = i;
    }
    lf.clear();
    for (const KeyLine& k : kls) {               // :30-38
        const double sp[3] = {k.startPointX, k.startPointY, 1.0}, ep[3] = {k.endPointX, k.endPointY, 1.0};
        double l[3] = {sp[1] * ep[2] - sp[2] * ep[1], sp[2] * ep[0] - sp[0] * ep[2], sp[0] * ep[1] - sp[1] * ep[0]};
        const double n = std::sqrt(l[0] * l[0] + l[1] * l[1] + l[2] * l[2]);
        for (int i = 0; i < 3; ++i) lf.push_back(l[i] / n);
    }
}

}  // namespace oracle
