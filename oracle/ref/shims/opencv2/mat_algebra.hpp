// TEST INFRASTRUCTURE ONLY.  Dense matrix algebra of the cv::Mat stand-in (double / float, small matrices) for the reference's
// src/LineExtractor.cpp: products accumulate left to right over k in double like cv::gemm's generic path, 3x3 inverse by cofactors like
// cv::invert, cv::SVD = OpenCV's Jacobi algorithm as restated in oracle/cvsvd.h.  Included from opencv.hpp inside no namespace.
#pragma once
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <ostream>
#include "../../../match.h"
#include "../../../cvsvd.h"

struct CvMat;                                            // legacy C type named by include/PnPsolver.h member declarations
namespace cv {

inline Mat newLike(int r, int c, int type) { return Mat(MatZeros{r, c, type, 0}); }
inline Mat operator*(const Mat& a, const Mat& b) {
    Mat m = newLike(a.rows, b.cols, a.type());
    for (int i = 0; i < a.rows; ++i)
        for (int j = 0; j < b.cols; ++j) {
            double s = 0;
            for (int k = 0; k < a.cols; ++k) s += a.getd(i, k) * b.getd(k, j);
            m.setd(i, j, s);
        }
    return m;
}
#define PSLAM_SHIM_EWISE(NAME, EXPR)                                                           \
    inline Mat NAME(const Mat& a, const Mat& b) {                                              \
        Mat m = newLike(a.rows, a.cols, a.type());                                             \
        for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) { const double x = a.getd(i, j), y = b.getd(i, j); m.setd(i, j, EXPR); } \
        return m;                                                                              \
    }
PSLAM_SHIM_EWISE(operator+, x + y)
PSLAM_SHIM_EWISE(operator-, x - y)
#undef PSLAM_SHIM_EWISE
inline Mat scaleMat(const Mat& a, double s) {
    Mat m = newLike(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.setd(i, j, a.getd(i, j) * s);
    return m;
}
inline Mat operator*(const Mat& a, double s) { return scaleMat(a, s); }
inline Mat operator*(double s, const Mat& a) { return scaleMat(a, s); }
inline Mat operator/(const Mat& a, double s) { return scaleMat(a, 1. / s); }         // MatExpr: a * (1 / s)
inline Mat operator-(const Mat& a) { return scaleMat(a, -1.0); }
inline Mat operator+(const Mat& a, double v) {
    Mat m = newLike(a.rows, a.cols, a.type());
    for (int i = 0; i < a.rows; ++i) for (int j = 0; j < a.cols; ++j) m.setd(i, j, a.getd(i, j) + v);
    return m;
}
inline Mat operator+(const MatZeros& z, double v) { return Mat(z) + v; }
inline Mat Mat::t() const {
    Mat m = newLike(cols, rows, type());
    for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) m.setd(j, i, getd(i, j));
    return m;
}
inline double Mat::dot(const Mat& o) const {
    double s = 0;
    for (int i = 0; i < rows; ++i) for (int j = 0; j < cols; ++j) s += getd(i, j) * o.getd(i, j);
    return s;
}
inline Mat Mat::cross(const Mat& o) const {
    auto e = [](const Mat& m, int i) { return m.cols == 1 ? m.getd(i, 0) : m.getd(0, i); };
    Mat m = newLike(rows, cols, type());
    const double a0 = e(*this, 0), a1 = e(*this, 1), a2 = e(*this, 2), b0 = e(o, 0), b1 = e(o, 1), b2 = e(o, 2);
    const double c[3] = {a1 * b2 - a2 * b1, a2 * b0 - a0 * b2, a0 * b1 - a1 * b0};
    for (int i = 0; i < 3; ++i) { if (cols == 1) m.setd(i, 0, c[i]); else m.setd(0, i, c[i]); }
    return m;
}
inline double determinant(const Mat& m) {
    if (m.rows == 2) return m.getd(0, 0) * m.getd(1, 1) - m.getd(0, 1) * m.getd(1, 0);
    return m.getd(0, 0) * (m.getd(1, 1) * m.getd(2, 2) - m.getd(1, 2) * m.getd(2, 1)) - m.getd(0, 1) * (m.getd(1, 0) * m.getd(2, 2) - m.getd(1, 2) * m.getd(2, 0)) +
           m.getd(0, 2) * (m.getd(1, 0) * m.getd(2, 1) - m.getd(1, 1) * m.getd(2, 0));
}
inline Mat Mat::inv(int) const {
    const int n = rows;
    Mat m = newLike(n, n, type());
    if (n == 3) {
        const double d = determinant(*this), id = d != 0 ? 1. / d : 0;
        auto a = [&](int i, int j) { return getd(i, j); };
        const double t[9] = {(a(1, 1) * a(2, 2) - a(1, 2) * a(2, 1)) * id, (a(0, 2) * a(2, 1) - a(0, 1) * a(2, 2)) * id, (a(0, 1) * a(1, 2) - a(0, 2) * a(1, 1)) * id,
                             (a(1, 2) * a(2, 0) - a(1, 0) * a(2, 2)) * id, (a(0, 0) * a(2, 2) - a(0, 2) * a(2, 0)) * id, (a(0, 2) * a(1, 0) - a(0, 0) * a(1, 2)) * id,
                             (a(1, 0) * a(2, 1) - a(1, 1) * a(2, 0)) * id, (a(0, 1) * a(2, 0) - a(0, 0) * a(2, 1)) * id, (a(0, 0) * a(1, 1) - a(0, 1) * a(1, 0)) * id};
        for (int i = 0; i < 9; ++i) m.setd(i / 3, i % 3, t[i]);
        return m;
    }
    std::vector<double> A((size_t)n * 2 * n, 0.0);                      // Gauss-Jordan with partial pivoting
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) A[(size_t)i * 2 * n + j] = getd(i, j); A[(size_t)i * 2 * n + n + i] = 1; }
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::fabs(A[(size_t)r * 2 * n + c]) > std::fabs(A[(size_t)p * 2 * n + c])) p = r;
        for (int j = 0; j < 2 * n; ++j) std::swap(A[(size_t)c * 2 * n + j], A[(size_t)p * 2 * n + j]);
        const double d = A[(size_t)c * 2 * n + c];
        if (d == 0) return newLike(n, n, type());
        for (int j = 0; j < 2 * n; ++j) A[(size_t)c * 2 * n + j] /= d;
        for (int r = 0; r < n; ++r) if (r != c) { const double f = A[(size_t)r * 2 * n + c]; if (f != 0) for (int j = 0; j < 2 * n; ++j) A[(size_t)r * 2 * n + j] -= f * A[(size_t)c * 2 * n + j]; }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) m.setd(i, j, A[(size_t)i * 2 * n + n + j]);
    return m;
}
inline double norm(const Mat& m) { return std::sqrt(m.dot(m)); }
inline double norm(const Mat& a, const Mat& b) { return norm(a - b); }
inline double norm(const Mat& a, const Mat& b, int normType) {          // NORM_HAMMING (6) over 8-bit rows, else L2
    if (normType != 6) return norm(a - b);
    int d = 0;
    for (int r = 0; r < a.rows; ++r) for (int c = 0; c < a.cols; ++c) d += __builtin_popcount((unsigned)(a.ptr(r)[c] ^ b.ptr(r)[c]));
    return d;
}
inline void transpose(const Mat& src, Mat& dst) { dst = src.t(); }
inline std::ostream& operator<<(std::ostream& os, const Mat::SizeTag&) { return os << "[size]"; }
struct Scalar { double v[4] = {0, 0, 0, 0}; Scalar() {} Scalar(double a, double b = 0, double c = 0, double d = 0) { v[0] = a; v[1] = b; v[2] = c; v[3] = d; }
                double& operator[](int i) { return v[i]; } const double& operator[](int i) const { return v[i]; } static Scalar all(double a) { return Scalar(a, a, a, a); } };
// cv::sum / cv::trace: one channel, accumulated in double in row-major order like OpenCV's sum_ / the diagonal walk of cv::trace
inline Scalar sum(const Mat& m) { double a = 0; for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) a += m.getd(r, c); return Scalar(a); }
inline Scalar trace(const Mat& m) { double a = 0; for (int i = 0; i < m.rows && i < m.cols; ++i) a += m.getd(i, i); return Scalar(a); }
enum { CV_BGR2GRAY = 6, CV_RGB2GRAY = 7, CV_BGRA2GRAY = 10, CV_RGBA2GRAY = 11 };
[[noreturn]] inline void shim_unreachable(const char* what) { std::fprintf(stderr, "oracle/ref/shims: %s is a compile-only stand-in\n", what); std::abort(); }
inline void cvtColor(const Mat&, Mat&, int) { shim_unreachable("cv::cvtColor"); }
template <class... A> inline void initUndistortRectifyMap(A&&...) { shim_unreachable("cv::initUndistortRectifyMap"); }
struct RNG { enum { UNIFORM = 0, NORMAL = 1 }; explicit RNG(uint64_t = 0) {} void fill(Mat&, int, const Scalar&, const Scalar&) { shim_unreachable("cv::RNG::fill"); } };
template <class E> inline void eigen2cv(const E& src, Mat& dst) { dst = newLike((int)src.rows(), (int)src.cols(), CV_64F); for (int r = 0; r < dst.rows; ++r) for (int c = 0; c < dst.cols; ++c) dst.setd(r, c, (double)src(r, c)); }
inline Mat& operator/=(Mat& m, double s) { for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) m.setd(r, c, m.depth() == CV_64F ? m.getd(r, c) / s : (double)((float)m.getd(r, c) / (float)s)); return m; }
inline std::ostream& operator<<(std::ostream& os, const Mat& m) {
    os << "[";
    for (int r = 0; r < m.rows; ++r) { for (int c = 0; c < m.cols; ++c) os << (c ? ", " : "") << m.getd(r, c); os << (r + 1 < m.rows ? ";\n " : ""); }
    return os << "]";
}
inline void undistortPoints(const Mat&, Mat&, const Mat&, const Mat&, const Mat&, const Mat&) { std::fprintf(stderr, "oracle/ref/shims: cv::undistortPoints is a compile-only stand-in (zero distortion returns before it, Frame.cc:543-547)\n"); std::abort(); }
inline void minMaxLoc(const Mat& m, double* minVal, double* maxVal = nullptr, Point* minLoc = nullptr, Point* maxLoc = nullptr) {
    double mn = m.getd(0, 0), mx = mn;
    Point pn(0, 0), px(0, 0);
    for (int i = 0; i < m.rows; ++i) for (int j = 0; j < m.cols; ++j) { const double v = m.getd(i, j); if (v < mn) { mn = v; pn = Point(j, i); } if (v > mx) { mx = v; px = Point(j, i); } }
    if (minVal) *minVal = mn; if (maxVal) *maxVal = mx; if (minLoc) *minLoc = pn; if (maxLoc) *maxLoc = px;
}

template <class T> class Mat_;
template <class T>
struct MatCommaInitializer_ {
    Mat_<T>* m; int idx;
    template <class U> MatCommaInitializer_& operator,(U v) { (*m)(idx / m->cols, idx % m->cols) = (T)v; ++idx; return *this; }
    operator Mat_<T>() const { return *m; }
    operator Mat() const { return *m; }
    Mat t() const { return Mat(*m).t(); }
};
template <class T>
class Mat_ : public Mat {
public:
    Mat_() {}
    Mat_(int r, int c) : Mat(MatZeros{r, c, sizeof(T) == 8 ? CV_64F : (sizeof(T) == 4 ? CV_32F : CV_8U), 0}) {}
    Mat_(const Mat& m) : Mat(m) {}
    static Mat_ eye(int r, int c) { return Mat_(Mat::eye(r, c, sizeof(T) == 8 ? CV_64F : CV_32F)); }
    T& operator()(int r, int c) { return this->template at<T>(r, c); }
    const T& operator()(int r, int c) const { return this->template at<T>(r, c); }
    template <class U> MatCommaInitializer_<T> operator<<(U v) { (*this)(0, 0) = (T)v; return MatCommaInitializer_<T>{this, 1}; }
};
template <class T> inline Mat operator*(const Mat& a, const MatCommaInitializer_<T>& b) { return a * Mat(b); }

class SVD {
public:
    enum { MODIFY_A = 1, NO_UV = 2, FULL_UV = 4 };
    Mat u, w, vt;
    SVD() {}
    SVD(const Mat& src, int flags = 0) { compute(src, w, u, vt, flags); }
    static void compute(const Mat& src, Mat& w, Mat& u, Mat& vt, int = 0) {
        const int m = src.rows, n = src.cols, k = m < n ? m : n;
        std::vector<double> A((size_t)m * n), W(k), U((size_t)m * k), Vt((size_t)k * n);
        for (int i = 0; i < m; ++i) for (int j = 0; j < n; ++j) A[(size_t)i * n + j] = src.getd(i, j);
        oracle::cv_svd<double>(A.data(), m, n, W.data(), U.data(), Vt.data());
        w = newLike(k, 1, CV_64F); u = newLike(m, k, CV_64F); vt = newLike(k, n, CV_64F);
        for (int i = 0; i < k; ++i) w.setd(i, 0, W[i]);
        for (int i = 0; i < m; ++i) for (int j = 0; j < k; ++j) u.setd(i, j, U[(size_t)i * k + j]);
        for (int i = 0; i < k; ++i) for (int j = 0; j < n; ++j) vt.setd(i, j, Vt[(size_t)i * n + j]);
    }
};

struct DMatch { int queryIdx = -1, trainIdx = -1, imgIdx = -1; float distance = 0; };
enum { NORM_HAMMING = 6 };
class BFMatcher {                                        // brute force, first minimum wins: the oracle's cv2-pinned one (oracle/match.h)
public:
    explicit BFMatcher(int = NORM_HAMMING, bool = false) {}
    void match(const Mat& q, const Mat& t, std::vector<DMatch>& out) const {
        std::vector<int32_t> idx(q.rows), dist(q.rows);
        oracle::bf_match(q.ptr<uint8_t>(0), q.rows, t.ptr<uint8_t>(0), t.rows, idx.data(), dist.data());
        out.clear();
        for (int i = 0; i < q.rows; ++i) if (idx[i] >= 0) { DMatch m; m.queryIdx = i; m.trainIdx = idx[i]; m.imgIdx = 0; m.distance = (float)dist[i]; out.push_back(m); }
    }
    void knnMatch(const Mat& q, const Mat& t, std::vector<std::vector<DMatch>>& out, int k) const {
        std::vector<int32_t> idx(2 * (size_t)q.rows), dist(2 * (size_t)q.rows);
        oracle::bf_knn2(q.ptr<uint8_t>(0), q.rows, t.ptr<uint8_t>(0), t.rows, idx.data(), dist.data());
        out.assign(q.rows, {});
        for (int i = 0; i < q.rows; ++i) for (int j = 0; j < 2 && j < k; ++j) if (idx[2 * i + j] >= 0) { DMatch m; m.queryIdx = i; m.trainIdx = idx[2 * i + j]; m.imgIdx = 0; m.distance = (float)dist[2 * i + j]; out[i].push_back(m); }
    }
};
template <class T> using Ptr = std::shared_ptr<T>;

class LineIterator {                                     // only met in code the tracker never reaches
public:
    int count = 0;
    LineIterator(const Mat&, Point a, Point b, int = 8) : p_(a), a_(a), b_(b) { const int dx = std::abs(b.x - a.x), dy = std::abs(b.y - a.y); count = (dx > dy ? dx : dy) + 1; }
    Point pos() const { return p_; }
    LineIterator& operator++() { ++i_; const double t = count > 1 ? (double)i_ / (count - 1) : 0; p_ = Point((int)std::lround(a_.x + t * (b_.x - a_.x)), (int)std::lround(a_.y + t * (b_.y - a_.y))); return *this; }
    LineIterator operator++(int) { LineIterator c = *this; ++*this; return c; }
private:
    Point p_, a_, b_; int i_ = 0;
};

namespace line_descriptor {
struct KeyLine {
    float angle = 0; int class_id = -1; int octave = 0; Point2f pt; float response = 0, size = 0;
    float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0, sPointInOctaveX = 0, sPointInOctaveY = 0, ePointInOctaveX = 0, ePointInOctaveY = 0;
    float lineLength = 0; int numOfPixels = 0;
    Point2f getStartPoint() const { return Point2f(startPointX, startPointY); }
    Point2f getEndPoint() const { return Point2f(endPointX, endPointY); }
};
class LSDDetector { public: static Ptr<LSDDetector> createLSDDetector() { return Ptr<LSDDetector>(); } };
class BinaryDescriptor { public: static Ptr<BinaryDescriptor> createBinaryDescriptor() { return Ptr<BinaryDescriptor>(); } };
}  // namespace line_descriptor

}  // namespace cv
