// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the few OpenCV container types the reference's PEAC headers and
// src/PlaneExtractor.cpp touch (cv::Mat as a typed byte buffer with ROI views, Vec3b, Range, tick counter), so that those
// reference sources compile UNMODIFIED from /root/reference in a container without OpenCV headers (oracle/Makefile target _ref).
// No image arithmetic lives here: every number the plane extractor produces comes from the reference's own code.
#pragma once
#include <algorithm>
#include <chrono>
#include <cassert>
#include <fstream>
#include <iostream>
#include <sstream>
#include <cmath>
#include <math.h>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)
#define CV_64FC1 CV_MAKETYPE(CV_64F, 1)

namespace cv {

struct Vec3b {
    uchar v[3];
    Vec3b() : v{0, 0, 0} {}
    Vec3b(uchar a, uchar b, uchar c) : v{a, b, c} {}
    explicit Vec3b(const uchar* p) : v{p[0], p[1], p[2]} {}
    uchar& operator[](int i) { return v[i]; }
    const uchar& operator[](int i) const { return v[i]; }
};
struct Vec2d { double v[2]; Vec2d() : v{0, 0} {} Vec2d(double a, double b) : v{a, b} {} };
template <class T>
struct Point3_ {
    T x, y, z;
    Point3_() : x(0), y(0), z(0) {}
    Point3_(T a, T b, T c) : x(a), y(b), z(c) {}
    template <class U> Point3_(const Point3_<U>& o) : x((T)o.x), y((T)o.y), z((T)o.z) {}
    T dot(const Point3_& o) const { return (T)(x * o.x + y * o.y + z * o.z); }
    double ddot(const Point3_& o) const { return (double)x * o.x + (double)y * o.y + (double)z * o.z; }
    Point3_ cross(const Point3_& o) const { return Point3_(y * o.z - z * o.y, z * o.x - x * o.z, x * o.y - y * o.x); }
};
typedef Point3_<float> Point3f;
typedef Point3_<double> Point3d;
template <class T> inline Point3_<T> operator+(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <class T> inline Point3_<T> operator-(const Point3_<T>& a, const Point3_<T>& b) { return Point3_<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <class T> inline Point3_<T> operator-(const Point3_<T>& a) { return Point3_<T>(-a.x, -a.y, -a.z); }
template <class T> inline Point3_<T> operator*(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x * s), (T)(a.y * s), (T)(a.z * s)); }
template <class T> inline Point3_<T> operator*(double s, const Point3_<T>& a) { return a * s; }
template <class T> inline Point3_<T> operator/(const Point3_<T>& a, double s) { return Point3_<T>((T)(a.x / s), (T)(a.y / s), (T)(a.z / s)); }
template <class T> inline Point3_<T>& operator+=(Point3_<T>& a, const Point3_<T>& b) { a = a + b; return a; }
template <class T> inline double norm(const Point3_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y + (double)p.z * p.z); }
template <class T>
struct Point_ {
    T x, y;
    Point_() : x(0), y(0) {}
    Point_(T a, T b) : x(a), y(b) {}
    template <class U> Point_(const Point_<U>& o) : x((T)o.x), y((T)o.y) {}
    Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }        // Point2f *= float: float products
    T dot(const Point_& o) const { return (T)(x * o.x + y * o.y); }
    double cross(const Point_& o) const { return (double)x * o.y - (double)y * o.x; }
};
template <class T> inline Point_<T> operator+(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x + b.x, a.y + b.y); }
template <class T> inline Point_<T> operator-(const Point_<T>& a, const Point_<T>& b) { return Point_<T>(a.x - b.x, a.y - b.y); }
template <class T> inline Point_<T> operator-(const Point_<T>& a) { return Point_<T>(-a.x, -a.y); }
template <class T> inline Point_<T> operator*(const Point_<T>& a, double s) { return Point_<T>((T)(a.x * s), (T)(a.y * s)); }
template <class T> inline Point_<T> operator*(double s, const Point_<T>& a) { return a * s; }
template <class T> inline Point_<T> operator/(const Point_<T>& a, double s) { return Point_<T>((T)(a.x / s), (T)(a.y / s)); }
template <class T> inline bool operator==(const Point_<T>& a, const Point_<T>& b) { return a.x == b.x && a.y == b.y; }
template <class T> inline double norm(const Point_<T>& p) { return std::sqrt((double)p.x * p.x + (double)p.y * p.y); }
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;
typedef Point_<double> Point2d;
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
struct Rect { int x = 0, y = 0, width = 0, height = 0; Rect() {} Rect(int a, int b, int w, int h) : x(a), y(b), width(w), height(h) {} };
struct KeyPoint {
    Point2f pt; float size = 0, angle = -1, response = 0; int octave = 0, class_id = -1;
    KeyPoint() {}
    KeyPoint(float x, float y, float s, float a = -1, float r = 0, int o = 0, int c = -1) : pt(x, y), size(s), angle(a), response(r), octave(o), class_id(c) {}
};
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

inline int64_t getTickCount() { return (int64_t)std::chrono::steady_clock::now().time_since_epoch().count(); }
inline double getTickFrequency() { return (double)std::chrono::steady_clock::period::den / std::chrono::steady_clock::period::num; }

struct MatZeros { int rows, cols, type, value; };          // Mat::zeros / Mat::ones: like cv::MatExpr, assigning one to a matrix that already
                                                           // has that size and type fills it IN PLACE (src/ORBextractor.cc:1037 relies on it)
class Mat {
public:
    int rows = 0, cols = 0;
    size_t step = 0;
    Mat() {}
    Mat(Size sz, int type) { create(sz.height, sz.width, type); }
    Mat(const MatZeros& e) { *this = e; }
    Mat& operator=(const MatZeros& e) { create(e.rows, e.cols, e.type); for (int r = 0; r < rows; ++r) std::memset(data_ + (size_t)r * step, 0, (size_t)cols * esz(type_)); if (e.value) setTo(e.value); return *this; }
    static MatZeros zeros(int r, int c, int type) { return MatZeros{r, c, type, 0}; }
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext, size_t stp = 0) : rows(r), cols(c), step(stp ? stp : (size_t)c * esz(type)), type_(type), data_((uchar*)ext) {}   // external data, not owned
    static MatZeros ones(int r, int c, int type) { return MatZeros{r, c, type, 1}; }
    void create(int r, int c, int type) {
        if (data_ && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step = (size_t)c * esz(type);
        owner_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step + 8], std::default_delete<uchar[]>());
        data_ = owner_.get();
    }
    void release() { owner_.reset(); data_ = nullptr; rows = cols = 0; }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
    int depth() const { return type_ & 7; }
    int type() const { return type_; }
    template <class T> T& at(int r, int c) { return *(T*)(data_ + (size_t)r * step + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *(const T*)(data_ + (size_t)r * step + (size_t)c * sizeof(T)); }
    Mat operator()(const Range& rr, const Range& cr) const {
        Mat m; m.rows = rr.end - rr.start; m.cols = cr.end - cr.start; m.type_ = type_; m.step = step; m.owner_ = owner_;
        m.data_ = data_ + (size_t)rr.start * step + (size_t)cr.start * esz(type_);
        return m;
    }
    Mat& setTo(int value) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                uchar* p = data_ + (size_t)r * step + (size_t)c * esz(type_);
                switch (depth()) {
                    case CV_8U: for (int k = 0; k < channels(); ++k) p[k] = k == 0 ? (uchar)value : 0; break;
                    case CV_16U: *(uint16_t*)p = (uint16_t)value; break;
                    case CV_32S: *(int32_t*)p = value; break;
                    default: *(float*)p = (float)value; break;
                }
            }
        return *this;
    }
    Mat& setTo(const Vec3b& value) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) std::memcpy(data_ + (size_t)r * step + (size_t)c * 3, value.v, 3);
        return *this;
    }
    Mat rowRange(int a, int b) const { return (*this)(Range(a, b), Range(0, cols)); }
    Mat colRange(int a, int b) const { return (*this)(Range(0, rows), Range(a, b)); }
    Mat operator()(const Rect& r) const { return (*this)(Range(r.y, r.y + r.height), Range(r.x, r.x + r.width)); }
    Mat clone() const {                                         // continuous copy (step = cols * elemSize)
        Mat m(rows, cols, type_);
        for (int r = 0; r < rows; ++r) std::memcpy(m.data_ + (size_t)r * m.step, data_ + (size_t)r * step, (size_t)cols * esz(type_));
        return m;
    }
    template <class T, class P> T& at(const P& p) const { return const_cast<Mat*>(this)->at<T>((int)p.y, (int)p.x); }     // at<T>(cv::Point)
    uchar* ptr(int r = 0) const { return data_ + (size_t)r * step; }
    template <class T> T* ptr(int r = 0) const { return (T*)(data_ + (size_t)r * step); }
    size_t step1() const { static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return step / d[type_ & 7]; }
    Mat col(int c) const { return colRange(c, c + 1); }
    Mat row(int r) const { return rowRange(r, r + 1); }
    Mat t() const;
    Mat inv(int = 0) const;
    double dot(const Mat& o) const;
    Mat cross(const Mat& o) const;
    void copyTo(Mat& dst) const { dst.create(rows, cols, type_); copyRows(dst); }
    void copyTo(Mat&& view) const { copyRows(view); }             // into a row / column / ROI view of the same size
    static Mat eye(Size sz, int type) { return eye(sz.height, sz.width, type); }
    void convertTo(Mat& dst, int type, double alpha = 1, double beta = 0) const {      // single channel, to float / double (Frame.cc:82 depth scaling)
        Mat out(rows, cols, type);
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                double v;
                switch (depth()) { case CV_8U: v = at<uchar>(r, c); break; case CV_16U: v = at<uint16_t>(r, c); break; case CV_32S: v = at<int32_t>(r, c); break;
                                   case CV_64F: v = at<double>(r, c); break; default: v = at<float>(r, c); }
                out.setd(r, c, v * alpha + beta);
            }
        dst = out;
    }
    Mat reshape(int cn) const {                                   // same rows, channels regrouped (Frame.cc:559-561); continuous data only
        Mat m = *this; const int total = cols * channels();
        m.type_ = (type_ & 7) | ((cn - 1) << 3); m.cols = total / cn;
        return m;
    }
    static Mat eye(int r, int c, int type) { Mat m(MatZeros{r, c, type, 0}); for (int i = 0; i < r && i < c; ++i) m.setd(i, i, 1.0); return m; }
    double getd(int r, int c) const { return depth() == CV_64F ? at<double>(r, c) : (double)at<float>(r, c); }
    void setd(int r, int c, double v) { if (depth() == CV_64F) at<double>(r, c) = v; else at<float>(r, c) = (float)v; }
    template <class T> T& at(int i) const { return cols == 1 ? const_cast<Mat*>(this)->at<T>(i, 0) : const_cast<Mat*>(this)->at<T>(i / cols, i % cols); }
    uchar* data() const { return data_; }
    static MatZeros zeros(Size sz, int type) { return MatZeros{sz.height, sz.width, type, 0}; }      // Size(width, height)
    explicit Mat(const Point3f& p) { create(3, 1, CV_32F); at<float>(0, 0) = p.x; at<float>(1, 0) = p.y; at<float>(2, 0) = p.z; }
    void resize(int nrows) { Mat m(nrows, cols, type_); for (int r = 0; r < nrows; ++r) std::memset(m.ptr(r), 0, m.step); for (int r = 0; r < nrows && r < rows; ++r) std::memcpy(m.ptr(r), ptr(r), step < m.step ? step : m.step); *this = m; }
    struct SizeTag {};                                          // only streamed in a log line of Tracking's constructor
    inline static const SizeTag size{};
    int channels() const { return (type_ >> 3) + 1; }
private:
    void copyRows(Mat& dst) const { for (int r = 0; r < rows; ++r) std::memcpy(dst.data_ + (size_t)r * dst.step, data_ + (size_t)r * step, (size_t)cols * esz(type_)); }
    static size_t esz(int type) { static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return d[type & 7] * (size_t)((type >> 3) + 1); }
    int type_ = 0;
    std::shared_ptr<uchar> owner_;
    uchar* data_ = nullptr;
};

}  // namespace cv

#include "mat_algebra.hpp"

// cv::FileStorage / FileNode: named by the YAML save / load members of Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h (:1453-1632), which the
// reference never calls (it loads ORBvoc.txt through loadFromTextFile, :1338-1434); inert placeholders so that the class template compiles.
#include <map>
#include <string>
namespace cv {
struct FileNode {
    double value = 0;                                       // set when the node comes from the settings table below
    FileNode() {}
    explicit FileNode(double v) : value(v) {}
    FileNode operator[](const std::string&) const { return FileNode(); }
    FileNode operator[](const char*) const { return FileNode(); }
    FileNode operator[](int) const { return FileNode(); }
    size_t size() const { return 0; }
    operator int() const { return (int)value; }
    operator double() const { return value; }
    operator float() const { return (float)value; }
    operator std::string() const { return std::string(); }
};
struct FileStorage {
    enum { READ = 0, WRITE = 1 };
    FileStorage() {}
    FileStorage(const std::string& name, int) : opened_(name == "pslam-settings-table") {}
    bool isOpened() const { return opened_; }
    void release() { opened_ = false; }
    // No YAML parser here: a storage opened under the name "pslam-settings-table" reads numeric keys from this process-wide table, which the
    // optimiser driver fills with the values of the reference's Examples/RGB-D/*.yaml (Plane.AngleInfo ...) before Config::SetParameterFile.
    static std::map<std::string, double>& table() { static std::map<std::string, double> t; return t; }
    FileNode operator[](const std::string& key) const { auto it = table().find(key); return it == table().end() || !opened_ ? FileNode() : FileNode(it->second); }
    bool opened_ = false;
    template <class T> FileStorage& operator<<(const T&) { return *this; }
};
}  // namespace cv

// ---- what src/ORBextractor.cc needs on top of the containers: array proxies and the six image primitives.  The primitives are the
// oracle's restatements (oracle/cvprims.cc), each pinned bit-for-bit against cv2 4.13 (tests/test_oracle_cvprims.py); everything
// else the extractor computes - cell grid, threshold fallback, quadtree distribution, orientation, steered BRIEF, scaling - is the
// reference's own code.
#include "../../../cvprims.h"

#define CV_PI 3.1415926535897932384626433832795

namespace cv {

enum { BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16, INTER_LINEAR = 1 };

inline int cvRound(double v) { return oracle::cv_round(v); }
inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }
inline float fastAtan2(float y, float x) { return oracle::fast_atan2_deg(y, x); }

class _InputArray {
public:
    _InputArray(const Mat& m) : m_(const_cast<Mat*>(&m)) {}
    Mat getMat() const { return *m_; }
    bool empty() const { return m_->empty(); }
protected:
    Mat* m_;
};
class _OutputArray : public _InputArray {
public:
    _OutputArray(Mat& m) : _InputArray(m) {}
    void create(int r, int c, int type) const { m_->create(r, c, type); }
    void release() const { m_->release(); }
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;

// cv::FAST(image, keypoints, threshold, nonmaxSuppression = true), TYPE_9_16
inline void FAST(const Mat& img, std::vector<KeyPoint>& kps, int threshold, bool /*nonmax*/ = true) {
    std::vector<oracle::FastKp> out;
    oracle::fast_detect(oracle::Img8{img.ptr(0), img.cols, img.rows, (int)img.step}, threshold, out);
    kps.clear();
    for (const oracle::FastKp& k : out) kps.push_back(KeyPoint((float)k.x, (float)k.y, 7.f, -1.f, (float)k.score));
}
inline void resize(const Mat& src, Mat& dst, Size sz, double = 0, double = 0, int = INTER_LINEAR) {
    dst.create(sz.height, sz.width, src.type());          // no-op for the pre-sized pyramid ROI: the result lands inside the bordered buffer
    std::vector<uint8_t> tmp((size_t)sz.width * sz.height);
    oracle::resize_linear_u8(oracle::Img8{src.ptr(0), src.cols, src.rows, (int)src.step}, tmp.data(), sz.width, sz.height);
    for (int r = 0; r < sz.height; ++r) std::memcpy(dst.ptr(r), &tmp[(size_t)r * sz.width], sz.width);
}
inline void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int /*BORDER_REFLECT_101 [+ BORDER_ISOLATED]*/) {
    (void)bottom; (void)left; (void)right;                // the extractor passes one width on all four sides
    dst.create(src.rows + 2 * top, src.cols + 2 * top, src.type());
    std::vector<uint8_t> tmp((size_t)dst.rows * dst.cols);
    oracle::copy_make_border_reflect101(oracle::Img8{src.ptr(0), src.cols, src.rows, (int)src.step}, tmp.data(), top);
    for (int r = 0; r < dst.rows; ++r) std::memcpy(dst.ptr(r), &tmp[(size_t)r * dst.cols], dst.cols);
}
inline void GaussianBlur(const Mat& src, Mat& dst, Size /*7x7*/, double /*2*/, double /*2*/, int /*BORDER_REFLECT_101*/) {
    std::vector<uint8_t> tmp((size_t)src.rows * src.cols);
    oracle::gaussian_blur_7x7_s2_u8(oracle::Img8{src.ptr(0), src.cols, src.rows, (int)src.step}, tmp.data());
    dst.create(src.rows, src.cols, src.type());
    for (int r = 0; r < src.rows; ++r) std::memcpy(dst.ptr(r), &tmp[(size_t)r * src.cols], src.cols);
}
struct KeyPointsFilter {                                   // only referenced by the dead ComputeKeyPointsOld (src/ORBextractor.cc:855-1032)
    static void retainBest(std::vector<KeyPoint>& k, int n) {
        std::stable_sort(k.begin(), k.end(), [](const KeyPoint& a, const KeyPoint& b) { return a.response > b.response; });
        if ((int)k.size() > n) k.resize(n);
    }
};

}  // namespace cv
using cv::cvRound; using cv::cvFloor; using cv::cvCeil;
