// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the few OpenCV container types the reference's PEAC headers and
// src/PlaneExtractor.cpp touch (cv::Mat as a typed byte buffer with ROI views, Vec3b, Range, tick counter), so that those
// reference sources compile UNMODIFIED from /root/reference in a container without OpenCV headers (oracle/Makefile target _ref).
// No image arithmetic lives here: every number the plane extractor produces comes from the reference's own code.
#pragma once
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_8U 0
#define CV_16U 2
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn) - 1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_16UC1 CV_MAKETYPE(CV_16U, 1)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)
#define CV_32FC1 CV_MAKETYPE(CV_32F, 1)

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
struct Point3f { float x = 0, y = 0, z = 0; };
struct Point2i { int x = 0, y = 0; };
struct Range { int start, end; Range(int s, int e) : start(s), end(e) {} };

inline int64_t getTickCount() { return (int64_t)std::chrono::steady_clock::now().time_since_epoch().count(); }
inline double getTickFrequency() { return (double)std::chrono::steady_clock::period::den / std::chrono::steady_clock::period::num; }

class Mat {
public:
    int rows = 0, cols = 0;
    Mat() {}
    Mat(int r, int c, int type) { create(r, c, type); }
    Mat(int r, int c, int type, void* ext) : rows(r), cols(c), type_(type), step_((size_t)c * esz(type)), data_((uchar*)ext) {}     // external data, not owned
    static Mat ones(int r, int c, int type) { Mat m(r, c, type); std::memset(m.data_, 0, (size_t)r * m.step_); m.setTo(1); return m; }
    void create(int r, int c, int type) {
        if (data_ && r == rows && c == cols && type == type_) return;
        rows = r; cols = c; type_ = type; step_ = (size_t)c * esz(type);
        owner_ = std::shared_ptr<uchar>(new uchar[(size_t)r * step_ + 8], std::default_delete<uchar[]>());
        data_ = owner_.get();
    }
    void release() { owner_.reset(); data_ = nullptr; rows = cols = 0; }
    bool empty() const { return data_ == nullptr || rows == 0 || cols == 0; }
    int depth() const { return type_ & 7; }
    int type() const { return type_; }
    template <class T> T& at(int r, int c) { return *(T*)(data_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <class T> const T& at(int r, int c) const { return *(const T*)(data_ + (size_t)r * step_ + (size_t)c * sizeof(T)); }
    template <class T> T& at(int i) { return at<T>(i / cols, i % cols); }                 // continuous matrices only (all uses here)
    Mat operator()(const Range& rr, const Range& cr) const {
        Mat m; m.rows = rr.end - rr.start; m.cols = cr.end - cr.start; m.type_ = type_; m.step_ = step_; m.owner_ = owner_;
        m.data_ = data_ + (size_t)rr.start * step_ + (size_t)cr.start * esz(type_);
        return m;
    }
    Mat& setTo(int value) {
        for (int r = 0; r < rows; ++r)
            for (int c = 0; c < cols; ++c) {
                uchar* p = data_ + (size_t)r * step_ + (size_t)c * esz(type_);
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
            for (int c = 0; c < cols; ++c) std::memcpy(data_ + (size_t)r * step_ + (size_t)c * 3, value.v, 3);
        return *this;
    }
    uchar* data() const { return data_; }
private:
    int channels() const { return (type_ >> 3) + 1; }
    static size_t esz(int type) { static const size_t d[8] = {1, 1, 2, 2, 4, 4, 8, 2}; return d[type & 7] * (size_t)((type >> 3) + 1); }
    int type_ = 0;
    size_t step_ = 0;
    std::shared_ptr<uchar> owner_;
    uchar* data_ = nullptr;
};

}  // namespace cv
