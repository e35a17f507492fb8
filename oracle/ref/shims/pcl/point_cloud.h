// TEST INFRASTRUCTURE ONLY.  The reference's headers and sources name PCL types for point-cloud members (plane boundary clouds, voxel filters, RANSAC
// refits, integral-image normals) that none of the pinned paths (PEAC, optimisers, matchers) executes.  These stand-ins exist so that the
// reference's translation units COMPILE; they pass EMPTY clouds through (MapPlane::UpdateCoefficientsAndPoints after a bundle adjustment refreshes the boundary
// cloud, which the drivers leave empty) and abort if a PCL algorithm would actually have to run.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <vector>
namespace pcl {
#ifndef PSLAM_PCL_STUB
#define PSLAM_PCL_STUB
#define PCL_ERROR(...) std::fprintf(stderr, __VA_ARGS__)
[[noreturn]] inline void stub_reached(const char* what) { std::fprintf(stderr, "oracle/ref/shims/pcl: %s is a compile-only stand-in\n", what); std::abort(); }
struct PointXYZRGB { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0; };
struct PointXYZRGBA { float x = 0, y = 0, z = 0; unsigned char r = 0, g = 0, b = 0, a = 0; };
struct Normal { float normal_x = 0, normal_y = 0, normal_z = 0, curvature = 0; };
template <class T> struct PointCloud {
    typedef std::shared_ptr<PointCloud<T>> Ptr;
    typedef std::shared_ptr<const PointCloud<T>> ConstPtr;
    std::vector<T> points;
    unsigned width = 0, height = 0;
    bool is_dense = true;
    size_t size() const { return points.size(); }
    bool empty() const { return points.empty(); }
    void push_back(const T& p) { points.push_back(p); }
    void clear() { points.clear(); }
    T& at(int col, int row) { return points[(size_t)row * width + col]; }
    const T& at(int col, int row) const { return points[(size_t)row * width + col]; }
    T& operator[](size_t i) { return points[i]; }
    const T& operator[](size_t i) const { return points[i]; }
    typename std::vector<T>::iterator begin() { return points.begin(); }
    typename std::vector<T>::iterator end() { return points.end(); }
    typename std::vector<T>::const_iterator begin() const { return points.begin(); }
    typename std::vector<T>::const_iterator end() const { return points.end(); }
    PointCloud& operator+=(const PointCloud& o) { points.insert(points.end(), o.points.begin(), o.points.end()); width = (unsigned)points.size(); height = 1; return *this; }
    Ptr makeShared() const { return Ptr(new PointCloud<T>(*this)); }
};
struct ModelCoefficients { typedef std::shared_ptr<ModelCoefficients> Ptr; std::vector<float> values; };
struct PointIndices { typedef std::shared_ptr<PointIndices> Ptr; std::vector<int> indices; };
enum { SACMODEL_PLANE = 0, SAC_RANSAC = 0 };
template <class T> struct VoxelGrid {
    void setLeafSize(float, float, float) {}
    template <class P> void setInputCloud(const P& c) { n_in_ = c ? c->size() : 0; }
    void filter(PointCloud<T>& out) { if (n_in_) stub_reached("pcl::VoxelGrid::filter on a non-empty cloud"); out.clear(); }     // empty in, empty out
    size_t n_in_ = 0;
};
template <class T> struct SACSegmentation {
    void setOptimizeCoefficients(bool) {}
    void setModelType(int) {}
    void setMethodType(int) {}
    void setDistanceThreshold(double) {}
    void setMaxIterations(int) {}
    template <class P> void setInputCloud(const P&) {}
    void segment(PointIndices&, ModelCoefficients&) { stub_reached("pcl::SACSegmentation::segment"); }
};
template <class T, class N> struct IntegralImageNormalEstimation {
    enum { COVARIANCE_MATRIX, AVERAGE_3D_GRADIENT, AVERAGE_DEPTH_CHANGE, SIMPLE_3D_GRADIENT };
    void setNormalEstimationMethod(int) {}
    void setMaxDepthChangeFactor(float) {}
    void setNormalSmoothingSize(float) {}
    template <class P> void setInputCloud(const P&) {}
    void compute(PointCloud<N>&) { stub_reached("pcl::IntegralImageNormalEstimation::compute"); }
};
template <class T, class M> void transformPointCloud(const PointCloud<T>& in, PointCloud<T>& out, const M&) { if (in.size()) stub_reached("pcl::transformPointCloud on a non-empty cloud"); out.clear(); }
#endif
}
