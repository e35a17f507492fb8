// TEST INFRASTRUCTURE ONLY.  Named by g2oAddition/Plane3D.h and EdgePlane.h for point-cloud members the optimiser path never touches.
#pragma once
#include <memory>
#include <vector>
namespace pcl {
#ifndef PSLAM_PCL_STUB
#define PSLAM_PCL_STUB
struct PointXYZRGB { float x = 0, y = 0, z = 0; };
template <class T> struct PointCloud { typedef std::shared_ptr<PointCloud<T>> Ptr; std::vector<T> points; };
#endif
}
