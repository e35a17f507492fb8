// TEST INFRASTRUCTURE ONLY - CPU oracle (see cvprims.h header note).
// Restatement of the plane post-processing of Frame::ComputePlanes (src/Frame.cc:647-753) and Frame::MaxPointDistanceFromPlane (:755-813), i.e. of the three
// PCL algorithms the reference calls (PCL 1.7 / 1.9, NOT in /root/reference and not installed here):
//   pcl::VoxelGrid<PointXYZRGB> (leaf 0.1 m)                       src/Frame.cc:674-679   -> mvPlanePoints (one centroid per occupied voxel, ascending voxel index)
//   pcl::SACSegmentation (SACMODEL_PLANE, SAC_RANSAC, optimise)    src/Frame.cc:777-786   -> mvPlaneCoefficients (RANSAC with PCL's mt19937(12345) sample
//                                                                                           shuffling, inlier least-squares refit via pcl::eigen33, sign kept)
//   pcl::IntegralImageNormalEstimation (AVERAGE_3D_GRADIENT, max depth change 0.05, smoothing 10)   src/Frame.cc:715-728   -> vSurfaceNormal
// PARITY UNPINNED: the upstream sources are absent; this follows PCL's published algorithms and the structure of its implementation as recalled (voxel index
// arithmetic, sample shuffling, the closed-form eigen33, the chamfer distance map and the integral images of 3-D gradients with their border policy).  Where
// PCL leaves an order to std::sort (points of one voxel) the original point order is used.  The CUDA path is held to this restatement (bit-exact voxel sets,
// counts and normals; plane coefficients to 1e-6: the closed-form eigen-solver goes through float atan2 / cos / sin of two different libms).
#pragma once
#include <cstdint>
#include <vector>

#include "peac.h"

namespace oracle {

struct PlanePostParams { float fx, fy, cx, cy, scale; double dist_th; };       // Plane.DistanceThreshold (TUM3.yaml: 0.05)

struct PostPlane {
    int src;                          // index of the PEAC plane it comes from (planes failing the distance check are dropped)
    float coef[4];                    // mvPlaneCoefficients[i]
    std::vector<float> points;        // mvPlanePoints[i]: xyz per voxel centroid
    int n_inliers, n_iterations;      // RANSAC bookkeeping (for the tests)
};

// planes / membership: the PEAC result on the same depth image
void compute_planes_post(const uint16_t* depth, int w, int h, const PlanePostParams& prm, const PeacResult& peac, std::vector<PostPlane>& out);

struct SurfaceNormal { float normal[3]; float cam[3]; float frame_xy[2]; };
// vSurfaceNormal: every 2nd row / column (odd ones) of the normals of the 3x sub-sampled organised cloud; NaN normals at the borders are kept like the reference
void surface_normals(const uint16_t* depth, int w, int h, const PlanePostParams& prm, std::vector<SurfaceNormal>& out);

// MapPlane::UpdateCoefficientsAndPoints() / (const Frame&, int id)   src/MapPlane.cc:298-365: the plane clouds of the observations brought into one frame by
// pcl::transformPointCloud (double 4x4 applied to float points, result rounded to float), concatenated and passed through the same VoxelGrid (leaf 0.1 m);
// the result becomes MapPlane::mvPlanePoints.  (The SACSegmentation call that follows in the reference writes into locals that are never read.)
// clouds: xyz triples; T: one row-major 4x4 per cloud.  PARITY UNPINNED like the rest of this file (PCL absent).
void map_plane_update(const std::vector<std::vector<float>>& clouds, const std::vector<const double*>& T, std::vector<float>& out_points);

// boost::mt19937(12345) through boost::uniform_int<>(0, INT_MAX) as pcl::SampleConsensusModel::rnd() draws it
struct PclRng {
    uint32_t mt[624]; int idx;
    explicit PclRng(uint32_t seed = 12345u);
    uint32_t next_u32();
    int rnd() { return (int)(next_u32() >> 1); }
};

}  // namespace oracle
