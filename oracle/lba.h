// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1853-2678) from the point where the graph is
// filled (:1971) to the point where the optimised estimates and the erase lists are handed back (:2462-2677), on top of
// a restatement of the vendored g2o pieces it drives:
//   graph set-up            src/Optimizer.cc:1992-2358 (vertices, edges, information, Huber deltas)
//   index mapping           Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:166-190,206-267 (active = has a level-0 edge; poses
//                           first, then marginalised landmarks, both in id order)
//   BlockSolver_6_3         Thirdparty/g2o/g2o/core/block_solver.hpp:140-290 (structure), :344-482 (Schur complement solve),
//                           :496-552 (buildSystem), :556-600 (setLambda / restoreDiagonal)
//   binary quadratic form   Thirdparty/g2o/g2o/core/base_binary_edge.hpp:55-118, numeric Jacobian :128-203 (error restored)
//   LM                      Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   linear solver           Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h:94-125 (Eigen SimplicialLDLT + AMD ordering;
//                           restated as a dense unpivoted LDL^T of the reduced system - same solution up to rounding)
//   edges                   EdgeSE3ProjectXYZ / EdgeStereoSE3ProjectXYZ types_six_dof_expmap.cpp:103-232 ; EdgeLineProjectXYZ
//                           include/EdgeLine.h:53-153 ; EdgePlane g2oAddition/EdgePlane.h:24-126 ; EdgeVerticalPlane /
//                           EdgeParallelPlane g2oAddition/Edge{Vertical,Parallel}Plane.h:21-108 ; VertexPlane::oplusImpl
//                           g2oAddition/VertexPlane.h:24-27 ; VertexSBAPointXYZ::oplusImpl types_sba.h:52-56
// The Map / KeyFrame pointer graph (which key frames are local / fixed, which observations exist; :1853-1969) stays on
// the host: the caller passes plain arrays.  Quirks kept: line edges gate the two endpoints jointly (:2382-2396); the
// vertical / parallel information uses angleInfo (:2274-2276).  Quirk left to the caller: the reference
// attaches every line edge to the *current* key frame's vertex, intrinsics and line function (:2169-2201) - pass that
// key frame's index in line_obs_kf to reproduce it.
// Pinned: the reference ships no tests for this path (its only call site is commented out, src/LocalMapping.cc:68) and
// Optimizer.cc cannot be compiled here, but its g2o, vertices and edges can (oracle/ref/lba_driver.cc -> oracle/_ref/libpose_ref.so):
// identical erase lists and iteration counts, poses within 2e-6 rad / 5e-6 m (tests/test_oracle_lba_ref.py, tests/golden/lba_reference.npz).
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

struct LbaProblem {
    int n_kf = 0;
    const float* kf_Tcw = nullptr;        // [n_kf][16] row-major KeyFrame::GetPose()
    const uint8_t* kf_fixed = nullptr;    // [n_kf]
    const float* kf_K = nullptr;          // [n_kf][5] fx fy cx cy bf
    int n_points = 0;
    const float* pt_Xw = nullptr;         // [n_points][3]
    int n_pt_obs = 0;
    const int32_t* pt_obs_kf = nullptr;
    const int32_t* pt_obs_pt = nullptr;
    const float* pt_obs_uvr = nullptr;    // [n][3] u, v, uR (< 0: monocular)
    const float* pt_obs_inv_sigma2 = nullptr;
    int n_lines = 0;
    const double* line_Xw = nullptr;      // [n_lines][6]
    int n_line_obs = 0;
    const int32_t* line_obs_kf = nullptr;
    const int32_t* line_obs_line = nullptr;
    const double* line_obs_l = nullptr;   // [n][3]
    int n_planes = 0;
    const float* plane_Xw = nullptr;      // [n_planes][4]
    int n_plane_obs[3] = {0, 0, 0};       // 0 plane, 1 vertical, 2 parallel
    const int32_t* plane_obs_kf[3] = {nullptr, nullptr, nullptr};
    const int32_t* plane_obs_plane[3] = {nullptr, nullptr, nullptr};
    const float* plane_obs_meas[3] = {nullptr, nullptr, nullptr};   // [n][4]
    double angle_info = 0.5, dist_info = 50, plane_chi = 100, vp_chi = 50;
};

struct LbaResult {
    std::vector<float> kf_Tcw;      // [n_kf][16]
    std::vector<double> kf_Tcw_d;
    std::vector<float> pt_Xw;       // [n_points][3]
    std::vector<double> pt_Xw_d;
    std::vector<double> line_Xw;    // [n_lines][6], float-rounded like Converter::toCvMat
    std::vector<double> line_Xw_d;
    std::vector<float> plane_Xw;    // [n_planes][4]
    std::vector<double> plane_Xw_d;
    std::vector<uint8_t> erase_pt, erase_line, erase_plane[3];
    int iterations[2] = {0, 0}, trials[2] = {0, 0};
    double chi2[2] = {0, 0}, lambda[2] = {0, 0};
};

void local_bundle_adjustment(const LbaProblem& P, LbaResult& out);

}  // namespace oracle
