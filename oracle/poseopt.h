// TEST INFRASTRUCTURE ONLY — CPU oracle (see cvprims.h header note).
// Restatement of Optimizer::PoseOptimization (src/Optimizer.cc:550-1275) on top of a restatement of the vendored,
// modified g2o it drives:
//   LM control flow       Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:61-189
//   optimize() loop       Thirdparty/g2o/g2o/core/sparse_optimizer.cpp:354-419 (active errors :61-76, robust chi2 :100-114)
//   quadratic form        Thirdparty/g2o/g2o/core/base_unary_edge.hpp:43-72, numeric Jacobian :82-122
//   Huber                 Thirdparty/g2o/g2o/core/robust_kernel_impl.cpp:78-91 ; chi2 / robustInformation base_edge.h:58-61,96-102
//   dense solve           Thirdparty/g2o/g2o/solvers/linear_solver_dense.h:65-113 (Eigen LDLT; restated as unpivoted LDL^T)
//   SE3Quat               Thirdparty/g2o/g2o/types/se3quat.h (exp :227-260, operator* :98-104, map :217, normalizeRotation :286)
//   point edges           types_six_dof_expmap.h:143-171,203-231 ; .cpp:266-306,335-364 (stereo uses a float 1/z, :300)
//   line edge             include/EdgeLine.h:155-245
//   plane edges           g2oAddition/EdgePlane.h:128-224, EdgeParallelPlane.h:110-194, EdgeVerticalPlane.h:111-195,
//                         g2oAddition/Plane3D.h (normalize :175-180, operator* :186-199, rotation :76-82, ominus :127-134,
//                         ominus_ver :136-153, ominus_par :155-173) ; Converter::toSE3Quat / toPlane3D src/Converter.cc:37-45,171-180
// Pinned: the reference ships no tests for this path and Optimizer.cc cannot be compiled here, but its g2o, vertices, edges and Converter can
// (oracle/ref/pose_driver.cc -> oracle/_ref/libpose_ref.so): identical inlier counts / outlier flags, poses within 6e-7 rad / 2e-6 m
// (tests/test_oracle_pose_ref.py, tests/golden/pose_reference.npz).
#pragma once
#include <cstdint>
#include <vector>

namespace oracle {

struct PoseProblem {
    float fx, fy, cx, cy, bf;
    // matched map points: world position (float, as MapPoint::GetWorldPos), undistorted keypoint, right coordinate
    // (< 0: monocular edge), inverse level sigma^2
    int n_points = 0;
    const float* Xw = nullptr;          // [n][3]
    const float* obs = nullptr;         // [n][3] = u, v, uR
    const float* inv_sigma2 = nullptr;  // [n]
    // matched map lines: world endpoints (double, MapLine::mWorldPos) and the observed 2-D line (mvKeyLineFunctions)
    int n_lines = 0;
    const double* line_Xw = nullptr;    // [n][6]
    const double* line_obs = nullptr;   // [n][3]
    // planes: frame plane coefficients (mvPlaneCoefficients, float 4) and matched map plane (world, float 4)
    int n_planes = 0, n_par = 0, n_ver = 0;
    const float* plane_meas = nullptr; const float* plane_map = nullptr;
    const float* par_meas = nullptr;   const float* par_map = nullptr;
    const float* ver_meas = nullptr;   const float* ver_map = nullptr;
    // settings (Plane.AngleInfo, DistanceInfo, ParallelInfo, VerticalInfo, Chi, VPChi)
    double angle_info = 0.5, dist_info = 50, par_info = 0.1, ver_info = 0.1, plane_chi = 100, vp_chi = 50;
};

struct PoseTraceRound { int lm_iterations; int trials; double chi2_final; double lambda_final; int n_bad; };

struct PoseResult {
    float Tcw[16];                 // optimised pose as the reference stores it (float 4x4, row-major)
    double Tcw_d[16];              // same before the float cast
    int n_inliers = 0;             // nInitialCorrespondences - nBad, or 0 when fewer than 3 correspondences
    std::vector<uint8_t> outlier_pt, outlier_line, outlier_plane, outlier_par, outlier_ver;
    PoseTraceRound rounds[4];
    int n_rounds = 0;
};

// Tcw_in: float 4x4 row-major (Frame::mTcw)
void pose_optimization(const PoseProblem& P, const float* Tcw_in, PoseResult& out);

// Optimizer::TranslationOptimization (src/Optimizer.cc:2995-3737): rotation frozen, every map point / line endpoint /
// plane normal is pre-rotated by R_cw (float, :3019,3066,3259) and the edges use SE3Quat::mapTrans (se3quat.h:221):
//   EdgeSE3ProjectXYZOnlyTranslation / EdgeStereoSE3ProjectXYZOnlyTranslation  types_six_dof_expmap.h:173-201,233-261, .cpp:368-375,404-432,463-485
//   EdgeLineProjectXYZOnlyTranslation  include/EdgeLine.h:247-337 ;  EdgePlaneOnlyTranslation  g2oAddition/EdgePlane.h:226-315
// Differences from PoseOptimization that are reproduced: only points count as correspondences (:3137-3139, :3245-3246), the
// function returns 0 before adding plane edges when fewer than 3 points are matched (:3198-3200), no parallel / vertical
// plane edges (:3215-3220), line errors are only recomputed for flagged lines and line outliers do not enter nBad.
void translation_optimization(const PoseProblem& P, const float* Tcw_in, PoseResult& out);

}  // namespace oracle
