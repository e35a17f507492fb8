// TEST INFRASTRUCTURE ONLY.  Optimizer::LocalBundleAdjustment (src/Optimizer.cc:1853-2678) run by the REFERENCE's own g2o (BlockSolver_6_3 with
// Schur complement, Levenberg-Marquardt, Huber kernels), its binary edges (EdgeSE3ProjectXYZ, EdgeStereoSE3ProjectXYZ, EdgeLineProjectXYZ, EdgePlane,
// EdgeVerticalPlane, EdgeParallelPlane) and vertices (VertexSE3Expmap, VertexSBAPointXYZ, VertexPlane), compiled unmodified against the stand-ins of
// oracle/ref/shims/ (Eigen = mini_eigen.hpp; LinearSolverEigen = a dense stand-in, see shims/Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h).
// The function itself walks KeyFrame / MapPoint / MapLine / MapPlane objects; its graph construction (:1985-2350), the two optimisations with the
// chi-square gating in between (:2354-2467) and the erase lists (:2470-2580) are restated here on the plain-array problem of the C ABI, in the
// reference's order, including its quirks: every line edge hangs on the key frame the caller names (the reference uses the outer pKF, :2169-2201),
// vertical / parallel plane edges carry angleInfo (:2274-2276).  Part of oracle/_ref/libpose_ref.so; pins oracle/lba.cc (tests/test_oracle_lba_ref.py).
#include <cmath>
#include <cstdint>
#include <vector>

#include "Converter.h"
#include "EdgeLine.h"
#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
#include "g2oAddition/EdgeParallelPlane.h"
#include "g2oAddition/EdgePlane.h"
#include "g2oAddition/EdgeVerticalPlane.h"
#include "pslam_abi.h"

using namespace Planar_SLAM;

static cv::Mat lba_coeff4(const float* v) { cv::Mat m(4, 1, CV_32F); for (int i = 0; i < 4; ++i) m.at<float>(i, 0) = v[i]; return m; }

extern "C" int ref_local_bundle_adjustment(const pslam_lba_problem* P, pslam_lba_result* R) {
    g2o::SparseOptimizer optimizer;
    g2o::BlockSolver_6_3::LinearSolverType* linearSolver = new g2o::LinearSolverEigen<g2o::BlockSolver_6_3::PoseMatrixType>();
    g2o::BlockSolver_6_3* solver_ptr = new g2o::BlockSolver_6_3(linearSolver);
    g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    optimizer.setAlgorithm(solver);
    unsigned long maxKFid = 0;
    for (int k = 0; k < P->n_kf; ++k) {
        cv::Mat T(4, 4, CV_32F);
        for (int i = 0; i < 16; ++i) T.at<float>(i / 4, i % 4) = P->kf_Tcw[16 * k + i];
        g2o::VertexSE3Expmap* vSE3 = new g2o::VertexSE3Expmap();
        vSE3->setEstimate(Converter::toSE3Quat(T));
        vSE3->setId(k);
        vSE3->setFixed(P->kf_fixed[k] != 0);
        optimizer.addVertex(vSE3);
        if ((unsigned long)k > maxKFid) maxKFid = k;
    }
    const float thHuberMono = sqrt(5.991), thHuberStereo = sqrt(7.815);
    long unsigned int maxMapPointId = maxKFid;
    std::vector<g2o::VertexSBAPointXYZ*> vPoints(P->n_points);
    for (int i = 0; i < P->n_points; ++i) {
        g2o::VertexSBAPointXYZ* vPoint = new g2o::VertexSBAPointXYZ();
        cv::Mat X(3, 1, CV_32F);
        for (int c = 0; c < 3; ++c) X.at<float>(c) = P->pt_Xw[3 * i + c];
        vPoint->setEstimate(Converter::toVector3d(X));
        const int id = i + maxKFid + 1;
        vPoint->setId(id);
        vPoint->setMarginalized(true);
        optimizer.addVertex(vPoint);
        if ((unsigned long)id > maxMapPointId) maxMapPointId = id;
        vPoints[i] = vPoint;
    }
    std::vector<g2o::EdgeSE3ProjectXYZ*> eMono(P->n_pt_obs, nullptr);
    std::vector<g2o::EdgeStereoSE3ProjectXYZ*> eStereo(P->n_pt_obs, nullptr);
    for (int j = 0; j < P->n_pt_obs; ++j) {
        const int k = P->pt_obs_kf[j], id = P->pt_obs_pt[j] + maxKFid + 1;
        const float* K = P->kf_K + 5 * k;
        const float invSigma2 = P->pt_obs_inv_sigma2[j];
        if (P->pt_obs_uvr[3 * j + 2] < 0) {
            Eigen::Matrix<double, 2, 1> obs;
            obs << P->pt_obs_uvr[3 * j], P->pt_obs_uvr[3 * j + 1];
            g2o::EdgeSE3ProjectXYZ* e = new g2o::EdgeSE3ProjectXYZ();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(id)));
            e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(k)));
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix2d::Identity() * invSigma2);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(thHuberMono);
            e->fx = K[0]; e->fy = K[1]; e->cx = K[2]; e->cy = K[3];
            optimizer.addEdge(e);
            eMono[j] = e;
        } else {
            Eigen::Matrix<double, 3, 1> obs;
            obs << P->pt_obs_uvr[3 * j], P->pt_obs_uvr[3 * j + 1], P->pt_obs_uvr[3 * j + 2];
            g2o::EdgeStereoSE3ProjectXYZ* e = new g2o::EdgeStereoSE3ProjectXYZ();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(id)));
            e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(k)));
            e->setMeasurement(obs);
            Eigen::Matrix3d Info = Eigen::Matrix3d::Identity() * invSigma2;
            e->setInformation(Info);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(thHuberStereo);
            e->fx = K[0]; e->fy = K[1]; e->cx = K[2]; e->cy = K[3]; e->bf = K[4];
            optimizer.addEdge(e);
            eStereo[j] = e;
        }
    }
    long unsigned int maxMapLineId = maxMapPointId;
    std::vector<g2o::VertexSBAPointXYZ*> vLs(P->n_lines), vLe(P->n_lines);
    for (int i = 0; i < P->n_lines; ++i) {
        g2o::VertexSBAPointXYZ* vs = new g2o::VertexSBAPointXYZ();
        vs->setEstimate(Eigen::Vector3d(P->line_Xw[6 * i], P->line_Xw[6 * i + 1], P->line_Xw[6 * i + 2]));
        const int id1 = (2 * i) + 1 + maxMapPointId;
        vs->setId(id1);
        vs->setMarginalized(true);
        optimizer.addVertex(vs);
        g2o::VertexSBAPointXYZ* ve = new g2o::VertexSBAPointXYZ();
        ve->setEstimate(Eigen::Vector3d(P->line_Xw[6 * i + 3], P->line_Xw[6 * i + 4], P->line_Xw[6 * i + 5]));
        const int id2 = (2 * (i + 1)) + maxMapPointId;
        ve->setId(id2);
        ve->setMarginalized(true);
        optimizer.addVertex(ve);
        if ((unsigned long)id2 > maxMapLineId) maxMapLineId = id2;
        vLs[i] = vs; vLe[i] = ve;
    }
    std::vector<EdgeLineProjectXYZ*> eLs(P->n_line_obs), eLe(P->n_line_obs);
    for (int j = 0; j < P->n_line_obs; ++j) {
        const int k = P->line_obs_kf[j], li = P->line_obs_line[j];
        const float* K = P->kf_K + 5 * k;
        Eigen::Vector3d lineObs(P->line_obs_l[3 * j], P->line_obs_l[3 * j + 1], P->line_obs_l[3 * j + 2]);
        for (int s = 0; s < 2; ++s) {
            EdgeLineProjectXYZ* e = new EdgeLineProjectXYZ();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(s == 0 ? (2 * li) + 1 + maxMapPointId : (2 * (li + 1)) + maxMapPointId)));
            e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(k)));
            e->setMeasurement(lineObs);
            e->setInformation(Eigen::Matrix3d::Identity());
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(thHuberStereo);
            e->fx = K[0]; e->fy = K[1]; e->cx = K[2]; e->cy = K[3];
            optimizer.addEdge(e);
            (s == 0 ? eLs : eLe)[j] = e;
        }
    }
    double angleInfo = P->angle_info; angleInfo = 3282.8 / (angleInfo * angleInfo);
    double disInfo = P->dist_info; disInfo = disInfo * disInfo;
    const double planeChi = P->plane_chi, VPplaneChi = P->vp_chi;
    const float deltaPlane = sqrt(planeChi), VPdeltaPlane = sqrt(VPplaneChi);
    std::vector<g2o::VertexPlane*> vPlanes(P->n_planes);
    Eigen::Matrix3d Info;
    Info << angleInfo, 0, 0, 0, angleInfo, 0, 0, 0, disInfo;
    Eigen::Matrix2d VPInfo;
    VPInfo << angleInfo, 0, 0, angleInfo;
    for (int i = 0; i < P->n_planes; ++i) {
        g2o::VertexPlane* vPlane = new g2o::VertexPlane();
        vPlane->setEstimate(Converter::toPlane3D(lba_coeff4(P->plane_Xw + 4 * i)));
        vPlane->setId(i + maxMapLineId + 1);
        vPlane->setMarginalized(true);
        optimizer.addVertex(vPlane);
        vPlanes[i] = vPlane;
    }
    std::vector<g2o::EdgePlane*> ePl(P->n_plane_obs[0]);
    std::vector<g2o::EdgeVerticalPlane*> eVer(P->n_plane_obs[1]);
    std::vector<g2o::EdgeParallelPlane*> ePar(P->n_plane_obs[2]);
    // the reference adds, per plane, its plane / vertical / parallel observations; the observation arrays are plane-major, so walking them family by
    // family visits the same edges (their order only matters to the rounding of the Hessian sums)
    for (int j = 0; j < P->n_plane_obs[0]; ++j) {
        g2o::EdgePlane* e = new g2o::EdgePlane();
        e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_kf[0][j])));
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_plane[0][j] + maxMapLineId + 1)));
        e->setMeasurement(Converter::toPlane3D(lba_coeff4(P->plane_obs_meas[0] + 4 * j)));
        e->setInformation(Info);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(deltaPlane);
        optimizer.addEdge(e);
        ePl[j] = e;
    }
    for (int j = 0; j < P->n_plane_obs[1]; ++j) {
        g2o::EdgeVerticalPlane* e = new g2o::EdgeVerticalPlane();
        e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_kf[1][j])));
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_plane[1][j] + maxMapLineId + 1)));
        e->setMeasurement(Converter::toPlane3D(lba_coeff4(P->plane_obs_meas[1] + 4 * j)));
        e->setInformation(VPInfo);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(VPdeltaPlane);
        optimizer.addEdge(e);
        eVer[j] = e;
    }
    for (int j = 0; j < P->n_plane_obs[2]; ++j) {
        g2o::EdgeParallelPlane* e = new g2o::EdgeParallelPlane();
        e->setVertex(1, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_kf[2][j])));
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(P->plane_obs_plane[2][j] + maxMapLineId + 1)));
        e->setMeasurement(Converter::toPlane3D(lba_coeff4(P->plane_obs_meas[2] + 4 * j)));
        e->setInformation(VPInfo);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(VPdeltaPlane);
        optimizer.addEdge(e);
        ePar[j] = e;
    }
    optimizer.initializeOptimization();
    R->iterations[0] = optimizer.optimize(5);
    for (int j = 0; j < P->n_pt_obs; ++j) {
        if (eMono[j]) { g2o::EdgeSE3ProjectXYZ* e = eMono[j]; if (e->chi2() > 5.991 || !e->isDepthPositive()) e->setLevel(1); e->setRobustKernel(0); }
    }
    for (int j = 0; j < P->n_pt_obs; ++j) {
        if (eStereo[j]) { g2o::EdgeStereoSE3ProjectXYZ* e = eStereo[j]; if (e->chi2() > 7.815 || !e->isDepthPositive()) e->setLevel(1); e->setRobustKernel(0); }
    }
    for (int j = 0; j < P->n_line_obs; ++j) {
        if (eLs[j]->chi2() > 7.815 || eLe[j]->chi2() > 7.815) { eLs[j]->setLevel(1); eLe[j]->setLevel(1); }
        eLs[j]->setRobustKernel(0); eLe[j]->setRobustKernel(0);
    }
    for (auto* e : ePl) { if (e->chi2() > planeChi) e->setLevel(1); e->setRobustKernel(0); }
    for (auto* e : eVer) { if (e->chi2() > VPplaneChi) e->setLevel(1); e->setRobustKernel(0); }
    for (auto* e : ePar) { if (e->chi2() > VPplaneChi) e->setLevel(1); e->setRobustKernel(0); }
    optimizer.initializeOptimization(0);
    R->iterations[1] = optimizer.optimize(10);
    for (int j = 0; j < P->n_pt_obs; ++j) {
        if (eMono[j]) R->erase_pt[j] = (eMono[j]->chi2() > 5.991 || !eMono[j]->isDepthPositive()) ? 1 : 0;
        else R->erase_pt[j] = (eStereo[j]->chi2() > 7.815 || !eStereo[j]->isDepthPositive()) ? 1 : 0;
    }
    for (int j = 0; j < P->n_line_obs; ++j) R->erase_line[j] = (eLs[j]->chi2() > 7.815 || eLe[j]->chi2() > 7.815) ? 1 : 0;
    for (int j = 0; j < P->n_plane_obs[0]; ++j) R->erase_plane[0][j] = ePl[j]->chi2() > planeChi ? 1 : 0;
    for (int j = 0; j < P->n_plane_obs[1]; ++j) R->erase_plane[1][j] = eVer[j]->chi2() > VPplaneChi ? 1 : 0;
    for (int j = 0; j < P->n_plane_obs[2]; ++j) R->erase_plane[2][j] = ePar[j]->chi2() > VPplaneChi ? 1 : 0;
    for (int k = 0; k < P->n_kf; ++k) {
        g2o::VertexSE3Expmap* v = static_cast<g2o::VertexSE3Expmap*>(optimizer.vertex(k));
        const Eigen::Matrix<double, 4, 4> T = v->estimate().to_homogeneous_matrix();
        for (int i = 0; i < 4; ++i) for (int c = 0; c < 4; ++c) { R->kf_Tcw_d[16 * k + 4 * i + c] = T(i, c); R->kf_Tcw[16 * k + 4 * i + c] = (float)T(i, c); }
    }
    for (int i = 0; i < P->n_points; ++i) for (int c = 0; c < 3; ++c) { R->pt_Xw_d[3 * i + c] = vPoints[i]->estimate()[c]; R->pt_Xw[3 * i + c] = (float)vPoints[i]->estimate()[c]; }
    for (int i = 0; i < P->n_lines; ++i)
        for (int c = 0; c < 3; ++c) {
            R->line_Xw_d[6 * i + c] = vLs[i]->estimate()[c]; R->line_Xw_d[6 * i + 3 + c] = vLe[i]->estimate()[c];
            R->line_Xw[6 * i + c] = (float)vLs[i]->estimate()[c]; R->line_Xw[6 * i + 3 + c] = (float)vLe[i]->estimate()[c];
        }
    for (int i = 0; i < P->n_planes; ++i) {
        const Eigen::Matrix<double, 4, 1> v = vPlanes[i]->estimate().toVector();
        for (int c = 0; c < 4; ++c) { R->plane_Xw_d[4 * i + c] = v[c]; R->plane_Xw[4 * i + c] = (float)v[c]; }
    }
    return 0;
}
