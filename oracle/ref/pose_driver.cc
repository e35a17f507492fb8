// TEST INFRASTRUCTURE ONLY.  The REFERENCE's own g2o (Thirdparty/g2o: sparse optimiser, BlockSolver_6_3, dense linear solver, Levenberg-Marquardt,
// Huber kernel, SE3Quat, VertexSE3Expmap, EdgeSE3ProjectXYZOnlyPose / EdgeStereoSE3ProjectXYZOnlyPose), its line edge (include/EdgeLine.h) and plane
// types / edges (g2oAddition/*.h) and src/Converter.cc, compiled unmodified from /root/reference against the stand-ins of oracle/ref/shims/
// (Eigen = the eager mini library there: same formulas, different rounding in reductions).  Optimizer::PoseOptimization itself (src/Optimizer.cc:550-1275)
// reads its inputs from Frame / MapPoint / MapLine / MapPlane objects, whose headers need the whole system; the graph construction and the four
// optimise-and-classify rounds are therefore restated here on the plain-array problem of the C ABI, line by line in the order of the reference.
// Built into oracle/_ref/libpose_ref.so by `make -C oracle ref`; pins oracle/poseopt.cc (tests/test_oracle_pose_ref.py).
#include <cmath>
#include <cstdint>
#include <vector>

#include "Converter.h"
#include "EdgeLine.h"
#include "Thirdparty/g2o/g2o/core/block_solver.h"
#include "Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.h"
#include "Thirdparty/g2o/g2o/core/robust_kernel_impl.h"
#include "Thirdparty/g2o/g2o/solvers/linear_solver_dense.h"
#include "Thirdparty/g2o/g2o/types/types_six_dof_expmap.h"
#include "g2oAddition/EdgeParallelPlane.h"
#include "g2oAddition/EdgePlane.h"
#include "g2oAddition/EdgeVerticalPlane.h"

using namespace Planar_SLAM;

extern "C" {
struct ref_pose_problem {                 // = orc_pose_problem / pslam_pose_problem
    float fx, fy, cx, cy, bf;
    int32_t n_points; const float* Xw; const float* obs; const float* inv_sigma2;
    int32_t n_lines; const double* line_Xw; const double* line_obs;
    int32_t n_planes, n_par, n_ver;
    const float *plane_meas, *plane_map, *par_meas, *par_map, *ver_meas, *ver_map;
    double angle_info, dist_info, par_info, ver_info, plane_chi, vp_chi;
};

static cv::Mat coeff4(const float* v) { cv::Mat m(4, 1, CV_32F); for (int i = 0; i < 4; ++i) m.at<float>(i, 0) = v[i]; return m; }

// returns nInitialCorrespondences - nBad; Tcw_d: optimised pose (row-major 4x4 double); iters: optimize() return values of the four rounds
int ref_pose_optimization(const ref_pose_problem* P, const float* Tcw_in, double* Tcw_d, uint8_t* o_pt, uint8_t* o_line, uint8_t* o_plane, uint8_t* o_par,
                          uint8_t* o_ver, int32_t* iters) {
    g2o::SparseOptimizer optimizer;
    g2o::BlockSolver_6_3::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolver_6_3::PoseMatrixType>();
    g2o::BlockSolver_6_3* solver_ptr = new g2o::BlockSolver_6_3(linearSolver);
    g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    optimizer.setAlgorithm(solver);
    cv::Mat mTcw(4, 4, CV_32F);
    for (int i = 0; i < 16; ++i) mTcw.at<float>(i / 4, i % 4) = Tcw_in[i];
    int nInitialCorrespondences = 0;
    g2o::VertexSE3Expmap* vSE3 = new g2o::VertexSE3Expmap();
    vSE3->setEstimate(Converter::toSE3Quat(mTcw));
    vSE3->setId(0);
    vSE3->setFixed(false);
    optimizer.addVertex(vSE3);
    std::vector<g2o::EdgeSE3ProjectXYZOnlyPose*> eMono; std::vector<int> iMono;
    std::vector<g2o::EdgeStereoSE3ProjectXYZOnlyPose*> eStereo; std::vector<int> iStereo;
    const float deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815);
    for (int i = 0; i < P->n_points; ++i) {
        o_pt[i] = 0;
        ++nInitialCorrespondences;
        const float invSigma2 = P->inv_sigma2[i];
        if (P->obs[3 * i + 2] < 0) {
            Eigen::Matrix<double, 2, 1> obs;
            obs << P->obs[3 * i], P->obs[3 * i + 1];
            g2o::EdgeSE3ProjectXYZOnlyPose* e = new g2o::EdgeSE3ProjectXYZOnlyPose();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix2d::Identity() * invSigma2);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(deltaMono);
            e->fx = P->fx; e->fy = P->fy; e->cx = P->cx; e->cy = P->cy;
            for (int k = 0; k < 3; ++k) e->Xw[k] = P->Xw[3 * i + k];
            optimizer.addEdge(e);
            eMono.push_back(e); iMono.push_back(i);
        } else {
            Eigen::Matrix<double, 3, 1> obs;
            obs << P->obs[3 * i], P->obs[3 * i + 1], P->obs[3 * i + 2];
            g2o::EdgeStereoSE3ProjectXYZOnlyPose* e = new g2o::EdgeStereoSE3ProjectXYZOnlyPose();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            e->setMeasurement(obs);
            Eigen::Matrix3d Info = Eigen::Matrix3d::Identity() * invSigma2;
            e->setInformation(Info);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(deltaStereo);
            e->fx = P->fx; e->fy = P->fy; e->cx = P->cx; e->cy = P->cy; e->bf = P->bf;
            for (int k = 0; k < 3; ++k) e->Xw[k] = P->Xw[3 * i + k];
            optimizer.addEdge(e);
            eStereo.push_back(e); iStereo.push_back(i);
        }
    }
    std::vector<EdgeLineProjectXYZOnlyPose*> eLs, eLe; std::vector<int> iLine;
    for (int i = 0; i < P->n_lines; ++i) {
        o_line[i] = 0;
        ++nInitialCorrespondences;
        Eigen::Vector3d line_obs(P->line_obs[3 * i], P->line_obs[3 * i + 1], P->line_obs[3 * i + 2]);
        for (int s = 0; s < 2; ++s) {
            EdgeLineProjectXYZOnlyPose* el = new EdgeLineProjectXYZOnlyPose();
            el->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            el->setMeasurement(line_obs);
            el->setInformation(Eigen::Matrix3d::Identity());
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            el->setRobustKernel(rk);
            rk->setDelta(deltaStereo);
            el->fx = P->fx; el->fy = P->fy; el->cx = P->cx; el->cy = P->cy;
            el->Xw = Eigen::Vector3d(P->line_Xw[6 * i + 3 * s], P->line_Xw[6 * i + 3 * s + 1], P->line_Xw[6 * i + 3 * s + 2]);
            optimizer.addEdge(el);
            (s == 0 ? eLs : eLe).push_back(el);
        }
        iLine.push_back(i);
    }
    double angleInfo = P->angle_info; angleInfo = 3282.8 / (angleInfo * angleInfo);
    double disInfo = P->dist_info; disInfo = disInfo * disInfo;
    double parInfo = P->par_info; parInfo = 3282.8 / (parInfo * parInfo);
    double verInfo = P->ver_info; verInfo = 3282.8 / (verInfo * verInfo);
    const double planeChi = P->plane_chi, VPplaneChi = P->vp_chi;
    const float deltaPlane = sqrt(planeChi), VPdeltaPlane = sqrt(VPplaneChi);
    std::vector<g2o::EdgePlaneOnlyPose*> ePl; std::vector<g2o::EdgeParallelPlaneOnlyPose*> ePar; std::vector<g2o::EdgeVerticalPlaneOnlyPose*> eVer;
    for (int i = 0; i < P->n_planes; ++i) {
        o_plane[i] = 0;
        ++nInitialCorrespondences;
        g2o::EdgePlaneOnlyPose* e = new g2o::EdgePlaneOnlyPose();
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
        e->setMeasurement(Converter::toPlane3D(coeff4(P->plane_meas + 4 * i)));
        Eigen::Matrix3d Info;
        Info << angleInfo, 0, 0, 0, angleInfo, 0, 0, 0, disInfo;
        e->setInformation(Info);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(deltaPlane);
        e->Xw = Converter::toPlane3D(coeff4(P->plane_map + 4 * i));
        optimizer.addEdge(e);
        ePl.push_back(e);
        e->computeError();
    }
    for (int i = 0; i < P->n_par; ++i) {
        o_par[i] = 0;
        ++nInitialCorrespondences;
        g2o::EdgeParallelPlaneOnlyPose* e = new g2o::EdgeParallelPlaneOnlyPose();
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
        e->setMeasurement(Converter::toPlane3D(coeff4(P->par_meas + 4 * i)));
        Eigen::Matrix2d Info;
        Info << parInfo, 0, 0, parInfo;
        e->setInformation(Info);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(VPdeltaPlane);
        e->Xw = Converter::toPlane3D(coeff4(P->par_map + 4 * i));
        optimizer.addEdge(e);
        ePar.push_back(e);
        e->computeError();
    }
    for (int i = 0; i < P->n_ver; ++i) {
        o_ver[i] = 0;
        ++nInitialCorrespondences;
        g2o::EdgeVerticalPlaneOnlyPose* e = new g2o::EdgeVerticalPlaneOnlyPose();
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
        e->setMeasurement(Converter::toPlane3D(coeff4(P->ver_meas + 4 * i)));
        Eigen::Matrix2d Info;
        Info << verInfo, 0, 0, verInfo;
        e->setInformation(Info);
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(VPdeltaPlane);
        e->Xw = Converter::toPlane3D(coeff4(P->ver_map + 4 * i));
        optimizer.addEdge(e);
        eVer.push_back(e);
        e->computeError();
    }
    for (int i = 0; i < 16; ++i) Tcw_d[i] = Tcw_in[i];
    for (int i = 0; i < 4; ++i) iters[i] = -1;
    if (nInitialCorrespondences < 3) return 0;
    const float chi2Mono[4] = {5.991, 5.991, 5.991, 5.991};
    const float chi2Stereo[4] = {7.815, 7.815, 7.815, 7.815};
    const int its[4] = {10, 10, 10, 10};
    int nBad = 0;
    for (size_t it = 0; it < 4; it++) {
        vSE3->setEstimate(Converter::toSE3Quat(mTcw));
        optimizer.initializeOptimization(0);
        iters[it] = optimizer.optimize(its[it]);
        nBad = 0;
        for (size_t i = 0; i < eMono.size(); i++) {
            g2o::EdgeSE3ProjectXYZOnlyPose* e = eMono[i];
            const int idx = iMono[i];
            if (o_pt[idx]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > chi2Mono[it]) { o_pt[idx] = 1; e->setLevel(1); nBad++; } else { o_pt[idx] = 0; e->setLevel(0); }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < eStereo.size(); i++) {
            g2o::EdgeStereoSE3ProjectXYZOnlyPose* e = eStereo[i];
            const int idx = iStereo[i];
            if (o_pt[idx]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > chi2Stereo[it]) { o_pt[idx] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_pt[idx] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < eLs.size(); i++) {
            EdgeLineProjectXYZOnlyPose *e1 = eLs[i], *e2 = eLe[i];
            const int idx = iLine[i];
            if (o_line[idx]) { e1->computeError(); e2->computeError(); }
            e1->computeError();
            e2->computeError();
            const float chi2_s = e1->chiline(), chi2_e = e2->chiline();
            if (chi2_s > 2 * chi2Mono[it] || chi2_e > 2 * chi2Mono[it]) { o_line[idx] = 1; e1->setLevel(1); e2->setLevel(1); nBad++; }
            else { o_line[idx] = 0; e1->setLevel(0); e2->setLevel(0); }
            if (it == 2) { e1->setRobustKernel(0); e2->setRobustKernel(0); }
        }
        for (size_t i = 0; i < ePl.size(); i++) {
            g2o::EdgePlaneOnlyPose* e = ePl[i];
            if (o_plane[i]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > planeChi) { o_plane[i] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_plane[i] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < ePar.size(); i++) {
            g2o::EdgeParallelPlaneOnlyPose* e = ePar[i];
            if (o_par[i]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > VPplaneChi) { o_par[i] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_par[i] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < eVer.size(); i++) {
            g2o::EdgeVerticalPlaneOnlyPose* e = eVer[i];
            if (o_ver[i]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > VPplaneChi) { o_ver[i] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_ver[i] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        if (optimizer.edges().size() < 10) break;
    }
    g2o::VertexSE3Expmap* vSE3_recov = static_cast<g2o::VertexSE3Expmap*>(optimizer.vertex(0));
    g2o::SE3Quat q = vSE3_recov->estimate();
    const Eigen::Matrix<double, 4, 4> T = q.to_homogeneous_matrix();
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Tcw_d[4 * i + j] = T(i, j);
    return nInitialCorrespondences - nBad;
}

// Optimizer::TranslationOptimization (src/Optimizer.cc:2995-3737) restated the same way: rotation frozen, points / line end points pre-rotated by the float
// R_cw, only point correspondences counted, early return before the plane edges when there are fewer than three, planes of the first family only.
int ref_translation_optimization(const ref_pose_problem* P, const float* Tcw_in, double* Tcw_d, uint8_t* o_pt, uint8_t* o_line, uint8_t* o_plane, int32_t* iters) {
    g2o::SparseOptimizer optimizer;
    g2o::BlockSolver_6_3::LinearSolverType* linearSolver = new g2o::LinearSolverDense<g2o::BlockSolver_6_3::PoseMatrixType>();
    g2o::BlockSolver_6_3* solver_ptr = new g2o::BlockSolver_6_3(linearSolver);
    g2o::OptimizationAlgorithmLevenberg* solver = new g2o::OptimizationAlgorithmLevenberg(solver_ptr);
    optimizer.setAlgorithm(solver);
    cv::Mat mTcw(4, 4, CV_32F);
    for (int i = 0; i < 16; ++i) mTcw.at<float>(i / 4, i % 4) = Tcw_in[i];
    int nInitialCorrespondences = 0;
    g2o::VertexSE3Expmap* vSE3 = new g2o::VertexSE3Expmap();
    vSE3->setEstimate(Converter::toSE3Quat(mTcw));
    vSE3->setId(0);
    vSE3->setFixed(false);
    optimizer.addVertex(vSE3);
    cv::Mat R_cw = mTcw.rowRange(0, 3).colRange(0, 3).clone();
    std::vector<g2o::EdgeSE3ProjectXYZOnlyTranslation*> eMono; std::vector<int> iMono;
    std::vector<g2o::EdgeStereoSE3ProjectXYZOnlyTranslation*> eStereo; std::vector<int> iStereo;
    const float deltaMono = sqrt(5.991), deltaStereo = sqrt(7.815);
    auto f3 = [](double a, double b, double c) { cv::Mat m(3, 1, CV_32F); m.at<float>(0) = (float)a; m.at<float>(1) = (float)b; m.at<float>(2) = (float)c; return m; };
    for (int i = 0; i < P->n_points; ++i) {
        o_pt[i] = 0;
        ++nInitialCorrespondences;
        const float invSigma2 = P->inv_sigma2[i];
        cv::Mat Xw = f3(P->Xw[3 * i], P->Xw[3 * i + 1], P->Xw[3 * i + 2]);
        cv::Mat Xc = R_cw * Xw;
        if (P->obs[3 * i + 2] < 0) {
            Eigen::Matrix<double, 2, 1> obs;
            obs << P->obs[3 * i], P->obs[3 * i + 1];
            g2o::EdgeSE3ProjectXYZOnlyTranslation* e = new g2o::EdgeSE3ProjectXYZOnlyTranslation();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            e->setMeasurement(obs);
            e->setInformation(Eigen::Matrix2d::Identity() * invSigma2);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(deltaMono);
            e->fx = P->fx; e->fy = P->fy; e->cx = P->cx; e->cy = P->cy;
            for (int k = 0; k < 3; ++k) e->Xc[k] = Xc.at<float>(k);
            optimizer.addEdge(e);
            eMono.push_back(e); iMono.push_back(i);
        } else {
            Eigen::Matrix<double, 3, 1> obs;
            obs << P->obs[3 * i], P->obs[3 * i + 1], P->obs[3 * i + 2];
            g2o::EdgeStereoSE3ProjectXYZOnlyTranslation* e = new g2o::EdgeStereoSE3ProjectXYZOnlyTranslation();
            e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            e->setMeasurement(obs);
            Eigen::Matrix3d Info = Eigen::Matrix3d::Identity() * invSigma2;
            e->setInformation(Info);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            e->setRobustKernel(rk);
            rk->setDelta(deltaStereo);
            e->fx = P->fx; e->fy = P->fy; e->cx = P->cx; e->cy = P->cy; e->bf = P->bf;
            for (int k = 0; k < 3; ++k) e->Xc[k] = Xc.at<float>(k);
            optimizer.addEdge(e);
            eStereo.push_back(e); iStereo.push_back(i);
        }
    }
    std::vector<EdgeLineProjectXYZOnlyTranslation*> eLs, eLe;
    for (int i = 0; i < P->n_lines; ++i) {
        o_line[i] = 0;
        Eigen::Vector3d line_obs(P->line_obs[3 * i], P->line_obs[3 * i + 1], P->line_obs[3 * i + 2]);
        for (int s = 0; s < 2; ++s) {
            EdgeLineProjectXYZOnlyTranslation* el = new EdgeLineProjectXYZOnlyTranslation();
            el->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
            el->setMeasurement(line_obs);
            el->setInformation(Eigen::Matrix3d::Identity() * 1);
            g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
            el->setRobustKernel(rk);
            rk->setDelta(deltaStereo);
            el->fx = P->fx; el->fy = P->fy; el->cx = P->cx; el->cy = P->cy;
            cv::Mat Xw = Converter::toCvVec(Eigen::Vector3d(P->line_Xw[6 * i + 3 * s], P->line_Xw[6 * i + 3 * s + 1], P->line_Xw[6 * i + 3 * s + 2]));
            cv::Mat Xc = R_cw * Xw;
            for (int k = 0; k < 3; ++k) el->Xc[k] = Xc.at<float>(k);
            optimizer.addEdge(el);
            (s == 0 ? eLs : eLe).push_back(el);
        }
    }
    for (int i = 0; i < 16; ++i) Tcw_d[i] = Tcw_in[i];
    for (int i = 0; i < 4; ++i) iters[i] = -1;
    for (int i = 0; i < P->n_planes; ++i) o_plane[i] = 0;
    if (nInitialCorrespondences < 3) return 0;
    double angleInfo = P->angle_info; angleInfo = 3282.8 / (angleInfo * angleInfo);
    double disInfo = P->dist_info; disInfo = disInfo * disInfo;
    const double planeChi = P->plane_chi;
    const float deltaPlane = sqrt(planeChi);
    std::vector<g2o::EdgePlaneOnlyTranslation*> ePl;
    for (int i = 0; i < P->n_planes; ++i) {
        g2o::EdgePlaneOnlyTranslation* e = new g2o::EdgePlaneOnlyTranslation();
        e->setVertex(0, dynamic_cast<g2o::OptimizableGraph::Vertex*>(optimizer.vertex(0)));
        e->setMeasurement(Converter::toPlane3D(coeff4(P->plane_meas + 4 * i)));
        Eigen::Matrix3d Info;
        Info << angleInfo, 0, 0, 0, angleInfo, 0, 0, 0, disInfo;
        e->setInformation(Info);
        g2o::Plane3D Xw = Converter::toPlane3D(coeff4(P->plane_map + 4 * i));
        Xw.rotateNormal(Converter::toMatrix3d(R_cw));
        e->Xc = Xw;
        g2o::RobustKernelHuber* rk = new g2o::RobustKernelHuber;
        e->setRobustKernel(rk);
        rk->setDelta(deltaPlane);
        optimizer.addEdge(e);
        ePl.push_back(e);
        e->computeError();
    }
    const float chi2Mono[4] = {5.991, 5.991, 5.991, 5.991};
    const float chi2Stereo[4] = {7.815, 7.815, 7.815, 7.815};
    const int its[4] = {10, 10, 10, 10};
    int nBad = 0;
    for (size_t it = 0; it < 4; it++) {
        vSE3->setEstimate(Converter::toSE3Quat(mTcw));
        optimizer.initializeOptimization(0);
        iters[it] = optimizer.optimize(its[it]);
        nBad = 0;
        for (size_t i = 0; i < eMono.size(); i++) {
            g2o::EdgeSE3ProjectXYZOnlyTranslation* e = eMono[i];
            const int idx = iMono[i];
            if (o_pt[idx]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > chi2Mono[it]) { o_pt[idx] = 1; e->setLevel(1); nBad++; } else { o_pt[idx] = 0; e->setLevel(0); }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < eStereo.size(); i++) {
            g2o::EdgeStereoSE3ProjectXYZOnlyTranslation* e = eStereo[i];
            const int idx = iStereo[i];
            if (o_pt[idx]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > chi2Stereo[it]) { o_pt[idx] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_pt[idx] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        for (size_t i = 0; i < eLs.size(); i++) {
            EdgeLineProjectXYZOnlyTranslation *e1 = eLs[i], *e2 = eLe[i];
            if (o_line[i]) { e1->computeError(); e2->computeError(); }
            const float chi2_s = e1->chiline(), chi2_e = e2->chiline();
            if (chi2_s > 2 * chi2Mono[it] || chi2_e > 2 * chi2Mono[it]) { o_line[i] = 1; e1->setLevel(1); e2->setLevel(1); }
            else { o_line[i] = 0; e1->setLevel(0); e2->setLevel(0); }
            if (it == 2) { e1->setRobustKernel(0); e2->setRobustKernel(0); }
        }
        for (size_t i = 0; i < ePl.size(); i++) {
            g2o::EdgePlaneOnlyTranslation* e = ePl[i];
            if (o_plane[i]) e->computeError();
            const float chi2 = e->chi2();
            if (chi2 > planeChi) { o_plane[i] = 1; e->setLevel(1); nBad++; } else { e->setLevel(0); o_plane[i] = 0; }
            if (it == 2) e->setRobustKernel(0);
        }
        if (optimizer.edges().size() < 10) break;
    }
    g2o::VertexSE3Expmap* vSE3_recov = static_cast<g2o::VertexSE3Expmap*>(optimizer.vertex(0));
    const Eigen::Matrix<double, 4, 4> T = vSE3_recov->estimate().to_homogeneous_matrix();
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) Tcw_d[4 * i + j] = T(i, j);
    return nInitialCorrespondences - nBad;
}
}
