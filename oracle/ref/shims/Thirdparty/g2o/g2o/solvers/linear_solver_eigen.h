// TEST INFRASTRUCTURE ONLY.  Stand-in for Thirdparty/g2o/g2o/solvers/linear_solver_eigen.h, the one g2o header that cannot be compiled against
// the dense Eigen stand-in: it is a thin wrapper over Eigen's SPARSE Cholesky (SimplicialLDLT subclassed to inject an AMD block ordering) and
// pokes at Eigen internals; all of its arithmetic is Eigen's.  LocalBundleAdjustment only needs "solve the Schur-reduced pose system", so the
// class keeps its name and interface and solves densely with the reference's own LinearSolverDense (same solution up to rounding).
#pragma once
#include <Thirdparty/g2o/g2o/solvers/linear_solver_dense.h>   // the reference's (only linear_solver_eigen.h is shadowed)

namespace g2o {
template <typename MatrixType>
class LinearSolverEigen : public LinearSolverDense<MatrixType> {
public:
    LinearSolverEigen() : LinearSolverDense<MatrixType>() {}
    bool blockOrdering() const { return false; }
    void setBlockOrdering(bool) {}
    void setWriteDebug(bool) {}
    bool writeDebug() const { return false; }
};
}  // namespace g2o
