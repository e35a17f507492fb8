// TEST INFRASTRUCTURE ONLY.  A small eager dense-matrix library with Eigen's names, enough for the reference's vendored g2o
// (Thirdparty/g2o), its SE3 / Plane3D types and edge classes to compile UNMODIFIED in a container without Eigen (oracle/ref/, `make -C
// oracle ref`).  Column-major storage like Eigen (g2o maps raw Hessian / Jacobian memory), fixed and dynamic sizes, Map, lvalue blocks,
// LDLT / LLT, Quaternion, AngleAxis, Isometry transform.  Products accumulate left to right over k; Eigen's own unrolled / vectorised
// reductions round differently, so results agree with a real Eigen build to rounding, not bit for bit - which is also the level at which the
// oracle's restatement of g2o is compared with this build (tests/test_oracle_pose_ref.py).
#pragma once
#include <algorithm>
#include <cassert>
#include <climits>
#include <cmath>
#include <cstddef>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Unaligned = 0, Aligned = 1 };
enum { AlignedBit = 0x80 };
enum { Affine = 1, Isometry = 2 };
enum { EigenvaluesOnly = 0x40, ComputeEigenvectors = 0x80 };
enum { Lower = 1, Upper = 2 };
inline void initParallel() {}
template <class T> using aligned_allocator = std::allocator<T>;
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW_IF_VECTORIZABLE_FIXED_SIZE(a, b)
#define EIGEN_STRONG_INLINE inline
#define EIGEN_VERSION_AT_LEAST(x, y, z) 1
template <class M, int UpLo = 1> class LDLT;
template <class M, int UpLo = 1> class LLT;

template <class S, int R, int C, int Opt = ColMajor, int MR = R, int MC = C> class Matrix;
template <class M, int MapOpt = Unaligned, class Stride = void> class Map;
template <class X> class Block;
template <class X> class Diagonal;
template <class D> struct traits;

template <class X> struct is_const_xpr { enum { value = 0 }; };

template <class Derived>
class DenseBase {
public:
    typedef typename traits<Derived>::Scalar Scalar;
    enum { RowsAtCompileTime = traits<Derived>::Rows, ColsAtCompileTime = traits<Derived>::Cols, SizeAtCompileTime = (traits<Derived>::Rows < 0 || traits<Derived>::Cols < 0) ? -1 : traits<Derived>::Rows * traits<Derived>::Cols,
           IsVectorAtCompileTime = traits<Derived>::Rows == 1 || traits<Derived>::Cols == 1, Flags = AlignedBit };
    typedef Matrix<Scalar, traits<Derived>::Rows, traits<Derived>::Cols> PlainObject;
    typedef Matrix<Scalar, traits<Derived>::Cols, traits<Derived>::Rows> TransposedPlain;
    Derived& derived() { return *static_cast<Derived*>(this); }
    const Derived& derived() const { return *static_cast<const Derived*>(this); }
    int rows() const { return derived().rows_(); }
    int cols() const { return derived().cols_(); }
    int size() const { return rows() * cols(); }
    Scalar coeff(int i, int j) const { return derived().c_(i, j); }
    Scalar& coeffRef(int i, int j) { return derived().r_(i, j); }
    Scalar operator()(int i, int j) const { return coeff(i, j); }
    Scalar& operator()(int i, int j) { return coeffRef(i, j); }
    Scalar vget(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
    Scalar& vref(int i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
    Scalar operator()(int i) const { return vget(i); }
    Scalar& operator()(int i) { return vref(i); }
    Scalar operator[](int i) const { return vget(i); }
    Scalar& operator[](int i) { return vref(i); }
    Scalar x() const { return vget(0); } Scalar y() const { return vget(1); } Scalar z() const { return vget(2); } Scalar w() const { return vget(3); }
    Scalar& x() { return vref(0); } Scalar& y() { return vref(1); } Scalar& z() { return vref(2); } Scalar& w() { return vref(3); }
    PlainObject eval() const { return PlainObject(*this); }
    TransposedPlain transpose() const {
        TransposedPlain t(cols(), rows());
        for (int i = 0; i < rows(); ++i) for (int j = 0; j < cols(); ++j) t.r_(j, i) = coeff(i, j);
        return t;
    }
    template <class O> Scalar dot(const DenseBase<O>& o) const { Scalar s = 0; for (int i = 0; i < size(); ++i) s += vget(i) * o.vget(i); return s; }
    Scalar squaredNorm() const { Scalar s = 0; for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j); return s; }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar sum() const { Scalar s = 0; for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) s += coeff(i, j); return s; }
    Scalar trace() const { Scalar s = 0; for (int i = 0; i < rows(); ++i) s += coeff(i, i); return s; }
    Scalar maxCoeff() const { Scalar m = coeff(0, 0); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) m = std::max(m, coeff(i, j)); return m; }
    Scalar minCoeff() const { Scalar m = coeff(0, 0); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) m = std::min(m, coeff(i, j)); return m; }
    PlainObject normalized() const { PlainObject p(*this); const Scalar n = norm(); for (int i = 0; i < p.size(); ++i) p.vref(i) = p.vget(i) / n; return p; }
    PlainObject cwiseAbs() const { PlainObject p(*this); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) p.r_(i, j) = std::abs(coeff(i, j)); return p; }
    PlainObject array() const { return PlainObject(*this); }
    PlainObject matrix() const { return PlainObject(*this); }
    bool allFinite() const { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) if (!std::isfinite((double)coeff(i, j))) return false; return true; }
    template <class O> Matrix<Scalar, 3, 1> cross(const DenseBase<O>& o) const {
        return Matrix<Scalar, 3, 1>(vget(1) * o.vget(2) - vget(2) * o.vget(1), vget(2) * o.vget(0) - vget(0) * o.vget(2), vget(0) * o.vget(1) - vget(1) * o.vget(0));
    }
    Scalar determinant() const;
    PlainObject inverse() const;
    struct LuProxy { PlainObject inv; template <class O> typename DenseBase<O>::PlainObject solve(const DenseBase<O>& b) const { return inv * b; } PlainObject inverse() const { return inv; } };
    LuProxy lu() const { return LuProxy{inverse()}; }
    LuProxy partialPivLu() const { return LuProxy{inverse()}; }
    LLT<PlainObject, 1> llt() const;
    LDLT<PlainObject, 1> ldlt() const;
    template <class O> bool operator==(const DenseBase<O>& o) const { if (rows() != o.rows() || cols() != o.cols()) return false; for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) if (coeff(i, j) != o.coeff(i, j)) return false; return true; }
    template <class O> bool operator!=(const DenseBase<O>& o) const { return !(*this == o); }
    // ---- mutators (Derived must be writable) ----
    Derived& noalias() { return derived(); }
    Derived& setZero() { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) = Scalar(0); return derived(); }
    Derived& setIdentity() { setZero(); for (int i = 0; i < std::min(rows(), cols()); ++i) coeffRef(i, i) = Scalar(1); return derived(); }
    Derived& fill(Scalar v) { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) = v; return derived(); }
    Derived& setConstant(Scalar v) { return fill(v); }
    void normalize() { const Scalar n = norm(); for (int i = 0; i < size(); ++i) vref(i) = vget(i) / n; }
    template <class O> Derived& assign(const DenseBase<O>& o) {
        derived().resizeLike_(o.rows(), o.cols());
        if ((const void*)&o == (const void*)this) return derived();
        typename DenseBase<O>::PlainObject tmp(o);            // eager: protects against aliasing
        for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) = tmp.c_(i, j);
        return derived();
    }
    template <class O> Derived& operator+=(const DenseBase<O>& o) { typename DenseBase<O>::PlainObject t(o); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) += t.c_(i, j); return derived(); }
    template <class O> Derived& operator-=(const DenseBase<O>& o) { typename DenseBase<O>::PlainObject t(o); for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) -= t.c_(i, j); return derived(); }
    Derived& operator*=(Scalar s) { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) *= s; return derived(); }
    Derived& operator/=(Scalar s) { for (int j = 0; j < cols(); ++j) for (int i = 0; i < rows(); ++i) coeffRef(i, j) /= s; return derived(); }
    // ---- blocks ----
    Block<Derived> block(int i, int j, int r, int c) { return Block<Derived>(derived(), i, j, r, c); }
    Block<const Derived> block(int i, int j, int r, int c) const { return Block<const Derived>(derived(), i, j, r, c); }
    template <int BR, int BC> Block<Derived> block(int i, int j) { return Block<Derived>(derived(), i, j, BR, BC); }
    template <int BR, int BC> Block<const Derived> block(int i, int j) const { return Block<const Derived>(derived(), i, j, BR, BC); }
    template <int BR, int BC> Block<Derived> topLeftCorner() { return Block<Derived>(derived(), 0, 0, BR, BC); }
    template <int BR, int BC> Block<const Derived> topLeftCorner() const { return Block<const Derived>(derived(), 0, 0, BR, BC); }
    Block<Derived> col(int j) { return Block<Derived>(derived(), 0, j, rows(), 1); }
    Block<const Derived> col(int j) const { return Block<const Derived>(derived(), 0, j, rows(), 1); }
    Block<Derived> row(int i) { return Block<Derived>(derived(), i, 0, 1, cols()); }
    Block<const Derived> row(int i) const { return Block<const Derived>(derived(), i, 0, 1, cols()); }
    Block<Derived> segment(int s, int n) { return cols() == 1 ? Block<Derived>(derived(), s, 0, n, 1) : Block<Derived>(derived(), 0, s, 1, n); }
    Block<const Derived> segment(int s, int n) const { return cols() == 1 ? Block<const Derived>(derived(), s, 0, n, 1) : Block<const Derived>(derived(), 0, s, 1, n); }
    template <int N> Block<Derived> segment(int s) { return segment(s, N); }
    template <int N> Block<const Derived> segment(int s) const { return segment(s, N); }
    Block<Derived> head(int n) { return segment(0, n); }
    Block<const Derived> head(int n) const { return segment(0, n); }
    template <int N> Block<Derived> head() { return segment(0, N); }
    template <int N> Block<const Derived> head() const { return segment(0, N); }
    Block<Derived> tail(int n) { return segment(size() - n, n); }
    template <int N> Block<Derived> tail() { return segment(size() - N, N); }
    template <int N> Block<const Derived> tail() const { return segment(size() - N, N); }
    Diagonal<Derived> diagonal() { return Diagonal<Derived>(derived()); }
    Diagonal<const Derived> diagonal() const { return Diagonal<const Derived>(derived()); }
};

template <class D> using MatrixBase = DenseBase<D>;

// ---- storage ----
template <class S, int R, int C, bool Dyn = (R < 0 || C < 0)> struct DenseStorage;
template <class S, int R, int C> struct DenseStorage<S, R, C, false> {
    S d[R * C > 0 ? R * C : 1];
    DenseStorage() { for (int i = 0; i < R * C; ++i) d[i] = S(0); }
    int rows() const { return R; } int cols() const { return C; }
    void resize(int, int) {}
    S* data() { return d; } const S* data() const { return d; }
};
template <class S, int R, int C> struct DenseStorage<S, R, C, true> {
    std::vector<S> d; int r = R < 0 ? 0 : R, c = C < 0 ? 0 : C;
    int rows() const { return r; } int cols() const { return c; }
    void resize(int rr, int cc) { if (rr != r || cc != c || (int)d.size() != rr * cc) { r = rr; c = cc; d.assign((size_t)rr * cc, S(0)); } }
    S* data() { return d.data(); } const S* data() const { return d.data(); }
};

template <class S, int R, int C, int Opt, int MR, int MC>
struct traits<Matrix<S, R, C, Opt, MR, MC>> { typedef S Scalar; enum { Rows = R, Cols = C, Options = Opt }; };

template <class S> struct CommaInit;

template <class S, int R, int C, int Opt, int MR, int MC>
class Matrix : public DenseBase<Matrix<S, R, C, Opt, MR, MC>> {
public:
    typedef DenseBase<Matrix> Base;
    typedef S Scalar;
    enum { Rows = R, Cols = C, Options = Opt };
    typedef Map<Matrix, Unaligned> MapType;
    typedef Map<const Matrix, Unaligned> ConstMapType;
    typedef Map<Matrix, Aligned> AlignedMapType;
    typedef Map<const Matrix, Aligned> ConstAlignedMapType;
    Matrix() {}
    explicit Matrix(int n) { if (R < 0 && C < 0) st.resize(n, n); else if (C == 1 || R < 0) st.resize(R < 0 ? n : R, C < 0 ? 1 : C); else st.resize(R, C < 0 ? n : C); if (C == 1 && R < 0) st.resize(n, 1); if (R == 1 && C < 0) st.resize(1, n); }
    Matrix(int r, int c) { if (R == 2 && C == 1) { st.d[0] = S(r); st.d[1] = S(c); } else st.resize(r, c); }
    Matrix(S a, S b) { static_assert(R * C == 2 || R < 0 || C < 0, "2-vector"); st.resize(R < 0 ? (int)a : R, C < 0 ? (int)b : C); if (R * C == 2) { st.data()[0] = a; st.data()[1] = b; } }
    Matrix(S a, S b, S c) { static_assert(R * C == 3, "3-vector"); st.d[0] = a; st.d[1] = b; st.d[2] = c; }
    Matrix(S a, S b, S c, S d) { static_assert(R * C == 4, "4-vector"); st.d[0] = a; st.d[1] = b; st.d[2] = c; st.d[3] = d; }
    Matrix(const Matrix& o) : st(o.st) {}
    template <class O> Matrix(const DenseBase<O>& o) { st.resize(o.rows(), o.cols()); for (int j = 0; j < o.cols(); ++j) for (int i = 0; i < o.rows(); ++i) r_(i, j) = o.coeff(i, j); }
    Matrix& operator=(const Matrix& o) { st = o.st; return *this; }
    template <class O> Matrix& operator=(const DenseBase<O>& o) { return this->assign(o); }
    int rows_() const { return st.rows(); }
    int cols_() const { return st.cols(); }
    int idx(int i, int j) const { return (Opt & RowMajor) ? i * st.cols() + j : i + j * st.rows(); }
    S c_(int i, int j) const { return st.data()[idx(i, j)]; }
    S& r_(int i, int j) { return st.data()[idx(i, j)]; }
    void resizeLike_(int r, int c) { st.resize(r, c); }
    void resize(int r, int c) { st.resize(r, c); }
    void resize(int n) { if (C == 1) st.resize(n, 1); else if (R == 1) st.resize(1, n); else st.resize(n, n); }
    void conservativeResize(int r, int c) { Matrix t(*this); st.resize(r, c); for (int j = 0; j < std::min(c, t.cols()); ++j) for (int i = 0; i < std::min(r, t.rows()); ++i) r_(i, j) = t.c_(i, j); }
    S* data() { return st.data(); }
    const S* data() const { return st.data(); }
    static Matrix Zero() { return Matrix(); }
    static Matrix Zero(int r, int c) { Matrix m; m.st.resize(r, c); m.setZero(); return m; }
    static Matrix Zero(int n) { Matrix m(n); m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Identity(int r, int c) { Matrix m; m.st.resize(r, c); m.setIdentity(); return m; }
    static Matrix Ones() { Matrix m; m.fill(S(1)); return m; }
    static Matrix Constant(S v) { Matrix m; m.fill(v); return m; }
    static Matrix UnitX() { Matrix m; m.vref(0) = 1; return m; }
    static Matrix UnitY() { Matrix m; m.vref(1) = 1; return m; }
    static Matrix UnitZ() { Matrix m; m.vref(2) = 1; return m; }
    CommaInit<Matrix> operator<<(S v);
    template <class O> CommaInit<Matrix> operator<<(const DenseBase<O>& o);
private:
    DenseStorage<S, R, C> st;
};

template <class M> struct CommaInit {
    M& m; int r, c, rowh;
    typedef typename M::Scalar S;
    CommaInit& operator,(S v) { if (c == m.cols()) { r += rowh; c = 0; rowh = 1; } m.r_(r, c++) = v; return *this; }
    template <class O> CommaInit& operator,(const DenseBase<O>& o) {
        if (c == m.cols()) { r += rowh; c = 0; }
        for (int i = 0; i < o.rows(); ++i) for (int j = 0; j < o.cols(); ++j) m.r_(r + i, c + j) = o.coeff(i, j);
        c += o.cols(); rowh = o.rows();
        return *this;
    }
    M& finished() { return m; }
    operator M&() { return m; }
};
template <class S, int R, int C, int Opt, int MR, int MC> CommaInit<Matrix<S, R, C, Opt, MR, MC>> Matrix<S, R, C, Opt, MR, MC>::operator<<(S v) {
    CommaInit<Matrix> ci{*this, 0, 0, 1};
    ci, v;
    return ci;
}
template <class S, int R, int C, int Opt, int MR, int MC> template <class O> CommaInit<Matrix<S, R, C, Opt, MR, MC>> Matrix<S, R, C, Opt, MR, MC>::operator<<(const DenseBase<O>& o) {
    CommaInit<Matrix> ci{*this, 0, 0, 1};
    ci, o;
    return ci;
}

// ---- Map ----
template <class M, int MapOpt, class Stride>
struct traits<Map<M, MapOpt, Stride>> { typedef typename std::remove_const<M>::type::Scalar Scalar; enum { Rows = std::remove_const<M>::type::Rows, Cols = std::remove_const<M>::type::Cols, Options = std::remove_const<M>::type::Options }; };
template <class M, int MapOpt, class Stride>
class Map : public DenseBase<Map<M, MapOpt, Stride>> {
public:
    typedef typename std::remove_const<M>::type Plain;
    typedef typename Plain::Scalar S;
    typedef typename std::conditional<std::is_const<M>::value, const S*, S*>::type Ptr;
    Map(Ptr p, int r = Plain::Rows, int c = Plain::Cols) : p_(const_cast<S*>(p)), r_n(Plain::Rows < 0 ? r : Plain::Rows), c_n(Plain::Cols < 0 ? (Plain::Cols == -1 && Plain::Rows != -1 && c == Plain::Cols ? 1 : c) : Plain::Cols) {
        if (Plain::Rows < 0 && Plain::Cols == 1) { r_n = r; c_n = 1; }                 // VectorXd::MapType(ptr, n)
        if (Plain::Rows == 1 && Plain::Cols < 0) { r_n = 1; c_n = r; }
    }
    Map(const Map& o) : p_(o.p_), r_n(o.r_n), c_n(o.c_n) {}
    Map& operator=(const Map& o) { return this->assign(o); }
    template <class O> Map& operator=(const DenseBase<O>& o) { return this->assign(o); }
    int rows_() const { return r_n; }
    int cols_() const { return c_n; }
    int idx(int i, int j) const { return (Plain::Options & RowMajor) ? i * c_n + j : i + j * r_n; }
    S c_(int i, int j) const { return p_[idx(i, j)]; }
    S& r_(int i, int j) { return p_[idx(i, j)]; }
    void resizeLike_(int, int) {}
    S* data() { return p_; }
    const S* data() const { return p_; }
    // placement-new re-seating, used by g2o: new (&map) MapType(ptr)
private:
    S* p_; int r_n, c_n;
};

// ---- Block (lvalue view) ----
template <class X>
struct traits<Block<X>> { typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar; enum { Rows = Dynamic, Cols = Dynamic, Options = 0 }; };
template <class X>
class Block : public DenseBase<Block<X>> {
public:
    typedef typename std::remove_const<X>::type XP;
    typedef typename traits<XP>::Scalar S;
    Block(X& x, int i, int j, int r, int c) : x_(const_cast<XP*>(&x)), i0(i), j0(j), r_n(r), c_n(c) {}
    Block(const Block& o) : x_(o.x_), i0(o.i0), j0(o.j0), r_n(o.r_n), c_n(o.c_n) {}
    Block& operator=(const Block& o) { return this->assign(o); }
    template <class O> Block& operator=(const DenseBase<O>& o) { return this->assign(o); }
    int rows_() const { return r_n; }
    int cols_() const { return c_n; }
    S c_(int i, int j) const { return x_->c_(i0 + i, j0 + j); }
    S& r_(int i, int j) { return x_->r_(i0 + i, j0 + j); }
    void resizeLike_(int, int) {}
private:
    XP* x_; int i0, j0, r_n, c_n;
};

// ---- Diagonal (lvalue view, n x 1) ----
template <class X>
struct traits<Diagonal<X>> { typedef typename traits<typename std::remove_const<X>::type>::Scalar Scalar; enum { Rows = Dynamic, Cols = 1, Options = 0 }; };
template <class X>
class Diagonal : public DenseBase<Diagonal<X>> {
public:
    typedef typename std::remove_const<X>::type XP;
    typedef typename traits<XP>::Scalar S;
    explicit Diagonal(X& x) : x_(const_cast<XP*>(&x)) {}
    Diagonal(const Diagonal& o) : x_(o.x_) {}
    Diagonal& operator=(const Diagonal& o) { return this->assign(o); }
    template <class O> Diagonal& operator=(const DenseBase<O>& o) { return this->assign(o); }
    int rows_() const { return std::min(x_->rows(), x_->cols()); }
    int cols_() const { return 1; }
    S c_(int i, int) const { return x_->c_(i, i); }
    S& r_(int i, int) { return x_->r_(i, i); }
    void resizeLike_(int, int) {}
    struct ArrayRef {
        Diagonal d;
        ArrayRef& operator+=(S v) { for (int i = 0; i < d.rows_(); ++i) d.r_(i, 0) += v; return *this; }
        ArrayRef& operator-=(S v) { for (int i = 0; i < d.rows_(); ++i) d.r_(i, 0) -= v; return *this; }
        ArrayRef& operator*=(S v) { for (int i = 0; i < d.rows_(); ++i) d.r_(i, 0) *= v; return *this; }
    };
    ArrayRef array() { return ArrayRef{*this}; }
private:
    XP* x_;
};

// ---- arithmetic (eager) ----
template <int A, int B> struct dim_pick { enum { value = A >= 0 ? A : B }; };
template <class A, class B>
Matrix<typename traits<A>::Scalar, dim_pick<traits<A>::Rows, traits<B>::Rows>::value, dim_pick<traits<A>::Cols, traits<B>::Cols>::value> operator+(const DenseBase<A>& a, const DenseBase<B>& b) {
    Matrix<typename traits<A>::Scalar, dim_pick<traits<A>::Rows, traits<B>::Rows>::value, dim_pick<traits<A>::Cols, traits<B>::Cols>::value> m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = a.coeff(i, j) + b.coeff(i, j);
    return m;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, dim_pick<traits<A>::Rows, traits<B>::Rows>::value, dim_pick<traits<A>::Cols, traits<B>::Cols>::value> operator-(const DenseBase<A>& a, const DenseBase<B>& b) {
    Matrix<typename traits<A>::Scalar, dim_pick<traits<A>::Rows, traits<B>::Rows>::value, dim_pick<traits<A>::Cols, traits<B>::Cols>::value> m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = a.coeff(i, j) - b.coeff(i, j);
    return m;
}
template <class A> typename DenseBase<A>::PlainObject operator-(const DenseBase<A>& a) {
    typename DenseBase<A>::PlainObject m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = -a.coeff(i, j);
    return m;
}
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const DenseBase<A>& a, const DenseBase<B>& b) {
    Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> m;
    m.resizeLike_(a.rows(), b.cols());
    for (int j = 0; j < b.cols(); ++j)
        for (int i = 0; i < a.rows(); ++i) {
            typename traits<A>::Scalar s = 0;
            for (int k = 0; k < a.cols(); ++k) s += a.coeff(i, k) * b.coeff(k, j);
            m.r_(i, j) = s;
        }
    return m;
}
template <class A> typename DenseBase<A>::PlainObject operator*(const DenseBase<A>& a, typename traits<A>::Scalar s) {
    typename DenseBase<A>::PlainObject m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = a.coeff(i, j) * s;
    return m;
}
template <class A> typename DenseBase<A>::PlainObject operator*(typename traits<A>::Scalar s, const DenseBase<A>& a) {
    typename DenseBase<A>::PlainObject m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = s * a.coeff(i, j);
    return m;
}
template <class A> typename DenseBase<A>::PlainObject operator/(const DenseBase<A>& a, typename traits<A>::Scalar s) {
    typename DenseBase<A>::PlainObject m(a);
    for (int j = 0; j < a.cols(); ++j) for (int i = 0; i < a.rows(); ++i) m.r_(i, j) = a.coeff(i, j) / s;
    return m;
}
template <class A> std::ostream& operator<<(std::ostream& os, const DenseBase<A>& a) {
    for (int i = 0; i < a.rows(); ++i) { for (int j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a.coeff(i, j); if (i + 1 < a.rows()) os << "\n"; }
    return os;
}

template <class D> typename DenseBase<D>::Scalar DenseBase<D>::determinant() const {
    const int n = rows();
    if (n == 1) return coeff(0, 0);
    if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(0, 1) * coeff(1, 0);
    if (n == 3) return coeff(0, 0) * (coeff(1, 1) * coeff(2, 2) - coeff(1, 2) * coeff(2, 1)) - coeff(0, 1) * (coeff(1, 0) * coeff(2, 2) - coeff(1, 2) * coeff(2, 0)) +
                       coeff(0, 2) * (coeff(1, 0) * coeff(2, 1) - coeff(1, 1) * coeff(2, 0));
    std::vector<Scalar> a((size_t)n * n);
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) a[(size_t)i * n + j] = coeff(i, j);
    Scalar det = 1;
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::abs(a[(size_t)r * n + c]) > std::abs(a[(size_t)p * n + c])) p = r;
        if (a[(size_t)p * n + c] == 0) return 0;
        if (p != c) { for (int j = 0; j < n; ++j) std::swap(a[(size_t)c * n + j], a[(size_t)p * n + j]); det = -det; }
        det *= a[(size_t)c * n + c];
        for (int r = c + 1; r < n; ++r) { const Scalar f = a[(size_t)r * n + c] / a[(size_t)c * n + c]; for (int j = c; j < n; ++j) a[(size_t)r * n + j] -= f * a[(size_t)c * n + j]; }
    }
    return det;
}
template <class D> typename DenseBase<D>::PlainObject DenseBase<D>::inverse() const {
    const int n = rows();
    PlainObject m; m.resizeLike_(n, n);
    if (n == 3) {                                              // cofactors / determinant, like Eigen's fixed-size 3x3 inverse
        const Scalar a = coeff(0, 0), b = coeff(0, 1), c = coeff(0, 2), d = coeff(1, 0), e = coeff(1, 1), f = coeff(1, 2), g = coeff(2, 0), h = coeff(2, 1), k = coeff(2, 2);
        const Scalar c00 = e * k - f * h, c10 = f * g - d * k, c20 = d * h - e * g;
        const Scalar invdet = Scalar(1) / (a * c00 + b * c10 + c * c20);
        m.r_(0, 0) = c00 * invdet; m.r_(1, 0) = c10 * invdet; m.r_(2, 0) = c20 * invdet;
        m.r_(0, 1) = (c * h - b * k) * invdet; m.r_(1, 1) = (a * k - c * g) * invdet; m.r_(2, 1) = (b * g - a * h) * invdet;
        m.r_(0, 2) = (b * f - c * e) * invdet; m.r_(1, 2) = (c * d - a * f) * invdet; m.r_(2, 2) = (a * e - b * d) * invdet;
        return m;
    }
    std::vector<Scalar> A((size_t)n * 2 * n, Scalar(0));          // Gauss-Jordan, partial pivoting
    for (int i = 0; i < n; ++i) { for (int j = 0; j < n; ++j) A[(size_t)i * 2 * n + j] = coeff(i, j); A[(size_t)i * 2 * n + n + i] = 1; }
    for (int c = 0; c < n; ++c) {
        int p = c;
        for (int r = c + 1; r < n; ++r) if (std::abs(A[(size_t)r * 2 * n + c]) > std::abs(A[(size_t)p * 2 * n + c])) p = r;
        for (int j = 0; j < 2 * n; ++j) std::swap(A[(size_t)c * 2 * n + j], A[(size_t)p * 2 * n + j]);
        const Scalar dd = A[(size_t)c * 2 * n + c];
        for (int j = 0; j < 2 * n; ++j) A[(size_t)c * 2 * n + j] /= dd;
        for (int r = 0; r < n; ++r) if (r != c) { const Scalar f = A[(size_t)r * 2 * n + c]; if (f != 0) for (int j = 0; j < 2 * n; ++j) A[(size_t)r * 2 * n + j] -= f * A[(size_t)c * 2 * n + j]; }
    }
    for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) m.r_(i, j) = A[(size_t)i * 2 * n + n + j];
    return m;
}

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d; typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, 2, 1> Vector2f; typedef Matrix<float, 3, 1> Vector3f; typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 2, 2> Matrix2d; typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d; typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, 3, 3> Matrix3f; typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<int, 2, 1> Vector2i; typedef Matrix<int, 3, 1> Vector3i; typedef Matrix<int, Dynamic, 1> VectorXi;

// ---- Cholesky: unpivoted LDL^T (Eigen's LDLT pivots; same solution up to rounding for the positive definite systems solved here) ----
template <class M, int UpLo>
class LDLT {
public:
    LDLT() {}
    template <class O> explicit LDLT(const DenseBase<O>& a) { compute(a); }
    template <class O> LDLT& compute(const DenseBase<O>& a) {
        n = a.rows(); L.assign((size_t)n * n, 0.0); D.assign(n, 0.0); positive = true;
        for (int j = 0; j < n; ++j) {
            double d = a.coeff(j, j);
            for (int k = 0; k < j; ++k) d -= L[(size_t)j * n + k] * L[(size_t)j * n + k] * D[k];
            D[j] = d;
            if (!(d > 0)) positive = false;
            L[(size_t)j * n + j] = 1;
            for (int i = j + 1; i < n; ++i) {
                double s = a.coeff(i, j);
                for (int k = 0; k < j; ++k) s -= L[(size_t)i * n + k] * L[(size_t)j * n + k] * D[k];
                L[(size_t)i * n + j] = s / d;
            }
        }
        return *this;
    }
    bool isPositive() const { return positive; }
    template <class O> Matrix<double, Dynamic, 1> solve(const DenseBase<O>& b) const {
        Matrix<double, Dynamic, 1> x(n);
        std::vector<double> y(n);
        for (int i = 0; i < n; ++i) { double s = b.vget(i); for (int k = 0; k < i; ++k) s -= L[(size_t)i * n + k] * y[k]; y[i] = s; }
        for (int i = 0; i < n; ++i) y[i] /= D[i];
        for (int i = n - 1; i >= 0; --i) { double s = y[i]; for (int k = i + 1; k < n; ++k) s -= L[(size_t)k * n + i] * x.vget(k); x.vref(i) = s; }
        return x;
    }
    int info() const { return positive ? 0 : 1; }
private:
    int n = 0; std::vector<double> L, D; bool positive = false;
};
template <class M, int UpLo> class LLT : public LDLT<M, UpLo> { public: using LDLT<M, UpLo>::LDLT; };
template <class D> LLT<typename DenseBase<D>::PlainObject, 1> DenseBase<D>::llt() const { return LLT<PlainObject, 1>(*this); }
template <class D> LDLT<typename DenseBase<D>::PlainObject, 1> DenseBase<D>::ldlt() const { return LDLT<PlainObject, 1>(*this); }
enum { Success = 0, NumericalIssue = 1 };

// ---- geometry ----
template <class S, int Opt = 0>
class Quaternion {
public:
    typedef Matrix<S, 3, 1> V3; typedef Matrix<S, 3, 3> M3;
    Quaternion() : q_(0, 0, 0, 1) {}
    Quaternion(S w, S x, S y, S z) : q_(x, y, z, w) {}
    template <class O> explicit Quaternion(const DenseBase<O>& m) { if (m.rows() == 3 && m.cols() == 3) fromMatrix(m); else for (int i = 0; i < 4; ++i) q_[i] = m.vget(i); }
    S x() const { return q_[0]; } S y() const { return q_[1]; } S z() const { return q_[2]; } S w() const { return q_[3]; }
    S& x() { return q_[0]; } S& y() { return q_[1]; } S& z() { return q_[2]; } S& w() { return q_[3]; }
    Matrix<S, 4, 1>& coeffs() { return q_; }
    const Matrix<S, 4, 1>& coeffs() const { return q_; }
    Block<Matrix<S, 4, 1>> vec() { return q_.template head<3>(); }
    V3 vec() const { return V3(q_[0], q_[1], q_[2]); }
    S norm() const { return q_.norm(); }
    S squaredNorm() const { return q_.squaredNorm(); }
    void normalize() { q_.normalize(); }
    Quaternion normalized() const { Quaternion r(*this); r.normalize(); return r; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const { const S n2 = squaredNorm(); return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2); }
    void setIdentity() { q_ = Matrix<S, 4, 1>(0, 0, 0, 1); }
    static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
    Quaternion operator*(const Quaternion& b) const {
        const Quaternion& a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(), a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(), a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion& operator*=(const Quaternion& b) { *this = *this * b; return *this; }
    M3 toRotationMatrix() const {
        M3 r;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w(), txx = tx * x(), txy = ty * x(), txz = tz * x(), tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        r(0, 0) = S(1) - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = S(1) - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = S(1) - (txx + tyy);
        return r;
    }
    M3 matrix() const { return toRotationMatrix(); }
    template <class O> V3 operator*(const DenseBase<O>& v) const {          // Eigen's _transformVector
        const V3 u = vec(), vv(v.vget(0), v.vget(1), v.vget(2));
        const V3 uv = u.cross(vv) * S(2);
        return vv + uv * w() + u.cross(uv);
    }
    template <class O> V3 _transformVector(const DenseBase<O>& v) const { return (*this) * v; }
    template <class O> Quaternion& operator=(const DenseBase<O>& m) { fromMatrix(m); return *this; }
private:
    template <class O> void fromMatrix(const DenseBase<O>& mat) {         // Eigen's quaternionbase_assign_impl<Other, 3, 3>
        S t = mat.coeff(0, 0) + mat.coeff(1, 1) + mat.coeff(2, 2);
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t; y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t; z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            int i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const int j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + S(1.0));
            q_[i] = S(0.5) * t;
            t = S(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            q_[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            q_[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
    }
    Matrix<S, 4, 1> q_;                                                 // x y z w, like Eigen's coeffs()
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S>
class AngleAxis {
public:
    typedef Matrix<S, 3, 1> V3; typedef Matrix<S, 3, 3> M3;
    AngleAxis() : a_(0), ax_(1, 0, 0) {}
    template <class O> AngleAxis(S angle, const DenseBase<O>& axis) : a_(angle), ax_(axis.vget(0), axis.vget(1), axis.vget(2)) {}
    S angle() const { return a_; }
    const V3& axis() const { return ax_; }
    operator Quaternion<S>() const { return toQuaternion(); }
    Quaternion<S> toQuaternion() const { const S s = std::sin(S(0.5) * a_), c = std::cos(S(0.5) * a_); return Quaternion<S>(c, s * ax_[0], s * ax_[1], s * ax_[2]); }
    Quaternion<S> operator*(const AngleAxis& o) const { return toQuaternion() * o.toQuaternion(); }
    Quaternion<S> operator*(const Quaternion<S>& o) const { return toQuaternion() * o; }
    M3 toRotationMatrix() const {                                       // Eigen's AngleAxis::toRotationMatrix
        M3 res;
        const S sin_a = std::sin(a_), cos_a = std::cos(a_);
        const V3 sin_axis = ax_ * sin_a, cos1_axis = ax_ * (S(1) - cos_a);
        S tmp;
        tmp = cos1_axis.x() * ax_.y(); res(0, 1) = tmp - sin_axis.z(); res(1, 0) = tmp + sin_axis.z();
        tmp = cos1_axis.x() * ax_.z(); res(0, 2) = tmp + sin_axis.y(); res(2, 0) = tmp - sin_axis.y();
        tmp = cos1_axis.y() * ax_.z(); res(1, 2) = tmp - sin_axis.x(); res(2, 1) = tmp + sin_axis.x();
        res(0, 0) = cos1_axis.x() * ax_.x() + cos_a; res(1, 1) = cos1_axis.y() * ax_.y() + cos_a; res(2, 2) = cos1_axis.z() * ax_.z() + cos_a;
        return res;
    }
    M3 matrix() const { return toRotationMatrix(); }
    template <class O> V3 operator*(const DenseBase<O>& v) const { return toRotationMatrix() * V3(v.vget(0), v.vget(1), v.vget(2)); }
private:
    S a_; V3 ax_;
};
typedef AngleAxis<double> AngleAxisd;

template <class S, int Dim, int Mode, int Opt = 0>
class Transform {
public:
    typedef Matrix<S, Dim + 1, Dim + 1> M; typedef Matrix<S, Dim, Dim> L; typedef Matrix<S, Dim, 1> V;
    Transform() { m_.setIdentity(); }
    template <class O> Transform(const DenseBase<O>& o) { set(o); }
    Transform(const Quaternion<S>& q) { m_.setIdentity(); linearRef() = q.toRotationMatrix(); }
    template <class O> Transform& operator=(const DenseBase<O>& o) { set(o); return *this; }
    Transform& operator=(const Quaternion<S>& q) { m_.setIdentity(); linearRef() = q.toRotationMatrix(); return *this; }
    static Transform Identity() { return Transform(); }
    void setIdentity() { m_.setIdentity(); }
    M& matrix() { return m_; }
    const M& matrix() const { return m_; }
    Block<M> linear() { return m_.template block<Dim, Dim>(0, 0); }
    L linear() const { return L(m_.template block<Dim, Dim>(0, 0)); }
    L rotation() const { return linear(); }                             // Isometry mode: the linear part is the rotation
    Block<M> translation() { return m_.template block<Dim, 1>(0, Dim); }
    V translation() const { return V(m_.template block<Dim, 1>(0, Dim)); }
    S operator()(int i, int j) const { return m_(i, j); }
    S& operator()(int i, int j) { return m_(i, j); }
    Transform inverse() const {
        Transform r;
        const L Rt = linear().transpose();
        r.linearRef() = Rt;
        r.translation() = -(Rt * translation());
        return r;
    }
    Transform operator*(const Transform& o) const { Transform r; r.m_ = m_ * o.m_; return r; }
    template <class O> V operator*(const DenseBase<O>& v) const { return linear() * V(v) + translation(); }
    Transform& translate(const V& t) { translation() = translation() + linear() * t; return *this; }
private:
    Block<M> linearRef() { return m_.template block<Dim, Dim>(0, 0); }
    template <class O> void set(const DenseBase<O>& o) {
        m_.setIdentity();
        for (int i = 0; i < o.rows(); ++i) for (int j = 0; j < o.cols(); ++j) m_(i, j) = o.coeff(i, j);
    }
    M m_;
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<double, 2, Isometry> Isometry2d;

}  // namespace Eigen
