// cafe_eigen_shim.hpp — TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// A small, eager, dense stand-in for the part of the Eigen 3 API that the reference's HS-DDP solver and HKD problem code
// use (HSDDPSolver/{header,source,common}, HKDMPC/HKD-TrajOpt, Reference/QuadReference, common/). Eigen itself is not in
// this image; with this header on the include path as <eigen3/Eigen/...> the reference's own solver sources compile
// UNCHANGED from where they lie under /root/reference (oracle/refbuild/Makefile), which lets the tests pin the oracle's
// restatement of the solver layer (and through it the GPU path) against the reference's own code.
//
// What it is not: Eigen's kernels. Every operation here is evaluated eagerly into a plain matrix with straightforward
// loops (products accumulate k = 0..K-1 in order, no FMA contraction requested, no vectorised reduction trees), so results
// agree with an Eigen build of the reference to rounding (~1e-13 relative per operation), not bit for bit. The control
// flow of the solver (iteration counts, line-search trials, regularisation steps, AL / ReB updates, MPC shifting) is the
// reference's own. LDLT follows Eigen 3.3's ldlt_inplace<Lower>::unblocked (diagonal pivoting, sign tracking, isPositive
// accepting semidefinite matrices) and LDLT::_solve_impl (pseudo-inverse of D).
//
// With EIGEN_INITIALIZE_MATRICES_BY_NAN defined (HSDDP_CPPTypes.h:6) fresh / resized floating-point storage is NaN.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <initializer_list>
#include <iomanip>
#include <iostream>
#include <limits>
#include <memory>
#include <sstream>
#include <string>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 7
#define CAFE_EIGEN_SHIM 1

namespace Eigen {

typedef std::ptrdiff_t Index;
constexpr int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2, DontAlignCols = 1 };
enum { Lower = 1, Upper = 2 };
constexpr int Infinity = -1;

template <class T> using aligned_allocator = std::allocator<T>;

template <class S, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <class S, int R, int C> class Block;
template <class D> struct traits;
template <class S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC>> { typedef S Scalar; enum { Rows = R, Cols = C }; };
template <class S, int R, int C> struct traits<Block<S, R, C>> { typedef S Scalar; enum { Rows = R, Cols = C }; };

namespace internal {
// contiguous storage with plain references for every scalar type (std::vector<bool> has none)
template <class S> class Buf {
  S* p = nullptr; std::size_t n = 0;
 public:
  Buf() {}
  Buf(std::size_t n_, const S& val) { assign(n_, val); }
  explicit Buf(std::size_t n_) { assign(n_, S()); }
  Buf(std::initializer_list<S> l) { p = l.size() ? new S[l.size()] : nullptr; n = l.size(); std::size_t i = 0; for (const S& x : l) p[i++] = x; }
  Buf(const Buf& o) { p = o.n ? new S[o.n] : nullptr; n = o.n; for (std::size_t i = 0; i < n; ++i) p[i] = o.p[i]; }
  Buf(Buf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  Buf& operator=(const Buf& o) { if (this != &o) { Buf t(o); swap(t); } return *this; }
  Buf& operator=(Buf&& o) noexcept { if (this != &o) { delete[] p; p = o.p; n = o.n; o.p = nullptr; o.n = 0; } return *this; }
  ~Buf() { delete[] p; }
  void assign(std::size_t n_, const S& val) { delete[] p; p = n_ ? new S[n_] : nullptr; n = n_; for (std::size_t i = 0; i < n; ++i) p[i] = val; }
  void swap(Buf& o) { std::swap(p, o.p); std::swap(n, o.n); }
  S* data() { return p; }
  const S* data() const { return p; }
  S& operator[](std::size_t i) { return p[i]; }
  const S& operator[](std::size_t i) const { return p[i]; }
  std::size_t size() const { return n; }
};
template <class S> inline S fresh_value() {
#ifdef EIGEN_INITIALIZE_MATRICES_BY_NAN
  if (std::numeric_limits<S>::has_quiet_NaN) return std::numeric_limits<S>::quiet_NaN();
#endif
  return S();
}
constexpr int prod_dim(int a, int b) { return (a == Dynamic || b == Dynamic) ? Dynamic : a * b; }
constexpr int pick_dim(int a, int b) { return a != Dynamic ? a : b; }
}  // namespace internal

struct IOFormat {
  int precision, flags;
  std::string coeffSep, rowSep, rowPrefix, rowSuffix, matPrefix, matSuffix;
  IOFormat(int p = -1, int f = 0, const std::string& cs = " ", const std::string& rs = "\n", const std::string& rp = "", const std::string& rsuf = "",
           const std::string& mp = "", const std::string& ms = "")
      : precision(p), flags(f), coeffSep(cs), rowSep(rs), rowPrefix(rp), rowSuffix(rsuf), matPrefix(mp), matSuffix(ms) {}
};
constexpr int StreamPrecision = -1, FullPrecision = -2;

template <class S> struct WithFormat {
  std::vector<S> v; Index r, c; IOFormat f;
  friend std::ostream& operator<<(std::ostream& os, const WithFormat& w) {
    std::streamsize old = 0;
    if (w.f.precision > 0) old = os.precision(w.f.precision);
    os << w.f.matPrefix;
    for (Index i = 0; i < w.r; ++i) {
      if (i) os << w.f.rowSep;
      os << w.f.rowPrefix;
      for (Index j = 0; j < w.c; ++j) { if (j) os << w.f.coeffSep; os << w.v[i + w.r * j]; }
      os << w.f.rowSuffix;
    }
    os << w.f.matSuffix;
    if (w.f.precision > 0) os.precision(old);
    return os;
  }
};

template <class S> struct DiagonalWrapper { std::vector<S> d; };

template <class D> class CommaInit;
// the upper / lower triangle (diagonal included) of a matrix as an assignable set of coefficients
template <class B, int Mode> struct TriangularView {
  B m;   // a Block onto the matrix
  explicit TriangularView(const B& b) : m(b) {}
  template <class O> TriangularView& operator=(const TriangularView<O, Mode>& o) {
    for (Index j = 0; j < m.cols(); ++j) for (Index i = 0; i < m.rows(); ++i) if (Mode == Upper ? i <= j : i >= j) m.coeffRef(i, j) = o.m.coeff(i, j);
    return *this;
  }
  TriangularView& operator=(const TriangularView& o) { return this->template operator=<B>(o); }
};
// a 1 x 1 matrix converts to its coefficient (inner products); a plain (non-template) conversion so that the built-in
// operators (double += x^T Q x) see it
template <class D, class S, bool On> struct ScalarConv {};
template <class D, class S> struct ScalarConv<D, S, true> { operator S() const { return static_cast<const D*>(this)->coeff(0, 0); } };

// ------------------------------------------------------------------------------------------------------------------
// DenseBase / MatrixBase: everything that Matrix and Block share. Derived supplies rows(), cols(), ref(i, j).
template <class D>
class DenseBase {
 public:
  typedef typename traits<D>::Scalar Scalar;
  enum { RowsAtCompileTime = traits<D>::Rows, ColsAtCompileTime = traits<D>::Cols,
         SizeAtCompileTime = internal::prod_dim(traits<D>::Rows, traits<D>::Cols), IsVectorAtCompileTime = (traits<D>::Rows == 1 || traits<D>::Cols == 1) };
  typedef Matrix<Scalar, traits<D>::Rows, traits<D>::Cols> PlainObject;
  typedef Block<Scalar, traits<D>::Cols, traits<D>::Rows> TransposeReturnType;   // a strided view: A.transpose() is assignable like Eigen's

  DenseBase() = default;
  DenseBase(const DenseBase&) = default;
  // assignment through a base reference reaches the derived object (const_cast<MatrixBase<T>&>(m) = ..., BarrelRollTO.cpp:44-52)
  DenseBase& operator=(const DenseBase& o) { if (this != &o) derived().assign_(o); return *this; }
  template <class O> D& operator=(const DenseBase<O>& o) { derived().assign_(o); return derived(); }
  D& derived() { return *static_cast<D*>(this); }
  const D& derived() const { return *static_cast<const D*>(this); }
  Index rows() const { return derived().rows_(); }
  Index cols() const { return derived().cols_(); }
  Index size() const { return rows() * cols(); }
  Scalar& coeffRef(Index i, Index j) { return derived().ref(i, j); }
  const Scalar& coeff(Index i, Index j) const { return const_cast<D&>(derived()).ref(i, j); }
  Scalar& operator()(Index i, Index j) { return derived().ref(i, j); }
  const Scalar& operator()(Index i, Index j) const { return coeff(i, j); }
  // linear (vector) access
  Scalar& lin(Index i) { return cols() == 1 ? derived().ref(i, 0) : (rows() == 1 ? derived().ref(0, i) : derived().ref(i % rows(), i / rows())); }
  const Scalar& lin(Index i) const { return const_cast<DenseBase*>(this)->lin(i); }
  Scalar& operator()(Index i) { return lin(i); }
  const Scalar& operator()(Index i) const { return lin(i); }
  Scalar& operator[](Index i) { return lin(i); }
  const Scalar& operator[](Index i) const { return lin(i); }
  Scalar& x() { return lin(0); } Scalar& y() { return lin(1); } Scalar& z() { return lin(2); } Scalar& w() { return lin(3); }
  const Scalar& x() const { return lin(0); } const Scalar& y() const { return lin(1); } const Scalar& z() const { return lin(2); } const Scalar& w() const { return lin(3); }

  PlainObject eval() const { return PlainObject(*this); }

  // ---- fill
  D& setConstant(const Scalar& v) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = v; return derived(); }
  D& setZero() { return setConstant(Scalar(0)); }
  D& setOnes() { return setConstant(Scalar(1)); }
  D& fill(const Scalar& v) { return setConstant(v); }
  D& setIdentity() { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0); return derived(); }

  // ---- views (strided blocks into the same storage)
  template <int BR, int BC> Block<Scalar, BR, BC> mk(Index i, Index j, Index r, Index c) const {
    D& d = const_cast<D&>(derived());
    return Block<Scalar, BR, BC>(d.ptr_(i, j), r, c, d.rs_(), d.cs_());
  }
  Block<Scalar, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const { return mk<Dynamic, Dynamic>(i, j, r, c); }
  template <int BR, int BC> Block<Scalar, BR, BC> block(Index i, Index j) const { return mk<BR, BC>(i, j, BR, BC); }
  template <int BR, int BC> Block<Scalar, BR, BC> block(Index i, Index j, Index r, Index c) const { return mk<BR, BC>(i, j, r, c); }
  Block<Scalar, 1, traits<D>::Cols> row(Index i) const { return mk<1, traits<D>::Cols>(i, 0, 1, cols()); }
  Block<Scalar, traits<D>::Rows, 1> col(Index j) const { return mk<traits<D>::Rows, 1>(0, j, rows(), 1); }
  Block<Scalar, Dynamic, traits<D>::Cols> topRows(Index n) const { return mk<Dynamic, traits<D>::Cols>(0, 0, n, cols()); }
  Block<Scalar, Dynamic, traits<D>::Cols> bottomRows(Index n) const { return mk<Dynamic, traits<D>::Cols>(rows() - n, 0, n, cols()); }
  Block<Scalar, Dynamic, traits<D>::Cols> middleRows(Index i, Index n) const { return mk<Dynamic, traits<D>::Cols>(i, 0, n, cols()); }
  Block<Scalar, traits<D>::Rows, Dynamic> leftCols(Index n) const { return mk<traits<D>::Rows, Dynamic>(0, 0, rows(), n); }
  Block<Scalar, traits<D>::Rows, Dynamic> rightCols(Index n) const { return mk<traits<D>::Rows, Dynamic>(0, cols() - n, rows(), n); }
  Block<Scalar, traits<D>::Rows, Dynamic> middleCols(Index j, Index n) const { return mk<traits<D>::Rows, Dynamic>(0, j, rows(), n); }
  template <int N> Block<Scalar, N, traits<D>::Cols> topRows() const { return mk<N, traits<D>::Cols>(0, 0, N, cols()); }
  template <int N> Block<Scalar, N, traits<D>::Cols> bottomRows() const { return mk<N, traits<D>::Cols>(rows() - N, 0, N, cols()); }
  template <int N> Block<Scalar, N, traits<D>::Cols> middleRows(Index i) const { return mk<N, traits<D>::Cols>(i, 0, N, cols()); }
  template <int N> Block<Scalar, traits<D>::Rows, N> leftCols() const { return mk<traits<D>::Rows, N>(0, 0, rows(), N); }
  template <int N> Block<Scalar, traits<D>::Rows, N> rightCols() const { return mk<traits<D>::Rows, N>(0, cols() - N, rows(), N); }
  template <int N> Block<Scalar, traits<D>::Rows, N> middleCols(Index j) const { return mk<traits<D>::Rows, N>(0, j, rows(), N); }
  template <int BR, int BC> Block<Scalar, BR, BC> topLeftCorner() const { return mk<BR, BC>(0, 0, BR, BC); }
  template <int BR, int BC> Block<Scalar, BR, BC> topRightCorner() const { return mk<BR, BC>(0, cols() - BC, BR, BC); }
  template <int BR, int BC> Block<Scalar, BR, BC> bottomLeftCorner() const { return mk<BR, BC>(rows() - BR, 0, BR, BC); }
  template <int BR, int BC> Block<Scalar, BR, BC> bottomRightCorner() const { return mk<BR, BC>(rows() - BR, cols() - BC, BR, BC); }
  Block<Scalar, Dynamic, Dynamic> topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
  Block<Scalar, Dynamic, Dynamic> topRightCorner(Index r, Index c) const { return block(0, cols() - c, r, c); }
  Block<Scalar, Dynamic, Dynamic> bottomLeftCorner(Index r, Index c) const { return block(rows() - r, 0, r, c); }
  Block<Scalar, Dynamic, Dynamic> bottomRightCorner(Index r, Index c) const { return block(rows() - r, cols() - c, r, c); }
  // vector segments (column or row vector)
  template <int N> Block<Scalar, (traits<D>::Cols == 1 ? N : 1), (traits<D>::Cols == 1 ? 1 : N)> seg_(Index i, Index n) const {
    if (traits<D>::Cols == 1 || cols() == 1) return mk<(traits<D>::Cols == 1 ? N : 1), (traits<D>::Cols == 1 ? 1 : N)>(i, 0, n, 1);
    return mk<(traits<D>::Cols == 1 ? N : 1), (traits<D>::Cols == 1 ? 1 : N)>(0, i, 1, n);
  }
  auto segment(Index i, Index n) const { return seg_<Dynamic>(i, n); }
  template <int N> auto segment(Index i, Index n = N) const { return seg_<N>(i, n); }
  auto head(Index n) const { return seg_<Dynamic>(0, n); }
  template <int N> auto head(Index n = N) const { return seg_<N>(0, n); }
  auto tail(Index n) const { return seg_<Dynamic>(size() - n, n); }
  template <int N> auto tail(Index n = N) const { return seg_<N>(size() - n, n); }
  // diagonal as a strided column view (writable: Q.diagonal() << ...)
  Block<Scalar, Dynamic, 1> diagonal(Index k = 0) const {   // k > 0: super-diagonal, k < 0: sub-diagonal
    D& d = const_cast<D&>(derived());
    const Index i0 = k < 0 ? -k : 0, j0 = k > 0 ? k : 0;
    return Block<Scalar, Dynamic, 1>(d.ptr_(i0, j0), std::min(rows() - i0, cols() - j0), 1, d.rs_() + d.cs_(), 0);
  }

  // ---- reductions
  Scalar sum() const { Scalar s = Scalar(0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s += coeff(i, j); return s; }
  Scalar prod() const { Scalar s = Scalar(1); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s *= coeff(i, j); return s; }
  Scalar trace() const { Scalar s = Scalar(0); for (Index i = 0; i < std::min(rows(), cols()); ++i) s += coeff(i, i); return s; }
  Scalar maxCoeff() const { Scalar s = coeff(0, 0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) > s) s = coeff(i, j); return s; }
  Scalar minCoeff() const { Scalar s = coeff(0, 0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) < s) s = coeff(i, j); return s; }
  Scalar mean() const { return sum() / Scalar(size()); }
  bool any() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j)) return true; return false; }
  bool all() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!coeff(i, j)) return false; return true; }
  Index count() const { Index n = 0; for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j)) ++n; return n; }
  Scalar squaredNorm() const { Scalar s = Scalar(0); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) s += coeff(i, j) * coeff(i, j); return s; }
  Scalar norm() const { using std::sqrt; return sqrt(squaredNorm()); }
  template <int P> Scalar lpNorm() const {
    using std::abs;
    if (P == 2) return norm();
    Scalar s = Scalar(0);
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) { const Scalar a = abs(coeff(i, j)); if (P == 1) s += a; else if (a > s) s = a; }
    return s;
  }
  template <class O> Scalar dot(const DenseBase<O>& o) const { Scalar s = Scalar(0); for (Index i = 0; i < size(); ++i) s += lin(i) * o.lin(i); return s; }
  bool hasNaN() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (coeff(i, j) != coeff(i, j)) return true; return false; }
  bool allFinite() const { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) if (!std::isfinite((double)coeff(i, j))) return false; return true; }

  // ---- coefficient-wise (comparisons yield 0 / 1 integers: std::vector<bool> has no plain references)
  template <class F> auto unary_(F f) const {
    typedef decltype(f(Scalar())) RS;
    Matrix<RS, traits<D>::Rows, traits<D>::Cols> r(rows(), cols());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r(i, j) = f(coeff(i, j));
    return r;
  }
  template <class O, class F> auto binary_(const DenseBase<O>& o, F f) const {
    typedef decltype(f(Scalar(), typename traits<O>::Scalar())) RS;
    assert(rows() == o.rows() && cols() == o.cols());
    Matrix<RS, internal::pick_dim(traits<D>::Rows, traits<O>::Rows), internal::pick_dim(traits<D>::Cols, traits<O>::Cols)> r(rows(), cols());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) r(i, j) = f(coeff(i, j), o.coeff(i, j));
    return r;
  }
  template <class U> auto cast() const { return unary_([](const Scalar& a) { return static_cast<U>(a); }); }
  auto cwiseAbs() const { return unary_([](const Scalar& a) { using std::abs; return (Scalar)abs(a); }); }
  auto cwiseAbs2() const { return unary_([](const Scalar& a) { return (Scalar)(a * a); }); }
  auto cwiseSqrt() const { return unary_([](const Scalar& a) { using std::sqrt; return (Scalar)sqrt(a); }); }
  auto cwiseInverse() const { return unary_([](const Scalar& a) { return (Scalar)(Scalar(1) / a); }); }
  auto cwiseEqual(const Scalar& s) const { return unary_([s](const Scalar& a) { return (int)(a == s); }); }
  auto cwiseNotEqual(const Scalar& s) const { return unary_([s](const Scalar& a) { return (int)(a != s); }); }
  template <class O> auto cwiseEqual(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const typename traits<O>::Scalar& b) { return (int)(a == b); }); }
  template <class O> auto cwiseNotEqual(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const typename traits<O>::Scalar& b) { return (int)(a != b); }); }
  template <class O> auto cwiseProduct(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const Scalar& b) { return (Scalar)(a * b); }); }
  template <class O> auto cwiseQuotient(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const Scalar& b) { return (Scalar)(a / b); }); }
  template <class O> auto cwiseMax(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const Scalar& b) { return a > b ? a : b; }); }
  template <class O> auto cwiseMin(const DenseBase<O>& o) const { return binary_(o, [](const Scalar& a, const Scalar& b) { return a < b ? a : b; }); }

  TransposeReturnType transpose() const {
    D& d = const_cast<D&>(derived());
    return TransposeReturnType(d.ptr_(0, 0), cols(), rows(), d.cs_(), d.rs_());
  }
  template <int Mode> TriangularView<Block<Scalar, traits<D>::Rows, traits<D>::Cols>, Mode> triangularView() const {
    return TriangularView<Block<Scalar, traits<D>::Rows, traits<D>::Cols>, Mode>(mk<traits<D>::Rows, traits<D>::Cols>(0, 0, rows(), cols()));
  }
  template <int RF, int CF> auto replicate() const {
    Matrix<Scalar, internal::prod_dim(traits<D>::Rows, RF), internal::prod_dim(traits<D>::Cols, CF)> r(rows() * RF, cols() * CF);
    for (Index j = 0; j < r.cols(); ++j) for (Index i = 0; i < r.rows(); ++i) r(i, j) = coeff(i % rows(), j % cols());
    return r;
  }
  auto replicate(Index rf, Index cf) const {
    Matrix<Scalar, Dynamic, Dynamic> r(rows() * rf, cols() * cf);
    for (Index j = 0; j < r.cols(); ++j) for (Index i = 0; i < r.rows(); ++i) r(i, j) = coeff(i % rows(), j % cols());
    return r;
  }
  DiagonalWrapper<Scalar> asDiagonal() const { DiagonalWrapper<Scalar> w; w.d.resize(size()); for (Index i = 0; i < size(); ++i) w.d[i] = lin(i); return w; }
  template <class O> Matrix<Scalar, 3, 1> cross(const DenseBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r[0] = lin(1) * o.lin(2) - lin(2) * o.lin(1); r[1] = lin(2) * o.lin(0) - lin(0) * o.lin(2); r[2] = lin(0) * o.lin(1) - lin(1) * o.lin(0);
    return r;
  }
  PlainObject normalized() const { PlainObject r(*this); const Scalar n = norm(); if (n > Scalar(0)) r /= n; return r; }
  // Gauss-Jordan with partial pivoting (small matrices only)
  PlainObject inverse() const {
    const Index n = rows();
    Matrix<Scalar, Dynamic, Dynamic> a(*this), b(n, n);
    b.setIdentity();
    for (Index k = 0; k < n; ++k) {
      Index p = k; using std::abs;
      for (Index i = k + 1; i < n; ++i) if (abs(a(i, k)) > abs(a(p, k))) p = i;
      if (p != k) for (Index j = 0; j < n; ++j) { std::swap(a(k, j), a(p, j)); std::swap(b(k, j), b(p, j)); }
      const Scalar d = a(k, k);
      for (Index j = 0; j < n; ++j) { a(k, j) /= d; b(k, j) /= d; }
      for (Index i = 0; i < n; ++i) if (i != k) { const Scalar f = a(i, k); if (f != Scalar(0)) for (Index j = 0; j < n; ++j) { a(i, j) -= f * a(k, j); b(i, j) -= f * b(k, j); } }
    }
    return PlainObject(b);
  }

  // ---- compound assignment
  template <class O> D& operator+=(const DenseBase<O>& o) { assert(rows() == o.rows() && cols() == o.cols()); PlainObject t(o); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) += t(i, j); return derived(); }
  template <class O> D& operator-=(const DenseBase<O>& o) { assert(rows() == o.rows() && cols() == o.cols()); PlainObject t(o); for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) -= t(i, j); return derived(); }
  D& operator*=(const Scalar& s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) *= s; return derived(); }
  D& operator/=(const Scalar& s) { for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) coeffRef(i, j) /= s; return derived(); }

  // ---- printing / comma initialiser
  WithFormat<Scalar> format(const IOFormat& f) const {
    WithFormat<Scalar> w; w.r = rows(); w.c = cols(); w.f = f; w.v.resize(size());
    for (Index j = 0; j < cols(); ++j) for (Index i = 0; i < rows(); ++i) w.v[i + w.r * j] = coeff(i, j);
    return w;
  }
  CommaInit<D> operator<<(const Scalar& s) { CommaInit<D> c(derived()); c.put(s); return c; }
  template <class O> CommaInit<D> operator<<(const DenseBase<O>& o) { CommaInit<D> c(derived()); c.put(o); return c; }
};
template <class D> using MatrixBase = DenseBase<D>;
template <class D> using EigenBase = DenseBase<D>;
template <class D> using PlainObjectBase = DenseBase<D>;

template <class D>
std::ostream& operator<<(std::ostream& os, const DenseBase<D>& m) {
  for (Index i = 0; i < m.rows(); ++i) { if (i) os << "\n"; for (Index j = 0; j < m.cols(); ++j) { if (j) os << " "; os << m.coeff(i, j); } }
  return os;
}

// fills row by row like Eigen's CommaInitializer; blocks advance by their own width and keep the row band
template <class D>
class CommaInit {
  D& m; Index row = 0, col = 0, band = 1;
 public:
  typedef typename traits<D>::Scalar Scalar;
  explicit CommaInit(D& m_) : m(m_) {}
  void put(const Scalar& s) {
    if (col == m.cols()) { row += band; col = 0; band = 1; }
    assert(row < m.rows());
    m.coeffRef(row, col) = s; ++col; band = 1;
  }
  template <class O> void put(const DenseBase<O>& o) {
    if (o.rows() == 0 || o.cols() == 0) return;
    if (col == m.cols()) { row += band; col = 0; }
    assert(row + o.rows() <= m.rows() && col + o.cols() <= m.cols());
    for (Index j = 0; j < o.cols(); ++j) for (Index i = 0; i < o.rows(); ++i) m.coeffRef(row + i, col + j) = o.coeff(i, j);
    col += o.cols(); band = o.rows();
  }
  CommaInit& operator,(const Scalar& s) { put(s); return *this; }
  template <class O> CommaInit& operator,(const DenseBase<O>& o) { put(o); return *this; }
  D& finished() { return m; }
};

// ------------------------------------------------------------------------------------------------------------------
template <class S, int R, int C, int Opt, int MR, int MC>
class Matrix : public DenseBase<Matrix<S, R, C, Opt, MR, MC>>, public ScalarConv<Matrix<S, R, C, Opt, MR, MC>, S, R == 1 && C == 1> {
  internal::Buf<S> v;
  Index r_, c_;
  typedef DenseBase<Matrix> Base;
 public:
  typedef S Scalar;
  friend class DenseBase<Matrix>;
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  S& ref(Index i, Index j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return v[i + r_ * j]; }
  S* ptr_(Index i, Index j) { return v.data() + i + r_ * j; }
  Index rs_() const { return 1; }
  Index cs_() const { return r_; }

  Matrix() : v((R > 0 ? R : 0) * (C > 0 ? C : 0), internal::fresh_value<S>()), r_(R > 0 ? R : 0), c_(C > 0 ? C : 0) {
    if (R == Dynamic && C == 1) { r_ = 0; c_ = 1; }
    if (R == 1 && C == Dynamic) { r_ = 1; c_ = 0; }
  }
  // Matrix(n) is a size for dynamic vectors; Matrix(r, c) are sizes for dynamic matrices and two coefficients for fixed 2-vectors
  template <class I, typename std::enable_if<std::is_integral<I>::value, int>::type = 0>
  explicit Matrix(I n) : r_(R == Dynamic ? (C == 1 || C == Dynamic ? (Index)n : 1) : R), c_(C == Dynamic ? (R == Dynamic ? 1 : (Index)n) : C) {
    if (R != Dynamic && C != Dynamic) { r_ = R; c_ = C; }
    if (R == Dynamic && C == Dynamic) { r_ = (Index)n; c_ = 1; }
    v.assign(r_ * c_, internal::fresh_value<S>());
    if (R == 1 && C == 1) v[0] = (S)n;
  }
  Matrix(Index r, Index c) : r_(r), c_(c) {
    if (R != Dynamic && C != Dynamic && R * C == 2 && !(R == r && C == c)) { r_ = R; c_ = C; v.assign(2, S()); v[0] = (S)r; v[1] = (S)c; return; }
    v.assign(r_ * c_, internal::fresh_value<S>());
  }
  Matrix(const S& a, const S& b, const S& c) : v{a, b, c}, r_(R == 1 ? 1 : 3), c_(R == 1 ? 3 : 1) {}
  Matrix(const S& a, const S& b, const S& c, const S& d) : v{a, b, c, d}, r_(R == 1 ? 1 : 4), c_(R == 1 ? 4 : 1) {}
  Matrix(const Matrix&) = default;
  Matrix(Matrix&&) = default;
  template <class O> Matrix(const DenseBase<O>& o) : r_(0), c_(0) { assign_(o); }
  Matrix(const DiagonalWrapper<S>& w) : r_(0), c_(0) {
    const Index n = (Index)w.d.size(); r_ = n; c_ = n; v.assign(n * n, S(0)); for (Index i = 0; i < n; ++i) v[i + n * i] = w.d[i];
  }
  Matrix& operator=(const Matrix& o) { if (this != &o) assign_(o); return *this; }
  Matrix& operator=(Matrix&& o) {
    if (this == &o) return *this;
    if ((R == Dynamic || o.r_ == r_) && (C == Dynamic || o.c_ == c_)) { v = std::move(o.v); r_ = o.r_; c_ = o.c_; } else assign_(o);
    return *this;
  }
  template <class O> Matrix& operator=(const DenseBase<O>& o) { assign_(o); return *this; }
  Matrix& operator=(const DiagonalWrapper<S>& w) { *this = Matrix(w); return *this; }

  template <class O> void assign_(const DenseBase<O>& o) {
    Index r = o.rows(), c = o.cols();
    // a vector may be assigned to a vector of the other orientation (Eigen transposes silently for vectors)
    const bool flip = (R == 1 && C != 1 && c == 1 && r != 1) || (C == 1 && R != 1 && r == 1 && c != 1);
    if (flip) std::swap(r, c);
    assert((R == Dynamic || R == r) && (C == Dynamic || C == c));
    if ((const void*)this == (const void*)&o && !flip) return;
    internal::Buf<S> t((size_t)(r * c));
    for (Index j = 0; j < c; ++j) for (Index i = 0; i < r; ++i) t[i + r * j] = (S)(flip ? o.coeff(j, i) : o.coeff(i, j));
    v.swap(t); r_ = r; c_ = c;
  }

  S* data() { return v.data(); }
  const S* data() const { return v.data(); }
  void resize(Index r, Index c) { assert((R == Dynamic || R == r) && (C == Dynamic || C == c)); if (r != r_ || c != c_) { r_ = r; c_ = c; v.assign(r * c, internal::fresh_value<S>()); } }
  void resize(Index n) { if (C == 1 || (R == Dynamic && C == Dynamic)) resize(n, 1); else resize(1, n); }
  void conservativeResize(Index r, Index c) {
    internal::Buf<S> t((size_t)(r * c), internal::fresh_value<S>());
    for (Index j = 0; j < std::min(c, c_); ++j) for (Index i = 0; i < std::min(r, r_); ++i) t[i + r * j] = v[i + r_ * j];
    v.swap(t); r_ = r; c_ = c;
  }
  void conservativeResize(Index n) { if (C == 1 || (R == Dynamic && C == Dynamic)) conservativeResize(n, 1); else conservativeResize(1, n); }
  using Base::setZero; using Base::setOnes; using Base::setConstant; using Base::setIdentity;
  Matrix& setZero(Index n) { resize(n); return Base::setZero(); }
  Matrix& setZero(Index r, Index c) { resize(r, c); return Base::setZero(); }
  Matrix& setOnes(Index n) { resize(n); return Base::setOnes(); }
  Matrix& setOnes(Index r, Index c) { resize(r, c); return Base::setOnes(); }
  Matrix& setConstant(Index n, const S& s) { resize(n); return Base::setConstant(s); }
  Matrix& setConstant(Index r, Index c, const S& s) { resize(r, c); return Base::setConstant(s); }
  Matrix& setIdentity(Index r, Index c) { resize(r, c); return Base::setIdentity(); }

  static Matrix Constant(const S& s) { Matrix m; m.Base::setConstant(s); return m; }
  static Matrix Constant(Index n, const S& s) { Matrix m; m.resize(n); m.Base::setConstant(s); return m; }
  static Matrix Constant(Index r, Index c, const S& s) { Matrix m; m.resize(r, c); m.Base::setConstant(s); return m; }
  static Matrix Zero() { return Constant(S(0)); }
  static Matrix Zero(Index n) { return Constant(n, S(0)); }
  static Matrix Zero(Index r, Index c) { return Constant(r, c, S(0)); }
  static Matrix Ones() { return Constant(S(1)); }
  static Matrix Ones(Index n) { return Constant(n, S(1)); }
  static Matrix Ones(Index r, Index c) { return Constant(r, c, S(1)); }
  static Matrix Identity() { Matrix m; m.Base::setIdentity(); return m; }
  static Matrix Identity(Index r, Index c) { Matrix m; m.resize(r, c); m.Base::setIdentity(); return m; }
};

template <class S, int R, int C>
class Block : public DenseBase<Block<S, R, C>>, public ScalarConv<Block<S, R, C>, S, R == 1 && C == 1> {
  S* p; Index r_, c_, rs, cs;
  typedef DenseBase<Block> Base;
 public:
  typedef S Scalar;
  friend class DenseBase<Block>;
  Block(S* p_, Index r, Index c, Index rs_in, Index cs_in) : p(p_), r_(r), c_(c), rs(rs_in), cs(cs_in) {}
  Block(const Block&) = default;
  Index rows_() const { return r_; }
  Index cols_() const { return c_; }
  S& ref(Index i, Index j) { assert(i >= 0 && i < r_ && j >= 0 && j < c_); return p[i * rs + j * cs]; }
  S* ptr_(Index i, Index j) { return p + i * rs + j * cs; }
  Index rs_() const { return rs; }
  Index cs_() const { return cs; }
  S* data() const { return p; }
  // assignment writes through (the source is evaluated first: aliasing-safe)
  template <class O> void assign_(const DenseBase<O>& o) {
    Index r = o.rows(), c = o.cols();
    const bool flip = (r != r_ || c != c_) && r == c_ && c == r_;
    assert(flip || (r == r_ && c == c_));
    Matrix<S, Dynamic, Dynamic> t(o);
    for (Index j = 0; j < c_; ++j) for (Index i = 0; i < r_; ++i) ref(i, j) = flip ? t(j, i) : t(i, j);
  }
  Block& operator=(const Block& o) { assign_(o); return *this; }
  template <class O> Block& operator=(const DenseBase<O>& o) { assign_(o); return *this; }
  Block& operator=(const DiagonalWrapper<S>& w) { Matrix<S, Dynamic, Dynamic> t(w); assign_(t); return *this; }
};

// Map<Matrix type>: a view of caller-owned contiguous column-major storage
template <class M, int = 0, class = void>
class Map : public Block<typename std::remove_const<typename M::Scalar>::type, traits<typename std::remove_const<M>::type>::Rows, traits<typename std::remove_const<M>::type>::Cols> {
  typedef typename std::remove_const<typename M::Scalar>::type S;
  typedef typename std::remove_const<M>::type MM;
  typedef Block<S, traits<MM>::Rows, traits<MM>::Cols> B;
 public:
  Map(const S* p) : B(const_cast<S*>(p), traits<MM>::Rows, traits<MM>::Cols, 1, traits<MM>::Rows) {}
  Map(const S* p, Index n) : B(const_cast<S*>(p), traits<MM>::Cols == 1 ? n : 1, traits<MM>::Cols == 1 ? 1 : n, 1, traits<MM>::Cols == 1 ? n : 1) {}
  Map(const S* p, Index r, Index c) : B(const_cast<S*>(p), r, c, 1, r) {}
  using B::operator=;
};
template <class M, int X, class Y> struct traits<Map<M, X, Y>> : traits<typename std::remove_const<M>::type> {};

// ---- arithmetic (eager) -------------------------------------------------------------------------------------------
#define CAFE_RES(A, B) Matrix<typename traits<A>::Scalar, internal::pick_dim(traits<A>::Rows, traits<B>::Rows), internal::pick_dim(traits<A>::Cols, traits<B>::Cols)>
template <class A, class B> CAFE_RES(A, B) operator+(const DenseBase<A>& a, const DenseBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  CAFE_RES(A, B) r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <class A, class B> CAFE_RES(A, B) operator-(const DenseBase<A>& a, const DenseBase<B>& b) {
  assert(a.rows() == b.rows() && a.cols() == b.cols());
  CAFE_RES(A, B) r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
#undef CAFE_RES
template <class A> typename DenseBase<A>::PlainObject operator-(const DenseBase<A>& a) {
  typename DenseBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = -a.coeff(i, j);
  return r;
}
// the scalar is a non-deduced parameter: anything convertible to the matrix' scalar type is accepted (int * Matrix<double>)
template <class A> typename DenseBase<A>::PlainObject operator*(const DenseBase<A>& a, const typename traits<A>::Scalar& s) {
  typename DenseBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = a.coeff(i, j) * s;
  return r;
}
template <class A> typename DenseBase<A>::PlainObject operator*(const typename traits<A>::Scalar& s, const DenseBase<A>& a) {
  typename DenseBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = s * a.coeff(i, j);
  return r;
}
template <class A> typename DenseBase<A>::PlainObject operator/(const DenseBase<A>& a, const typename traits<A>::Scalar& s) {
  typename DenseBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = a.coeff(i, j) / s;
  return r;
}
// matrix product: plain inner-product loops, k ascending
template <class A, class B>
Matrix<typename traits<A>::Scalar, traits<A>::Rows, traits<B>::Cols> operator*(const DenseBase<A>& a, const DenseBase<B>& b) {
  typedef typename traits<A>::Scalar S;
  assert(a.cols() == b.rows());
  Matrix<S, traits<A>::Rows, traits<B>::Cols> r(a.rows(), b.cols());
  const Index K = a.cols();
  for (Index j = 0; j < b.cols(); ++j)
    for (Index i = 0; i < a.rows(); ++i) {
      S s = S(0);
      for (Index k = 0; k < K; ++k) s += a.coeff(i, k) * b.coeff(k, j);
      r(i, j) = s;
    }
  return r;
}
template <class S, class B> typename DenseBase<B>::PlainObject operator*(const DiagonalWrapper<S>& d, const DenseBase<B>& b) {
  assert((Index)d.d.size() == b.rows());
  typename DenseBase<B>::PlainObject r(b.rows(), b.cols());
  for (Index j = 0; j < b.cols(); ++j) for (Index i = 0; i < b.rows(); ++i) r(i, j) = d.d[i] * b.coeff(i, j);
  return r;
}
template <class A, class S> typename DenseBase<A>::PlainObject operator*(const DenseBase<A>& a, const DiagonalWrapper<S>& d) {
  assert((Index)d.d.size() == a.cols());
  typename DenseBase<A>::PlainObject r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) r(i, j) = a.coeff(i, j) * d.d[j];
  return r;
}
template <class S> DiagonalWrapper<S> operator*(const typename std::common_type<S>::type& s, const DiagonalWrapper<S>& d) { DiagonalWrapper<S> r(d); for (auto& x : r.d) x = s * x; return r; }
template <class S> DiagonalWrapper<S> operator*(const DiagonalWrapper<S>& d, const typename std::common_type<S>::type& s) { DiagonalWrapper<S> r(d); for (auto& x : r.d) x = x * s; return r; }
template <class A, class B> bool operator==(const DenseBase<A>& a, const DenseBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (Index j = 0; j < a.cols(); ++j) for (Index i = 0; i < a.rows(); ++i) if (!(a.coeff(i, j) == b.coeff(i, j))) return false;
  return true;
}
template <class A, class B> bool operator!=(const DenseBase<A>& a, const DenseBase<B>& b) { return !(a == b); }

// ---- LDLT (Eigen 3.3: Eigen/src/Cholesky/LDLT.h, ldlt_inplace<Lower>::unblocked and LDLT::_solve_impl) -------------
template <class MatrixType, int UpLo = 0>
class LDLT {
  typedef typename MatrixType::Scalar S;
  Matrix<S, Dynamic, Dynamic> m;
  std::vector<Index> tr;
  int sign = 0;   // 0 zero, 1 positive semidefinite, -1 negative semidefinite, 2 indefinite
  bool init = false;
 public:
  LDLT() {}
  template <class O> explicit LDLT(const DenseBase<O>& a) { compute(a); }
  template <class O> LDLT& compute(const DenseBase<O>& a) {
    using std::abs;
    m = a;
    const Index n = m.rows();
    tr.assign(n, 0);
    sign = 0; init = true;
    if (n <= 1) {
      for (Index j = 0; j < n; ++j) tr[j] = j;
      if (n == 1) { const S v = m(0, 0); sign = v > S(0) ? 1 : (v < S(0) ? -1 : 0); }
      return *this;
    }
    std::vector<S> temp(n, S(0));
    for (Index k = 0; k < n; ++k) {
      // largest |diagonal| of the trailing block, first index on ties
      Index piv = k; S best = abs(m(k, k));
      for (Index i = k + 1; i < n; ++i) { const S v = abs(m(i, i)); if (v > best) { best = v; piv = i; } }
      tr[k] = piv;
      if (piv != k) {
        const Index s = n - piv - 1;
        for (Index j = 0; j < k; ++j) std::swap(m(k, j), m(piv, j));
        for (Index i = 0; i < s; ++i) std::swap(m(piv + 1 + i, k), m(piv + 1 + i, piv));
        std::swap(m(k, k), m(piv, piv));
        for (Index i = k + 1; i < piv; ++i) std::swap(m(i, k), m(piv, i));
      }
      const Index rs = n - k - 1;
      if (k > 0) {
        for (Index j = 0; j < k; ++j) temp[j] = m(j, j) * m(k, j);
        S s = S(0);
        for (Index j = 0; j < k; ++j) s += m(k, j) * temp[j];
        m(k, k) -= s;
        for (Index i = 0; i < rs; ++i) { S t = S(0); for (Index j = 0; j < k; ++j) t += m(k + 1 + i, j) * temp[j]; m(k + 1 + i, k) -= t; }
      }
      const S akk = m(k, k);
      const bool pivot_is_valid = abs(akk) > S(0);
      if (k == 0 && !pivot_is_valid) { sign = 0; for (Index j = 0; j < n; ++j) tr[j] = j; return *this; }
      if (rs > 0 && pivot_is_valid) for (Index i = 0; i < rs; ++i) m(k + 1 + i, k) /= akk;
      if (sign == 1) { if (akk < S(0)) sign = 2; }
      else if (sign == -1) { if (akk > S(0)) sign = 2; }
      else if (sign == 0) { if (akk > S(0)) sign = 1; else if (akk < S(0)) sign = -1; }
    }
    return *this;
  }
  bool isPositive() const { return sign == 1 || sign == 0; }
  bool isNegative() const { return sign == -1 || sign == 0; }
  int info() const { return 0; }
  Matrix<S, Dynamic, 1> vectorD() const { return Matrix<S, Dynamic, 1>(m.diagonal()); }
  template <class O> typename DenseBase<O>::PlainObject solve(const DenseBase<O>& b) const {
    using std::abs;
    const Index n = m.rows();
    Matrix<S, Dynamic, Dynamic> X(b);
    const Index nc = X.cols();
    for (Index k = 0; k < n; ++k) if (tr[k] != k) for (Index j = 0; j < nc; ++j) std::swap(X(k, j), X(tr[k], j));
    const S tol = S(1) / std::numeric_limits<S>::max();
    for (Index j = 0; j < nc; ++j) {
      for (Index i = 0; i < n; ++i) { S s = X(i, j); for (Index l = 0; l < i; ++l) s -= m.coeff(i, l) * X(l, j); X(i, j) = s; }
      for (Index i = 0; i < n; ++i) { const S d = m.coeff(i, i); if (abs(d) > tol) X(i, j) /= d; else X(i, j) = S(0); }
      for (Index i = n - 1; i >= 0; --i) { S s = X(i, j); for (Index l = i + 1; l < n; ++l) s -= m.coeff(l, i) * X(l, j); X(i, j) = s; }
    }
    for (Index k = n - 1; k >= 0; --k) if (tr[k] != k) for (Index j = 0; j < nc; ++j) std::swap(X(k, j), X(tr[k], j));
    return typename DenseBase<O>::PlainObject(X);
  }
};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<int, 4, 1> Vector4i;

}  // namespace Eigen
