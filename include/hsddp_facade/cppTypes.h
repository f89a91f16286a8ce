// cppTypes.h — the vector / matrix aliases of the reference (common/cppTypes.h, HSDDPSolver/common/HSDDP_CPPTypes.h:1-60):
// DVec, DMat, VecM, MatMN, Vec3, Vec4 ... With Eigen on the include path they ARE the reference's Eigen typedefs. Without Eigen
// (this repository's build image has none) a minimal column-major stand-in provides what the solver-facing call sequences use
// (testMHPCProblem.cpp:52-89, MHPCLocomotion.cpp:52-150, HKDMPC.cpp:60-140): element access, setZero / setConstant / Zero,
// the comma initialiser, replicate<r,1>(), head / tail / segment copies, fixed <-> dynamic conversion, + - and scalar *.
#pragma once
#include <cstddef>
#if !defined(CAFE_FACADE_NO_EIGEN) && defined(__has_include)
#if __has_include(<Eigen/Dense>)
#define CAFE_FACADE_HAVE_EIGEN 1
#endif
#endif

#ifdef CAFE_FACADE_HAVE_EIGEN
#include <Eigen/Dense>
template <typename T> using DVec = Eigen::Matrix<T, Eigen::Dynamic, 1>;
template <typename T> using DMat = Eigen::Matrix<T, Eigen::Dynamic, Eigen::Dynamic>;
template <typename T, size_t N> using VecM = Eigen::Matrix<T, (int)N, 1>;
template <typename T, size_t M, size_t N> using MatMN = Eigen::Matrix<T, (int)M, (int)N>;
template <typename T> using Vec2 = Eigen::Matrix<T, 2, 1>;
template <typename T> using Vec3 = Eigen::Matrix<T, 3, 1>;
template <typename T> using Vec4 = Eigen::Matrix<T, 4, 1>;
template <typename T> using Vec12 = Eigen::Matrix<T, 12, 1>;
template <typename T> using Mat3 = Eigen::Matrix<T, 3, 3>;
#else
#include <cassert>
#include <vector>
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
namespace cafe_shim {
constexpr int Dynamic = -1;
template <typename T, int R, int C>
class Matrix {
 public:
  typedef T Scalar;
  Matrix() : r_(R < 0 ? 0 : R), c_(C < 0 ? 0 : C), v_((size_t)r_ * c_, T(0)) {}
  explicit Matrix(int n) : r_(R < 0 ? n : R), c_(C < 0 ? 1 : C), v_((size_t)r_ * c_, T(0)) {}
  Matrix(int r, int c) : r_(r), c_(c), v_((size_t)r * c, T(0)) { static_assert(R < 0 && C < 0, "size constructor of a dynamic matrix"); }
  Matrix(T a, T b, T c) : r_(3), c_(1), v_{a, b, c} { static_assert(R == 3 && C == 1, "three-coefficient constructor of Vec3"); }
  Matrix(T a, T b, T c, T d) : r_(4), c_(1), v_{a, b, c, d} { static_assert(R == 4 && C == 1, "four-coefficient constructor of Vec4"); }
  template <typename T2, int R2, int C2>
  Matrix(const Matrix<T2, R2, C2>& o) : r_(o.rows()), c_(o.cols()), v_((size_t)o.size()) {
    assert((R < 0 || R == o.rows()) && (C < 0 || C == o.cols()));
    for (int i = 0; i < size(); ++i) v_[i] = (T)o.data()[i];
  }
  int rows() const { return r_; }
  int cols() const { return c_; }
  int size() const { return r_ * c_; }
  T* data() { return v_.data(); }
  const T* data() const { return v_.data(); }
  T& operator[](int i) { return v_[i]; }
  const T& operator[](int i) const { return v_[i]; }
  T& operator()(int i) { return v_[i]; }
  const T& operator()(int i) const { return v_[i]; }
  T& operator()(int i, int j) { return v_[i + (size_t)r_ * j]; }
  const T& operator()(int i, int j) const { return v_[i + (size_t)r_ * j]; }
  Matrix& setZero() { for (auto& x : v_) x = T(0); return *this; }
  Matrix& setZero(int n) { resize(n, C < 0 ? 1 : C); return setZero(); }
  Matrix& setZero(int r, int c) { resize(r, c); return setZero(); }
  Matrix& setConstant(T s) { for (auto& x : v_) x = s; return *this; }
  Matrix& setOnes() { return setConstant(T(1)); }
  void resize(int r, int c) { r_ = r; c_ = c; v_.assign((size_t)r * c, T(0)); }
  void resize(int n) { resize(n, 1); }
  static Matrix Zero() { return Matrix(); }
  static Matrix Zero(int n) { Matrix m(n); return m; }
  static Matrix Zero(int r, int c) { Matrix m; m.resize(r, c); return m; }
  static Matrix Constant(T s) { Matrix m; m.setConstant(s); return m; }
  template <typename T2> Matrix<T2, R, C> cast() const { Matrix<T2, R, C> o; o.resize(r_, c_); for (int i = 0; i < size(); ++i) o.data()[i] = (T2)v_[i]; return o; }
  template <int RR, int CC> Matrix<T, (R < 0 ? -1 : R * RR), (C < 0 ? -1 : C * CC)> replicate() const {
    Matrix<T, (R < 0 ? -1 : R * RR), (C < 0 ? -1 : C * CC)> o; o.resize(r_ * RR, c_ * CC);
    for (int bj = 0; bj < CC; ++bj) for (int bi = 0; bi < RR; ++bi) for (int j = 0; j < c_; ++j) for (int i = 0; i < r_; ++i) o(bi * r_ + i, bj * c_ + j) = (*this)(i, j);
    return o;
  }
  Matrix<T, Dynamic, 1> segment(int i0, int n) const { Matrix<T, Dynamic, 1> o(n); for (int i = 0; i < n; ++i) o[i] = v_[i0 + i]; return o; }
  Matrix<T, Dynamic, 1> head(int n) const { return segment(0, n); }
  Matrix<T, Dynamic, 1> tail(int n) const { return segment(size() - n, n); }
  template <int N> Matrix<T, N, 1> segment(int i0) const { Matrix<T, N, 1> o; for (int i = 0; i < N; ++i) o[i] = v_[i0 + i]; return o; }
  template <int N> Matrix<T, N, 1> head() const { return segment<N>(0); }
  template <int N> Matrix<T, N, 1> tail() const { return segment<N>(size() - N); }
  T norm() const { T s = 0; for (auto x : v_) s += x * x; return std::sqrt(s); }
  Matrix& operator+=(const Matrix& o) { for (int i = 0; i < size(); ++i) v_[i] += o.v_[i]; return *this; }
  Matrix& operator-=(const Matrix& o) { for (int i = 0; i < size(); ++i) v_[i] -= o.v_[i]; return *this; }
  Matrix& operator*=(T s) { for (auto& x : v_) x *= s; return *this; }
  friend Matrix operator+(Matrix a, const Matrix& b) { return a += b; }
  friend Matrix operator-(Matrix a, const Matrix& b) { return a -= b; }
  friend Matrix operator*(Matrix a, T s) { return a *= s; }
  friend Matrix operator*(T s, Matrix a) { return a *= s; }
  // comma initialiser: m << a, b, c;  (scalars and vectors, filled in storage order: what the call sites do with column vectors)
  struct Comma {
    Matrix& m; int at;
    Comma& operator,(T s) { assert(at < m.size()); m.v_[at++] = s; return *this; }
    template <int R2, int C2> Comma& operator,(const Matrix<T, R2, C2>& o) { for (int i = 0; i < o.size(); ++i) { assert(at < m.size()); m.v_[at++] = o.data()[i]; } return *this; }
  };
  Comma operator<<(T s) { Comma c{*this, 0}; c, s; return c; }
  template <int R2, int C2> Comma operator<<(const Matrix<T, R2, C2>& o) { Comma c{*this, 0}; c, o; return c; }

 private:
  int r_, c_;
  std::vector<T> v_;
};
}  // namespace cafe_shim
#include <cmath>
template <typename T> using DVec = cafe_shim::Matrix<T, -1, 1>;
template <typename T> using DMat = cafe_shim::Matrix<T, -1, -1>;
template <typename T, size_t N> using VecM = cafe_shim::Matrix<T, (int)N, 1>;
template <typename T, size_t M, size_t N> using MatMN = cafe_shim::Matrix<T, (int)M, (int)N>;
template <typename T> using Vec2 = cafe_shim::Matrix<T, 2, 1>;
template <typename T> using Vec3 = cafe_shim::Matrix<T, 3, 1>;
template <typename T> using Vec4 = cafe_shim::Matrix<T, 4, 1>;
template <typename T> using Vec12 = cafe_shim::Matrix<T, 12, 1>;
template <typename T> using Mat3 = cafe_shim::Matrix<T, 3, 3>;
#endif
