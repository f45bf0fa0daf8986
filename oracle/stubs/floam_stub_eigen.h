// ORACLE — TEST INFRASTRUCTURE ONLY.  Stand-in for the subset of Eigen 3.3 that the reference's class sources use, so that
// /root/reference/src/{laserProcessingClass,dataHandler,lidar,lidarOptimization,odomEstimationClass,laserMappingClass}.cpp
// compile UNMODIFIED into oracle/_ref (oracle/Makefile).  Eigen itself is an un-vendored dependency of the reference
// (CMakeLists.txt:21 `find_package(Eigen3)`, implied 3.3.4) and is absent from this image.
//
// Everything is evaluated eagerly, coefficient by coefficient, in the order of Eigen's coefficient-based (lazy) product /
// cwise evaluators: sums over the inner index run k = 0,1,2,... left to right.  The iterative routines
// (SelfAdjointEigenSolver, ColPivHouseholderQR, quaternion <-> matrix) forward to oracle/linalg.h, which follows the published
// Eigen 3.3.4 sources.  Results can differ from a real Eigen build in the last ulp where Eigen vectorises or re-associates;
// integer/byte-exact parts of the path (feature extraction) do not depend on this header's arithmetic at all.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <memory>
#include <ostream>
#include <sstream>
#include <type_traits>
#include <vector>
#include "../linalg.h"

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 __attribute__((aligned(16)))
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3

namespace Eigen {

enum { ColMajor = 0, RowMajor = 1 };
const int Dynamic = -1;
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };

template <class T>
class aligned_allocator : public std::allocator<T> {
 public:
  template <class U> struct rebind { typedef aligned_allocator<U> other; };
  aligned_allocator() {}
  template <class U> aligned_allocator(const aligned_allocator<U>&) {}
};

template <class S, int R, int C, int Opt = ColMajor> class Matrix;
template <class S, int R, int C> class View;
template <class S> class DynView;
template <class T> class Map;
template <class S> class Quaternion;
template <class S> class AngleAxis;
template <class S, int Dim, int Mode> class Transform;

// ---------------------------------------------------------------------------------------------------------------------
// DenseBase: CRTP base of Matrix / View / Map.  Derived supplies at(i,j) (and a mutable at for writable types).
template <class D, class S, int R, int C>
struct DenseBase {
  typedef S Scalar;
  enum { RowsAtCompileTime = R, ColsAtCompileTime = C, SizeAtCompileTime = R * C };
  typedef Matrix<S, R, C> PlainObject;
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  int rows() const { return R; }
  int cols() const { return C; }
  int size() const { return R * C; }

  S operator()(int i, int j) const { return derived().at(i, j); }
  S& operator()(int i, int j) { return derived().at(i, j); }
  S operator()(int i) const { return lin(i); }
  S& operator()(int i) { return lin(i); }
  S operator[](int i) const { return lin(i); }
  S& operator[](int i) { return lin(i); }
  S coeff(int i) const { return lin(i); }
  S& coeffRef(int i) { return lin(i); }
  S coeff(int i, int j) const { return derived().at(i, j); }
  S& coeffRef(int i, int j) { return derived().at(i, j); }
  S x() const { return lin(0); }
  S y() const { return lin(1); }
  S z() const { return lin(2); }
  S w() const { return lin(3); }
  S& x() { return lin(0); }
  S& y() { return lin(1); }
  S& z() { return lin(2); }
  S& w() { return lin(3); }

  PlainObject eval() const {
    PlainObject r;
    for (int j = 0; j < C; ++j)
      for (int i = 0; i < R; ++i) r.at(i, j) = derived().at(i, j);
    return r;
  }
  Matrix<S, C, R> transpose() const {
    Matrix<S, C, R> r;
    for (int j = 0; j < C; ++j)
      for (int i = 0; i < R; ++i) r.at(j, i) = derived().at(i, j);
    return r;
  }
  S squaredNorm() const {  // linear order, left to right
    S s = lin(0) * lin(0);
    for (int k = 1; k < R * C; ++k) s += lin(k) * lin(k);
    return s;
  }
  S norm() const { return std::sqrt(squaredNorm()); }
  S sum() const {
    S s = lin(0);
    for (int k = 1; k < R * C; ++k) s += lin(k);
    return s;
  }
  template <class O>
  S dot(const DenseBase<O, S, R, C>& o) const {
    S s = lin(0) * o.derived().lin_c(0);
    for (int k = 1; k < R * C; ++k) s += lin(k) * o.derived().lin_c(k);
    return s;
  }
  template <class O>
  Matrix<S, 3, 1> cross(const DenseBase<O, S, R, C>& o) const {  // MatrixBase::cross (3-vectors)
    static_assert(R * C == 3, "cross: 3-vectors only");
    const O& b = o.derived();
    Matrix<S, 3, 1> r;
    r.at(0, 0) = lin(1) * b.lin_c(2) - lin(2) * b.lin_c(1);
    r.at(1, 0) = lin(2) * b.lin_c(0) - lin(0) * b.lin_c(2);
    r.at(2, 0) = lin(0) * b.lin_c(1) - lin(1) * b.lin_c(0);
    return r;
  }
  PlainObject normalized() const {
    const S n = norm();
    PlainObject r = eval();
    if (n > S(0)) for (int k = 0; k < R * C; ++k) r.lin(k) = r.lin(k) / n;
    return r;
  }
  void normalize() {  // MatrixBase::normalize: *this /= norm() when squaredNorm() > 0
    const S z2 = squaredNorm();
    if (z2 > S(0)) { const S n = std::sqrt(z2); for (int k = 0; k < R * C; ++k) lin(k) = lin(k) / n; }
  }
  D& setZero() {
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = S(0);
    return derived();
  }
  D& setIdentity() {
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = (i == j) ? S(1) : S(0);
    return derived();
  }
  D& setOnes() {
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = S(1);
    return derived();
  }
  template <class T>
  Matrix<T, R, C> cast() const {
    Matrix<T, R, C> r;
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = static_cast<T>(derived().at(i, j));
    return r;
  }
  Matrix<S, R, 1> col(int j) const {
    Matrix<S, R, 1> r;
    for (int i = 0; i < R; ++i) r.at(i, 0) = derived().at(i, j);
    return r;
  }
  Matrix<S, 1, C> row(int i) const {
    Matrix<S, 1, C> r;
    for (int j = 0; j < C; ++j) r.at(0, j) = derived().at(i, j);
    return r;
  }
  template <int BR, int BC>
  Matrix<S, BR, BC> block(int i0, int j0) const {
    Matrix<S, BR, BC> r;
    for (int j = 0; j < BC; ++j) for (int i = 0; i < BR; ++i) r.at(i, j) = derived().at(i0 + i, j0 + j);
    return r;
  }
  template <int BR, int BC>
  View<S, BR, BC> block(int i0, int j0) { return View<S, BR, BC>(&derived().at(i0, j0), derived().row_stride(), derived().col_stride()); }
  DynView<S> topRows(int n) { return DynView<S>(&derived().at(0, 0), n, C, derived().row_stride(), derived().col_stride()); }
  DynView<S> bottomRows(int n) { return DynView<S>(&derived().at(R - n, 0), n, C, derived().row_stride(), derived().col_stride()); }
  PlainObject matrix() const { return eval(); }
  // `m << a, b, c, ...;` fills coefficients row by row (Eigen's CommaInitializer)
  struct CommaInit {
    D& m; int k;
    CommaInit& operator,(S v) { m.at(k / C, k % C) = v; ++k; return *this; }
  };
  CommaInit operator<<(S v) { derived().at(0, 0) = v; return CommaInit{derived(), 1}; }
  struct DiagonalWrapper { PlainObject v; };
  DiagonalWrapper asDiagonal() const { return DiagonalWrapper{eval()}; }

  template <class O> D& operator+=(const DenseBase<O, S, R, C>& o) {
    PlainObject t = o.eval();
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = derived().at(i, j) + t.at(i, j);
    return derived();
  }
  template <class O> D& operator-=(const DenseBase<O, S, R, C>& o) {
    PlainObject t = o.eval();
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = derived().at(i, j) - t.at(i, j);
    return derived();
  }
  D& operator*=(S s) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = derived().at(i, j) * s; return derived(); }
  D& operator/=(S s) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) derived().at(i, j) = derived().at(i, j) / s; return derived(); }

  // linear (vector-style) coefficient access: index k walks a vector, or column-major order of a matrix
  S lin_c(int k) const { return (C == 1) ? derived().at(k, 0) : (R == 1) ? derived().at(0, k) : derived().at(k % R, k / R); }
  S lin(int k) const { return lin_c(k); }
  S& lin(int k) { return (C == 1) ? derived().at(k, 0) : (R == 1) ? derived().at(0, k) : derived().at(k % R, k / R); }
};

template <class D, class S, int R, int C>
void assign_dense(D& dst, const Matrix<S, R, C>& src) {
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) dst.at(i, j) = src.at(i, j);
}

// ---------------------------------------------------------------------------------------------------------------------
template <class S, int R, int C, int Opt>
class Matrix : public DenseBase<Matrix<S, R, C, Opt>, S, R, C> {
  typedef DenseBase<Matrix<S, R, C, Opt>, S, R, C> Base;

 public:
  enum { IsRowMajor = (Opt & RowMajor) ? 1 : 0 };
  Matrix() {}
  // vector constructors (any arithmetic argument types, converted like Eigen's Scalar conversions)
  template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
  Matrix(A a, B b) { static_assert(R * C == 2, "2 coefficients"); d_[0] = S(a); d_[1] = S(b); }
  template <class A, class B, class E, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<E>::value>::type>
  Matrix(A a, B b, E c) { static_assert(R * C == 3, "3 coefficients"); d_[0] = S(a); d_[1] = S(b); d_[2] = S(c); }
  template <class A, class B, class E, class F,
            class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<E>::value && std::is_arithmetic<F>::value>::type>
  Matrix(A a, B b, E c, F e) { static_assert(R * C == 4, "4 coefficients"); d_[0] = S(a); d_[1] = S(b); d_[2] = S(c); d_[3] = S(e); }
  explicit Matrix(const S* p) { for (int k = 0; k < R * C; ++k) d_[k] = p[k]; }
  template <class O>
  Matrix(const DenseBase<O, S, R, C>& o) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) at(i, j) = o.derived().at(i, j); }
  template <class O>
  Matrix& operator=(const DenseBase<O, S, R, C>& o) {
    Matrix<S, R, C> t = o.eval();
    for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) at(i, j) = t.at(i, j);
    return *this;
  }

  S at(int i, int j) const { return d_[IsRowMajor ? i * C + j : j * R + i]; }
  S& at(int i, int j) { return d_[IsRowMajor ? i * C + j : j * R + i]; }
  std::ptrdiff_t row_stride() const { return IsRowMajor ? C : 1; }
  std::ptrdiff_t col_stride() const { return IsRowMajor ? 1 : R; }
  const S* data() const { return d_; }
  S* data() { return d_; }

  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Ones() { Matrix m; m.setOnes(); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Constant(S v) { Matrix m; for (int k = 0; k < R * C; ++k) m.d_[k] = v; return m; }
  static Matrix Unit(int k) { Matrix m; m.setZero(); m.d_[k] = S(1); return m; }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }
  static Matrix UnitW() { return Unit(3); }

  // decompositions the reference calls (declared below)
  struct ColPivQrProxy {
    Matrix<S, R, C> a;
    template <class B> Matrix<S, C, 1> solve(const DenseBase<B, S, R, 1>& b) const {
      double A[R * C], bb[R], x[3];
      static_assert(C == 3, "colPivHouseholderQr().solve: rows x 3 systems only (src/odomEstimationClass.cpp:220)");
      for (int i = 0; i < R; ++i) { for (int j = 0; j < C; ++j) A[i * C + j] = a.at(i, j); bb[i] = b.derived().at(i, 0); }
      fo::colpiv_qr_solve_nx3(A, bb, R, x);
      return Matrix<S, C, 1>(x[0], x[1], x[2]);
    }
  };
  ColPivQrProxy colPivHouseholderQr() const { return ColPivQrProxy{this->eval()}; }

 private:
  S d_[R * C];
};

// dynamic-size matrix: only what src/odomEstimationNode.cpp:106-112 needs (`const Eigen::MatrixXd m = poses[i].matrix(); m(r,c)`)
template <class S, int Opt>
class Matrix<S, Dynamic, Dynamic, Opt> {
 public:
  Matrix() : r_(0), c_(0) {}
  template <int R, int C, int O>
  Matrix(const Matrix<S, R, C, O>& o) : r_(R), c_(C), d_(R * C) { for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) d_[j * R + i] = o.at(i, j); }
  template <class W, class = decltype(W::v)>
  Matrix(const W& diag) : r_(diag.v.size()), c_(diag.v.size()), d_((size_t)diag.v.size() * diag.v.size(), S(0)) {   // vec.asDiagonal()
    for (int i = 0; i < r_; ++i) d_[i * r_ + i] = diag.v.lin_c(i);
  }
  S operator()(int i, int j) const { return d_[j * r_ + i]; }
  S& operator()(int i, int j) { return d_[j * r_ + i]; }
  int rows() const { return r_; }
  int cols() const { return c_; }

 private:
  int r_, c_;
  std::vector<S> d_;
};

// strided, writable window on somebody else's storage (block<>(), linear(), translation(), Map)
template <class S, int R, int C>
class View : public DenseBase<View<S, R, C>, S, R, C> {
 public:
  View(S* p, std::ptrdiff_t rs, std::ptrdiff_t cs) : p_(p), rs_(rs), cs_(cs) {}
  View(const View& o) : p_(o.p_), rs_(o.rs_), cs_(o.cs_) {}
  S at(int i, int j) const { return p_[i * rs_ + j * cs_]; }
  S& at(int i, int j) { return p_[i * rs_ + j * cs_]; }
  std::ptrdiff_t row_stride() const { return rs_; }
  std::ptrdiff_t col_stride() const { return cs_; }
  S* data() { return p_; }
  const S* data() const { return p_; }
  View& operator=(const View& o) { Matrix<S, R, C> t = o.eval(); assign_dense(*this, t); return *this; }  // copies coefficients
  template <class O>
  View& operator=(const DenseBase<O, S, R, C>& o) { Matrix<S, R, C> t = o.eval(); assign_dense(*this, t); return *this; }

 private:
  S* p_;
  std::ptrdiff_t rs_, cs_;
};

// run-time sized window: topRows(n) / bottomRows(n).setIdentity() / setZero() (src/lidarOptimization.cpp:95-97)
template <class S>
class DynView {
 public:
  DynView(S* p, int r, int c, std::ptrdiff_t rs, std::ptrdiff_t cs) : p_(p), r_(r), c_(c), rs_(rs), cs_(cs) {}
  DynView& setIdentity() { for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) p_[i * rs_ + j * cs_] = (i == j) ? S(1) : S(0); return *this; }
  DynView& setZero() { for (int i = 0; i < r_; ++i) for (int j = 0; j < c_; ++j) p_[i * rs_ + j * cs_] = S(0); return *this; }
  S& operator()(int i, int j) { return p_[i * rs_ + j * cs_]; }
  int rows() const { return r_; }
  int cols() const { return c_; }

 private:
  S* p_;
  int r_, c_;
  std::ptrdiff_t rs_, cs_;
};

// Map<Matrix<...>> and Map<const Matrix<...>>
template <class S, int R, int C, int Opt>
class Map<Matrix<S, R, C, Opt>> : public DenseBase<Map<Matrix<S, R, C, Opt>>, S, R, C> {
  enum { RM = (Opt & RowMajor) ? 1 : 0 };

 public:
  explicit Map(S* p) : p_(p) {}
  Map(const Map& o) : p_(o.p_) {}  // Eigen: copying a Map copies the pointer ...
  S at(int i, int j) const { return p_[RM ? i * C + j : j * R + i]; }
  S& at(int i, int j) { return p_[RM ? i * C + j : j * R + i]; }
  std::ptrdiff_t row_stride() const { return RM ? C : 1; }
  std::ptrdiff_t col_stride() const { return RM ? 1 : R; }
  S* data() { return p_; }
  const S* data() const { return p_; }
  Map& operator=(const Map& o) { Matrix<S, R, C> t = o.eval(); assign_dense(*this, t); return *this; }  // ... assigning copies coefficients
  template <class O>
  Map& operator=(const DenseBase<O, S, R, C>& o) { Matrix<S, R, C> t = o.eval(); assign_dense(*this, t); return *this; }

 private:
  S* p_;
};
template <class S, int R, int C, int Opt>
class Map<const Matrix<S, R, C, Opt>> : public DenseBase<Map<const Matrix<S, R, C, Opt>>, S, R, C> {
  enum { RM = (Opt & RowMajor) ? 1 : 0 };

 public:
  explicit Map(const S* p) : p_(p) {}
  S at(int i, int j) const { return p_[RM ? i * C + j : j * R + i]; }
  S& at(int i, int j) { return const_cast<S&>(p_[RM ? i * C + j : j * R + i]); }  // never written through (const Map)
  std::ptrdiff_t row_stride() const { return RM ? C : 1; }
  std::ptrdiff_t col_stride() const { return RM ? 1 : R; }
  const S* data() const { return p_; }

 private:
  const S* p_;
};

// ---------------------------------------------------------------------------------------------------------------------
// cwise operators and products (eager)
template <class A, class B, class S, int R, int C>
Matrix<S, R, C> operator+(const DenseBase<A, S, R, C>& a, const DenseBase<B, S, R, C>& b) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = a.derived().at(i, j) + b.derived().at(i, j);
  return r;
}
template <class A, class B, class S, int R, int C>
Matrix<S, R, C> operator-(const DenseBase<A, S, R, C>& a, const DenseBase<B, S, R, C>& b) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = a.derived().at(i, j) - b.derived().at(i, j);
  return r;
}
template <class A, class S, int R, int C>
Matrix<S, R, C> operator-(const DenseBase<A, S, R, C>& a) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = -a.derived().at(i, j);
  return r;
}
template <class A, class S, int R, int C, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<S, R, C> operator*(const DenseBase<A, S, R, C>& a, T s) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = a.derived().at(i, j) * S(s);
  return r;
}
template <class A, class S, int R, int C, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<S, R, C> operator*(T s, const DenseBase<A, S, R, C>& a) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = S(s) * a.derived().at(i, j);
  return r;
}
template <class A, class S, int R, int C, class T, class = typename std::enable_if<std::is_arithmetic<T>::value>::type>
Matrix<S, R, C> operator/(const DenseBase<A, S, R, C>& a, T s) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j) for (int i = 0; i < R; ++i) r.at(i, j) = a.derived().at(i, j) / S(s);
  return r;
}
// coefficient-based product: res(i,j) = sum_k lhs(i,k)*rhs(k,j), k ascending
template <class A, class B, class S, int R, int K, int C>
Matrix<S, R, C> operator*(const DenseBase<A, S, R, K>& a, const DenseBase<B, S, K, C>& b) {
  Matrix<S, R, C> r;
  for (int j = 0; j < C; ++j)
    for (int i = 0; i < R; ++i) {
      S s = a.derived().at(i, 0) * b.derived().at(0, j);
      for (int k = 1; k < K; ++k) s += a.derived().at(i, k) * b.derived().at(k, j);
      r.at(i, j) = s;
    }
  return r;
}

// operator<<(ostream, matrix) with Eigen's default IOFormat: stream precision, columns aligned to the widest coefficient, coefficients
// separated by one blank, rows by a newline (Eigen/src/Core/IO.h, print_matrix)
template <class A, class S, int R, int C>
std::ostream& operator<<(std::ostream& s, const DenseBase<A, S, R, C>& m) {
  std::streamsize width = 0;
  for (int j = 0; j < C; ++j)
    for (int i = 0; i < R; ++i) {
      std::stringstream sstr;
      sstr.copyfmt(s);
      sstr << m.derived().at(i, j);
      width = std::max<std::streamsize>(width, (std::streamsize)sstr.str().length());
    }
  for (int i = 0; i < R; ++i) {
    if (width) s.width(width);
    s << m.derived().at(i, 0);
    for (int j = 1; j < C; ++j) {
      s << " ";
      if (width) s.width(width);
      s << m.derived().at(i, j);
    }
    if (i < R - 1) s << "\n";
  }
  return s;
}

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;

// ---------------------------------------------------------------------------------------------------------------------
// Quaternion (coefficients stored x,y,z,w like Eigen's coeffs())
namespace internal {
template <class S> inline fo::Quat to_fo(const S* c) { return fo::Quat{double(c[0]), double(c[1]), double(c[2]), double(c[3])}; }
template <class M> inline fo::Mat3 mat3_to_fo(const M& m) {
  fo::Mat3 r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.m[i][j] = double(m.at(i, j));
  return r;
}
template <class S> inline Matrix<S, 3, 3> mat3_from_fo(const fo::Mat3& m) {
  Matrix<S, 3, 3> r;
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) r.at(i, j) = S(m.m[i][j]);
  return r;
}
}  // namespace internal

template <class D, class S>
struct QuaternionBase {
  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  const S* c() const { return derived().coeff_ptr(); }
  S x() const { return c()[0]; }
  S y() const { return c()[1]; }
  S z() const { return c()[2]; }
  S w() const { return c()[3]; }
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(c()[0], c()[1], c()[2]); }
  Matrix<S, 4, 1> coeffs() const { return Matrix<S, 4, 1>(c()[0], c()[1], c()[2], c()[3]); }
  S squaredNorm() const { return c()[0] * c()[0] + c()[1] * c()[1] + c()[2] * c()[2] + c()[3] * c()[3]; }
  S norm() const { return std::sqrt(squaredNorm()); }
  template <class O>
  Quaternion<S> operator*(const QuaternionBase<O, S>& o) const {  // internal::quat_product (scalar path)
    const S* a = c(); const S* b = o.c();
    return Quaternion<S>(a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2], a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1],
                         a[3] * b[1] + a[1] * b[3] + a[2] * b[0] - a[0] * b[2], a[3] * b[2] + a[2] * b[3] + a[0] * b[1] - a[1] * b[0]);
  }
  Quaternion<S> operator*(const AngleAxis<S>& aa) const { return *this * Quaternion<S>(aa); }
  template <class V>
  Matrix<S, 3, 1> operator*(const DenseBase<V, S, 3, 1>& v) const {  // QuaternionBase::_transformVector
    const Matrix<S, 3, 1> qv = vec();
    Matrix<S, 3, 1> uv = qv.cross(v);
    uv += uv;
    return v + w() * uv + qv.cross(uv);
  }
  Quaternion<S> conjugate() const { return Quaternion<S>(c()[3], -c()[0], -c()[1], -c()[2]); }
  Quaternion<S> inverse() const {  // QuaternionBase::inverse
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion<S>(c()[3] / n2, -c()[0] / n2, -c()[1] / n2, -c()[2] / n2);
    return Quaternion<S>(S(0), S(0), S(0), S(0));
  }
  Quaternion<S> normalized() const { const S n = norm(); return Quaternion<S>(c()[3] / n, c()[0] / n, c()[1] / n, c()[2] / n); }
  Matrix<S, 3, 3> toRotationMatrix() const { return internal::mat3_from_fo<S>(fo::quat_to_matrix(internal::to_fo(c()))); }
  Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
};

template <class S>
class Quaternion : public QuaternionBase<Quaternion<S>, S> {
 public:
  Quaternion() {}
  Quaternion(S w, S x, S y, S z) { q_[0] = x; q_[1] = y; q_[2] = z; q_[3] = w; }
  template <class O> Quaternion(const QuaternionBase<O, S>& o) { std::memcpy(q_, o.c(), sizeof(q_)); }
  explicit Quaternion(const AngleAxis<S>& aa) {  // QuaternionBase::operator=(AngleAxis)
    const S ha = S(0.5) * aa.angle();
    q_[3] = std::cos(ha);
    const S s = std::sin(ha);
    q_[0] = s * aa.axis()(0); q_[1] = s * aa.axis()(1); q_[2] = s * aa.axis()(2);
  }
  template <class M>
  explicit Quaternion(const DenseBase<M, S, 3, 3>& m) {  // quaternionbase_assign_impl<Other,3,3>
    const fo::Quat q = fo::quat_from_matrix(internal::mat3_to_fo(m.derived()));
    q_[0] = S(q.x); q_[1] = S(q.y); q_[2] = S(q.z); q_[3] = S(q.w);
  }
  template <class O> Quaternion& operator=(const QuaternionBase<O, S>& o) { S t[4]; std::memcpy(t, o.c(), sizeof(t)); std::memcpy(q_, t, sizeof(t)); return *this; }
  Quaternion& operator=(const AngleAxis<S>& aa) { *this = Quaternion(aa); return *this; }
  static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
  void normalize() { const S n = this->norm(); for (int k = 0; k < 4; ++k) q_[k] = q_[k] / n; }
  const S* coeff_ptr() const { return q_; }
  S* coeff_ptr() { return q_; }
  S& x() { return q_[0]; }
  S& y() { return q_[1]; }
  S& z() { return q_[2]; }
  S& w() { return q_[3]; }
  using QuaternionBase<Quaternion<S>, S>::x;
  using QuaternionBase<Quaternion<S>, S>::y;
  using QuaternionBase<Quaternion<S>, S>::z;
  using QuaternionBase<Quaternion<S>, S>::w;

 private:
  S q_[4];
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <class S>
class Map<Quaternion<S>> : public QuaternionBase<Map<Quaternion<S>>, S> {
 public:
  explicit Map(S* p) : p_(p) {}
  Map(const Map& o) : p_(o.p_) {}
  const S* coeff_ptr() const { return p_; }
  Map& operator=(const Map& o) { S t[4]; std::memcpy(t, o.p_, sizeof(t)); std::memcpy(p_, t, sizeof(t)); return *this; }
  template <class O> Map& operator=(const QuaternionBase<O, S>& o) { S t[4]; std::memcpy(t, o.c(), sizeof(t)); std::memcpy(p_, t, sizeof(t)); return *this; }

 private:
  S* p_;
};
template <class S>
class Map<const Quaternion<S>> : public QuaternionBase<Map<const Quaternion<S>>, S> {
 public:
  explicit Map(const S* p) : p_(p) {}
  const S* coeff_ptr() const { return p_; }

 private:
  const S* p_;
};

template <class S>
class AngleAxis {
 public:
  AngleAxis() : angle_(0) {}
  template <class V> AngleAxis(S angle, const DenseBase<V, S, 3, 1>& axis) : angle_(angle), axis_(axis) {}
  template <class O> explicit AngleAxis(const QuaternionBase<O, S>& q) { from_quat(Quaternion<S>(q)); }
  template <class M> explicit AngleAxis(const DenseBase<M, S, 3, 3>& m) { from_quat(Quaternion<S>(m)); }  // fromRotationMatrix: via quaternion
  S angle() const { return angle_; }
  const Matrix<S, 3, 1>& axis() const { return axis_; }
  Quaternion<S> operator*(const AngleAxis& o) const { return Quaternion<S>(*this) * Quaternion<S>(o); }
  template <class O> Quaternion<S> operator*(const QuaternionBase<O, S>& o) const { return Quaternion<S>(*this) * o; }
  Matrix<S, 3, 3> toRotationMatrix() const { return Quaternion<S>(*this).toRotationMatrix(); }

 private:
  void from_quat(const Quaternion<S>& q) {  // AngleAxis::operator=(QuaternionBase), Eigen 3.3
    S n = q.vec().norm();
    if (n < std::numeric_limits<S>::epsilon()) n = q.vec().norm();  // stableNorm in Eigen; same value for finite inputs
    if (n != S(0)) {
      angle_ = S(2) * std::atan2(n, std::abs(q.w()));
      if (q.w() < S(0)) n = -n;
      axis_ = q.vec() / n;
    } else {
      angle_ = S(0);
      axis_ = Matrix<S, 3, 1>(S(1), S(0), S(0));
    }
  }
  S angle_;
  Matrix<S, 3, 1> axis_;
};
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;

// ---------------------------------------------------------------------------------------------------------------------
// Transform<S,3,Mode>: 4x4 column-major storage like Eigen (Isometry and Affine modes)
template <class S, int Dim, int Mode>
class Transform {
  static_assert(Dim == 3, "3-D transforms only");

 public:
  typedef Matrix<S, 4, 4> MatrixType;
  Transform() {}
  template <int OtherMode> Transform(const Transform<S, 3, OtherMode>& o) : m_(o.matrix()) {}
  template <class O> explicit Transform(const QuaternionBase<O, S>& q) { m_.setIdentity(); linear() = q.toRotationMatrix(); }
  template <class M> explicit Transform(const DenseBase<M, S, 4, 4>& m) : m_(m) {}
  static Transform Identity() { Transform t; t.m_.setIdentity(); return t; }
  void setIdentity() { m_.setIdentity(); }
  const MatrixType& matrix() const { return m_; }
  MatrixType& matrix() { return m_; }
  S operator()(int i, int j) const { return m_.at(i, j); }
  S& operator()(int i, int j) { return m_.at(i, j); }
  View<S, 3, 3> linear() { return View<S, 3, 3>(m_.data(), 1, 4); }
  Matrix<S, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
  Matrix<S, 3, 3> rotation() const { return linear(); }  // Eigen 3.3 transform_rotation_impl<Isometry>: linear() as is
  View<S, 3, 1> translation() { return View<S, 3, 1>(m_.data() + 12, 1, 4); }
  Matrix<S, 3, 1> translation() const { return m_.template block<3, 1>(0, 3); }
  // transform_transform_product_impl (non-projective): linear = L1 L2 ; translation = L1 t2 + t1
  template <int OtherMode>
  Transform operator*(const Transform<S, 3, OtherMode>& o) const {
    Transform r;
    r.m_.setIdentity();
    const Matrix<S, 3, 3> L = linear();
    r.linear() = L * o.linear();
    r.translation() = L * o.translation() + translation();
    return r;
  }
  template <class V>
  Matrix<S, 3, 1> operator*(const DenseBase<V, S, 3, 1>& v) const { return linear() * v + translation(); }
  // Transform::inverse(Isometry): R^T, -R^T t.  (Affine mode would invert the linear part; the path never does that.)
  Transform inverse() const {
    Transform r;
    r.m_.setIdentity();
    if (Mode == Isometry) {
      const Matrix<S, 3, 3> Rt = linear().transpose();
      r.linear() = Rt;
      r.translation() = -(Rt * translation());
    } else {   // Affine: general 3x3 inverse by cofactors (Eigen's compute_inverse_size3_helper), then -L^-1 t  (src/utils.cpp:35, off the hot path)
      const Matrix<S, 3, 3> L = linear();
      auto cof = [&](int i, int j) {
        const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
        return L.at(i1, j1) * L.at(i2, j2) - L.at(i1, j2) * L.at(i2, j1);
      };
      const S c0 = cof(0, 0), c1 = cof(1, 0), c2 = cof(2, 0);
      const S det = c0 * L.at(0, 0) + c1 * L.at(1, 0) + c2 * L.at(2, 0);
      const S invdet = S(1) / det;
      Matrix<S, 3, 3> Li;
      Li.at(0, 0) = c0 * invdet; Li.at(0, 1) = c1 * invdet; Li.at(0, 2) = c2 * invdet;
      Li.at(1, 0) = cof(0, 1) * invdet; Li.at(1, 1) = cof(1, 1) * invdet; Li.at(1, 2) = cof(2, 1) * invdet;
      Li.at(2, 0) = cof(0, 2) * invdet; Li.at(2, 1) = cof(1, 2) * invdet; Li.at(2, 2) = cof(2, 2) * invdet;
      r.linear() = Li;
      r.translation() = -(Li * translation());
    }
    return r;
  }
  template <class T>
  Transform<T, 3, Mode> cast() const { Transform<T, 3, Mode> r; r.matrix() = m_.template cast<T>(); return r; }
  template <class O> Transform& rotate(const QuaternionBase<O, S>& q) { const Matrix<S, 3, 3> L = linear(); linear() = L * q.toRotationMatrix(); return *this; }
  template <class V> Transform& pretranslate(const DenseBase<V, S, 3, 1>& v) { translation() = translation() + v; return *this; }
  template <class V> Transform& translate(const DenseBase<V, S, 3, 1>& v) { const Matrix<S, 3, 3> L = linear(); translation() = translation() + L * v; return *this; }

 private:
  MatrixType m_;
};
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<double, 3, Affine> Affine3d;
typedef Transform<float, 3, Affine> Affine3f;

// ---------------------------------------------------------------------------------------------------------------------
// SelfAdjointEigenSolver<Matrix3d> (iterative QL path, src/odomEstimationClass.cpp:175)
template <class MatrixType>
class SelfAdjointEigenSolver {
 public:
  template <class M>
  explicit SelfAdjointEigenSolver(const DenseBase<M, double, 3, 3>& a) {
    const fo::Eigen3 e = fo::self_adjoint_eigen3(internal::mat3_to_fo(a.derived()));
    for (int i = 0; i < 3; ++i) { values_(i) = e.values[i]; for (int j = 0; j < 3; ++j) vectors_.at(i, j) = e.vectors[i][j]; }
  }
  const Matrix<double, 3, 3>& eigenvectors() const { return vectors_; }
  const Matrix<double, 3, 1>& eigenvalues() const { return values_; }

 private:
  Matrix<double, 3, 3> vectors_;
  Matrix<double, 3, 1> values_;
};

}  // namespace Eigen
