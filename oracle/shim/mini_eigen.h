// ORACLE / TEST INFRASTRUCTURE -- not product code, never linked into libvicon2gt_b200.so.
//
// A small stand-in for the part of Eigen's INTERFACE that the reference's hot-path sources use
// (src/utils/quat_ops.h, src/utils/rpy_ops.h, src/cpi/CpiBase.h, src/cpi/CpiV1.h, src/meas/*.cpp,
// src/gtsam/*.cpp).  Eigen itself is not in this image, so oracle/Makefile.ref compiles those
// reference files WHERE THEY LIE under /root/reference against this header to obtain
// oracle/_ref/libv2gt_ref.so -- the reference's own arithmetic, statement by statement, with only
// the matrix container swapped.  Everything here is eager, row-major, run-time sized and slow on
// purpose; products and sums are evaluated in the textbook order (sum over k ascending), which is
// also what Eigen's small fixed-size products do, so results agree with a real Eigen build to
// round-off.  Where Eigen has a specific small-size algorithm (3x3 inverse by cofactors, unblocked
// LLT with its early exit on a non-positive pivot) the same algorithm is used.
//
// Fixed-size objects are NaN-filled on default construction (Eigen leaves them uninitialised), so a
// reliance on uninitialised memory shows up as NaN instead of passing by luck.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <memory>
#include <map> // the real Eigen headers pull these in; the reference relies on that (src/sim/BsplineSE3.h uses std::map)
#include <string>
#include <type_traits>
#include <vector>

namespace Eigen {

const int Dynamic = -1;
template <class T, int R, int C> class Matrix;
typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, Dynamic, 1> VectorXd;
template <class M> class BlockRef;
class LLTResult;
class QRResult;

template <class T> using aligned_allocator = std::allocator<T>;

inline void mini_fail(const char *what) {
  std::fprintf(stderr, "mini_eigen: %s\n", what);
  std::abort();
}

// ---- common read interface (CRTP)
template <class D> class Base {
public:
  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }
  int rows() const { return derived().rows_(); }
  int cols() const { return derived().cols_(); }
  int size() const { return rows() * cols(); }
  double coeff(int i, int j) const { return derived().get_(i, j); }
  double operator()(int i, int j) const { return coeff(i, j); }
  double operator()(int i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }

  MatrixXd eval() const;
  MatrixXd transpose() const;
  MatrixXd inverse() const;
  MatrixXd block(int i, int j, int r, int c) const;
  MatrixXd head(int n) const;
  MatrixXd tail(int n) const;
  double norm() const { return std::sqrt(squaredNorm()); }
  double squaredNorm() const {
    double s = 0;
    for (int i = 0; i < rows(); i++)
      for (int j = 0; j < cols(); j++)
        s += coeff(i, j) * coeff(i, j);
    return s;
  }
  double trace() const {
    double s = 0;
    for (int i = 0; i < rows() && i < cols(); i++)
      s += coeff(i, i);
    return s;
  }
  template <class E> double dot(const Base<E> &o) const {
    if (size() != o.size())
      mini_fail("dot: size mismatch");
    double s = 0;
    for (int i = 0; i < size(); i++)
      s += (*this)(i)*o(i);
    return s;
  }
  LLTResult llt() const;
  QRResult colPivHouseholderQr() const;
};

// ---- comma initialiser:  m << a, b, c;   scalars fill row-major; a vector item appends its entries
template <class Target> class CommaInit {
  Target &t_;
  int k_ = 0;
  void put(double v) {
    int n = t_.rows() * t_.cols();
    if (k_ >= n)
      mini_fail("comma initialiser: too many coefficients");
    t_.set_(k_ / t_.cols(), k_ % t_.cols(), v);
    k_++;
  }

public:
  CommaInit(Target &t) : t_(t) {}
  CommaInit &add(double v) {
    put(v);
    return *this;
  }
  template <class E> CommaInit &add(const Base<E> &m) {
    if (t_.cols() != 1 || m.cols() != 1)
      mini_fail("comma initialiser: only column vectors can be stacked");
    for (int i = 0; i < m.rows(); i++)
      put(m(i, 0));
    return *this;
  }
  CommaInit &operator,(double v) { return add(v); }
  template <class E> CommaInit &operator,(const Base<E> &m) { return add(m); }
  Target &finished() { return t_; }
};

// ---- owning matrix
template <class T, int R, int C> class Matrix : public Base<Matrix<T, R, C>> {
  static_assert(std::is_same<T, double>::value, "mini_eigen: double only");
  int r_, c_;
  std::vector<double> a_;
  void alloc(int r, int c, double fill) {
    if ((R != Dynamic && r != R) || (C != Dynamic && c != C))
      mini_fail("size mismatch on a fixed-size matrix");
    r_ = r;
    c_ = c;
    a_.assign((size_t)r * c, fill);
  }

public:
  int rows_() const { return r_; }
  int cols_() const { return c_; }
  double get_(int i, int j) const {
    if (i < 0 || j < 0 || i >= r_ || j >= c_)
      mini_fail("index out of range");
    return a_[(size_t)i * c_ + j];
  }
  void set_(int i, int j, double v) { ref_(i, j) = v; }
  double &ref_(int i, int j) {
    if (i < 0 || j < 0 || i >= r_ || j >= c_)
      mini_fail("index out of range");
    return a_[(size_t)i * c_ + j];
  }

  Matrix() { alloc(R == Dynamic ? 0 : R, C == Dynamic ? 0 : C, std::numeric_limits<double>::quiet_NaN()); }
  Matrix(const Matrix &) = default;
  Matrix &operator=(const Matrix &) = default;
  template <class E> Matrix(const Base<E> &o) {
    alloc(o.rows(), o.cols(), 0.0);
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++)
        a_[(size_t)i * c_ + j] = o.coeff(i, j);
  }
  // (x, y) of a 2-vector, or (rows, cols) of a run-time sized matrix
  Matrix(double x, double y) {
    if (R != Dynamic && C != Dynamic && R * C == 2) {
      alloc(R, C, 0.0);
      a_[0] = x;
      a_[1] = y;
    } else {
      alloc((int)x, (int)y, std::numeric_limits<double>::quiet_NaN());
    }
  }
  Matrix(double x, double y, double z) {
    static_assert(R == Dynamic || C == Dynamic || R * C == 3, "3-coefficient constructor on a non 3-vector");
    alloc(R == Dynamic ? 3 : R, C == Dynamic ? 1 : C, 0.0);
    a_[0] = x, a_[1] = y, a_[2] = z;
  }
  Matrix(double x, double y, double z, double w) {
    static_assert(R == Dynamic || C == Dynamic || R * C == 4, "4-coefficient constructor on a non 4-vector");
    alloc(R == Dynamic ? 4 : R, C == Dynamic ? 1 : C, 0.0);
    a_[0] = x, a_[1] = y, a_[2] = z, a_[3] = w;
  }
  template <class E> Matrix &operator=(const Base<E> &o) {
    Matrix tmp(o); // evaluate first: the right-hand side may alias *this
    if (R != Dynamic && C != Dynamic && (tmp.r_ != R || tmp.c_ != C))
      mini_fail("assignment: size mismatch");
    r_ = tmp.r_, c_ = tmp.c_;
    a_.swap(tmp.a_);
    return *this;
  }

  using Base<Matrix>::operator();
  using Base<Matrix>::block;
  using Base<Matrix>::head;
  using Base<Matrix>::tail;
  double &operator()(int i, int j) { return ref_(i, j); }
  double &operator()(int i) { return c_ == 1 ? ref_(i, 0) : ref_(0, i); }
  BlockRef<Matrix> block(int i, int j, int r, int c);
  BlockRef<Matrix> head(int n);
  BlockRef<Matrix> tail(int n);

  static Matrix filled(int r, int c, double v) {
    Matrix m;
    m.alloc(r, c, v);
    return m;
  }
  static Matrix Zero() { return filled(R, C, 0.0); }
  static Matrix Zero(int r, int c) { return filled(r, c, 0.0); }
  static Matrix Zero(int n) { return filled(R == 1 ? 1 : n, R == 1 ? n : 1, 0.0); }
  static Matrix Constant(double v) { return filled(R, C, v); }
  static Matrix Identity() { return Identity(R, C); }
  static Matrix Identity(int r, int c) {
    Matrix m = filled(r, c, 0.0);
    for (int i = 0; i < r && i < c; i++)
      m.ref_(i, i) = 1.0;
    return m;
  }
  Matrix &setZero() {
    a_.assign(a_.size(), 0.0);
    return *this;
  }
  const Matrix &matrix() const { return *this; }
  Matrix &setIdentity() {
    *this = Identity(r_, c_);
    return *this;
  }
  CommaInit<Matrix> operator<<(double v) {
    CommaInit<Matrix> ci(*this);
    ci.add(v);
    return ci;
  }
  template <class E> CommaInit<Matrix> operator<<(const Base<E> &m) {
    CommaInit<Matrix> ci(*this);
    ci.add(m);
    return ci;
  }
  template <class E> Matrix &operator+=(const Base<E> &o) { return *this = *this + o; }
  template <class E> Matrix &operator-=(const Base<E> &o) { return *this = *this - o; }
  Matrix &operator*=(double s) { return *this = *this * s; }
  Matrix &operator/=(double s) { return *this = *this / s; }
};

typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<double, 4, 4> Matrix4d;

// ---- writable view of a rectangular part of a matrix
template <class M> class BlockRef : public Base<BlockRef<M>> {
  M &m_;
  int i0_, j0_, r_, c_;

public:
  BlockRef(M &m, int i0, int j0, int r, int c) : m_(m), i0_(i0), j0_(j0), r_(r), c_(c) {
    if (i0 < 0 || j0 < 0 || r < 0 || c < 0 || i0 + r > m.rows() || j0 + c > m.cols())
      mini_fail("block out of range");
  }
  int rows_() const { return r_; }
  int cols_() const { return c_; }
  double get_(int i, int j) const {
    if (i < 0 || j < 0 || i >= r_ || j >= c_)
      mini_fail("block index out of range");
    return static_cast<const M &>(m_).get_(i0_ + i, j0_ + j);
  }
  void set_(int i, int j, double v) {
    if (i < 0 || j < 0 || i >= r_ || j >= c_)
      mini_fail("block index out of range");
    m_.set_(i0_ + i, j0_ + j, v);
  }
  template <class E> BlockRef &assign(const Base<E> &o) {
    if (o.rows() != r_ || o.cols() != c_)
      mini_fail("block assignment: size mismatch");
    MatrixXd tmp(o); // the right-hand side may alias the parent
    for (int i = 0; i < r_; i++)
      for (int j = 0; j < c_; j++)
        set_(i, j, tmp.coeff(i, j));
    return *this;
  }
  template <class E> BlockRef &operator=(const Base<E> &o) { return assign(o); }
  BlockRef &operator=(const BlockRef &o) { return assign(o); }
  template <class E> BlockRef &operator+=(const Base<E> &o) { return assign(*this + o); }
  template <class E> BlockRef &operator-=(const Base<E> &o) { return assign(*this - o); }
  CommaInit<BlockRef> operator<<(double v) {
    CommaInit<BlockRef> ci(*this);
    ci.add(v);
    return ci;
  }
  template <class E> CommaInit<BlockRef> operator<<(const Base<E> &m) {
    CommaInit<BlockRef> ci(*this);
    ci.add(m);
    return ci;
  }
};

template <class T, int R, int C> BlockRef<Matrix<T, R, C>> Matrix<T, R, C>::block(int i, int j, int r, int c) {
  return BlockRef<Matrix>(*this, i, j, r, c);
}
template <class T, int R, int C> BlockRef<Matrix<T, R, C>> Matrix<T, R, C>::head(int n) {
  return c_ == 1 ? BlockRef<Matrix>(*this, 0, 0, n, 1) : BlockRef<Matrix>(*this, 0, 0, 1, n);
}
template <class T, int R, int C> BlockRef<Matrix<T, R, C>> Matrix<T, R, C>::tail(int n) {
  return c_ == 1 ? BlockRef<Matrix>(*this, r_ - n, 0, n, 1) : BlockRef<Matrix>(*this, 0, c_ - n, 1, n);
}

// ---- arithmetic (always evaluated into a run-time sized matrix)
template <class A, class B> MatrixXd operator+(const Base<A> &a, const Base<B> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols())
    mini_fail("operator+: size mismatch");
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return o;
}
template <class A, class B> MatrixXd operator-(const Base<A> &a, const Base<B> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols())
    mini_fail("operator-: size mismatch");
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return o;
}
template <class A> MatrixXd operator-(const Base<A> &a) {
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = -a.coeff(i, j);
  return o;
}
template <class A, class B> MatrixXd operator*(const Base<A> &a, const Base<B> &b) {
  if (a.cols() != b.rows())
    mini_fail("operator*: inner size mismatch");
  MatrixXd o = MatrixXd::Zero(a.rows(), b.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < b.cols(); j++) {
      double s = a.coeff(i, 0) * b.coeff(0, j);
      for (int k = 1; k < a.cols(); k++)
        s += a.coeff(i, k) * b.coeff(k, j);
      o(i, j) = s;
    }
  return o;
}
template <class A> MatrixXd operator*(const Base<A> &a, double s) {
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = a.coeff(i, j) * s;
  return o;
}
template <class A> MatrixXd operator*(double s, const Base<A> &a) {
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = s * a.coeff(i, j);
  return o;
}
template <class A> MatrixXd operator/(const Base<A> &a, double s) {
  MatrixXd o = MatrixXd::Zero(a.rows(), a.cols());
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      o(i, j) = a.coeff(i, j) / s;
  return o;
}
template <class A, class B> bool operator==(const Base<A> &a, const Base<B> &b) {
  if (a.rows() != b.rows() || a.cols() != b.cols())
    mini_fail("operator==: size mismatch");
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      if (!(a.coeff(i, j) == b.coeff(i, j)))
        return false;
  return true;
}
template <class A, class B> bool operator!=(const Base<A> &a, const Base<B> &b) { return !(a == b); }
template <class A> std::ostream &operator<<(std::ostream &os, const Base<A> &a) {
  for (int i = 0; i < a.rows(); i++) {
    for (int j = 0; j < a.cols(); j++)
      os << (j ? " " : "") << a.coeff(i, j);
    if (i + 1 < a.rows())
      os << "\n";
  }
  return os;
}

// ---- Base members that return matrices
template <class D> MatrixXd Base<D>::eval() const { return MatrixXd(*this); }
template <class D> MatrixXd Base<D>::transpose() const {
  MatrixXd o = MatrixXd::Zero(cols(), rows());
  for (int i = 0; i < rows(); i++)
    for (int j = 0; j < cols(); j++)
      o(j, i) = coeff(i, j);
  return o;
}
template <class D> MatrixXd Base<D>::block(int i0, int j0, int r, int c) const {
  if (i0 < 0 || j0 < 0 || r < 0 || c < 0 || i0 + r > rows() || j0 + c > cols())
    mini_fail("block out of range");
  MatrixXd o = MatrixXd::Zero(r, c);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++)
      o(i, j) = coeff(i0 + i, j0 + j);
  return o;
}
template <class D> MatrixXd Base<D>::head(int n) const { return cols() == 1 ? block(0, 0, n, 1) : block(0, 0, 1, n); }
template <class D> MatrixXd Base<D>::tail(int n) const {
  return cols() == 1 ? block(rows() - n, 0, n, 1) : block(0, cols() - n, 1, n);
}

// general solve A X = B by Gaussian elimination with partial pivoting (used for n != 3 inverses and
// for the QR stand-in); an exactly zero A gives X = 0, which is what Eigen's rank-revealing
// ColPivHouseholderQR::solve returns for a rank-0 matrix
inline MatrixXd mini_solve(MatrixXd A, MatrixXd B, bool zero_if_singular) {
  int n = A.rows();
  if (A.cols() != n || B.rows() != n)
    mini_fail("solve: shape");
  double amax = 0;
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++)
      amax = std::fmax(amax, std::fabs(A(i, j)));
  if (amax == 0 && zero_if_singular)
    return MatrixXd::Zero(n, B.cols());
  for (int k = 0; k < n; k++) {
    int p = k;
    for (int i = k + 1; i < n; i++)
      if (std::fabs(A(i, k)) > std::fabs(A(p, k)))
        p = i;
    if (p != k) {
      for (int j = 0; j < n; j++)
        std::swap(A(k, j), A(p, j));
      for (int j = 0; j < B.cols(); j++)
        std::swap(B(k, j), B(p, j));
    }
    for (int i = k + 1; i < n; i++) {
      double f = A(i, k) / A(k, k);
      if (f == 0)
        continue;
      for (int j = k; j < n; j++)
        A(i, j) -= f * A(k, j);
      for (int j = 0; j < B.cols(); j++)
        B(i, j) -= f * B(k, j);
    }
  }
  MatrixXd X = MatrixXd::Zero(n, B.cols());
  for (int j = 0; j < B.cols(); j++)
    for (int i = n - 1; i >= 0; i--) {
      double s = B(i, j);
      for (int k = i + 1; k < n; k++)
        s -= A(i, k) * X(k, j);
      X(i, j) = s / A(i, i);
    }
  return X;
}

template <class D> MatrixXd Base<D>::inverse() const {
  if (rows() != cols())
    mini_fail("inverse of a non-square matrix");
  if (rows() == 3) { // Eigen's compute_inverse<.., 3>: cofactors, one division by the determinant
    const Base<D> &m = *this;
    auto cof = [&](int i, int j) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      return m(i1, j1) * m(i2, j2) - m(i1, j2) * m(i2, j1);
    };
    double c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
    double det = c00 * m(0, 0) + c10 * m(1, 0) + c20 * m(2, 0);
    double invdet = 1.0 / det;
    MatrixXd o = MatrixXd::Zero(3, 3);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        o(j, i) = cof(i, j) * invdet;
    return o;
  }
  return mini_solve(MatrixXd(*this), MatrixXd::Identity(rows(), rows()), false);
}

// LLT as Eigen's unblocked kernel (llt_inplace<Lower>::unblocked): column by column, stop at the
// first non-positive pivot leaving the rest of the matrix as it was; matrixL() is the lower triangle
class LLTResult {
  MatrixXd m_;

public:
  explicit LLTResult(const MatrixXd &a) : m_(a) {
    int n = m_.rows();
    for (int k = 0; k < n; k++) {
      double x = m_(k, k);
      for (int j = 0; j < k; j++)
        x -= m_(k, j) * m_(k, j);
      if (x <= 0)
        break;
      x = std::sqrt(x);
      m_(k, k) = x;
      for (int i = k + 1; i < n; i++) {
        double s = m_(i, k);
        for (int j = 0; j < k; j++)
          s -= m_(i, j) * m_(k, j);
        m_(i, k) = s / x;
      }
    }
  }
  MatrixXd matrixL() const {
    MatrixXd o = m_;
    for (int i = 0; i < o.rows(); i++)
      for (int j = i + 1; j < o.cols(); j++)
        o(i, j) = 0.0;
    return o;
  }
};
class QRResult {
  MatrixXd m_;

public:
  explicit QRResult(const MatrixXd &a) : m_(a) {}
  template <class E> MatrixXd solve(const Base<E> &b) const { return mini_solve(m_, MatrixXd(b), true); }
};
template <class D> LLTResult Base<D>::llt() const { return LLTResult(MatrixXd(*this)); }
template <class D> QRResult Base<D>::colPivHouseholderQr() const { return QRResult(MatrixXd(*this)); }

} // namespace Eigen
