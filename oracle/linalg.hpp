// ORACLE (test infrastructure, NOT product code) -- container only (no reference counterpart).
//
// Tiny fixed-size dense matrix type used by the CPU restatement of the
// vicon2gt build-and-solve path.  The reference uses Eigen (absent from this
// image); this header gives the oracle the few operations it needs so that the
// restatement can follow the reference's formulas one to one.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
// --impl reference legs may use anything in oracle/.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstring>
#include <initializer_list>

namespace orc {

template <int R, int C> struct Mat {
  double a[R * C];
  Mat() { std::memset(a, 0, sizeof(a)); }
  Mat(std::initializer_list<double> v) {
    std::memset(a, 0, sizeof(a));
    int i = 0;
    for (double x : v) {
      if (i < R * C)
        a[i++] = x;
    }
  }
  static Mat Zero() { return Mat(); }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); i++)
      m(i, i) = 1.0;
    return m;
  }
  double &operator()(int r, int c) { return a[r * C + c]; }
  double operator()(int r, int c) const { return a[r * C + c]; }
  // vector style access (column or row vectors)
  double &operator()(int i) { return a[i]; }
  double operator()(int i) const { return a[i]; }
  double &operator[](int i) { return a[i]; }
  double operator[](int i) const { return a[i]; }

  Mat<C, R> transpose() const {
    Mat<C, R> t;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C; c++)
        t(c, r) = (*this)(r, c);
    return t;
  }
  double squaredNorm() const {
    double s = 0;
    for (int i = 0; i < R * C; i++)
      s += a[i] * a[i];
    return s;
  }
  double norm() const { return std::sqrt(squaredNorm()); }
  double trace() const {
    double s = 0;
    for (int i = 0; i < (R < C ? R : C); i++)
      s += (*this)(i, i);
    return s;
  }
  template <int BR, int BC> Mat<BR, BC> block(int r0, int c0) const {
    Mat<BR, BC> b;
    for (int r = 0; r < BR; r++)
      for (int c = 0; c < BC; c++)
        b(r, c) = (*this)(r0 + r, c0 + c);
    return b;
  }
  template <int BR, int BC> void setBlock(int r0, int c0, const Mat<BR, BC> &b) {
    for (int r = 0; r < BR; r++)
      for (int c = 0; c < BC; c++)
        (*this)(r0 + r, c0 + c) = b(r, c);
  }
  bool hasNaN() const {
    for (int i = 0; i < R * C; i++)
      if (std::isnan(a[i]))
        return true;
    return false;
  }
  bool operator==(const Mat &o) const {
    for (int i = 0; i < R * C; i++)
      if (a[i] != o.a[i])
        return false;
    return true;
  }
  bool operator!=(const Mat &o) const { return !(*this == o); }
  Mat &operator+=(const Mat &o) {
    for (int i = 0; i < R * C; i++)
      a[i] += o.a[i];
    return *this;
  }
  Mat &operator-=(const Mat &o) {
    for (int i = 0; i < R * C; i++)
      a[i] -= o.a[i];
    return *this;
  }
  Mat &operator*=(double s) {
    for (int i = 0; i < R * C; i++)
      a[i] *= s;
    return *this;
  }
};

template <int R, int C> Mat<R, C> operator+(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; i++)
    z.a[i] = x.a[i] + y.a[i];
  return z;
}
template <int R, int C> Mat<R, C> operator-(const Mat<R, C> &x, const Mat<R, C> &y) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; i++)
    z.a[i] = x.a[i] - y.a[i];
  return z;
}
template <int R, int C> Mat<R, C> operator-(const Mat<R, C> &x) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; i++)
    z.a[i] = -x.a[i];
  return z;
}
template <int R, int C> Mat<R, C> operator*(double s, const Mat<R, C> &x) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; i++)
    z.a[i] = s * x.a[i];
  return z;
}
template <int R, int C> Mat<R, C> operator*(const Mat<R, C> &x, double s) { return s * x; }
template <int R, int C> Mat<R, C> operator/(const Mat<R, C> &x, double s) {
  Mat<R, C> z;
  for (int i = 0; i < R * C; i++)
    z.a[i] = x.a[i] / s;
  return z;
}
template <int R, int K, int C> Mat<R, C> operator*(const Mat<R, K> &x, const Mat<K, C> &y) {
  Mat<R, C> z;
  for (int r = 0; r < R; r++)
    for (int k = 0; k < K; k++) {
      const double xv = x(r, k);
      if (xv == 0.0)
        continue; // exact: adding 0*y changes nothing for finite y (sparse F, G, H blocks)
      for (int c = 0; c < C; c++)
        z(r, c) += xv * y(k, c);
    }
  return z;
}
template <int N> double dot(const Mat<N, 1> &x, const Mat<N, 1> &y) {
  double s = 0;
  for (int i = 0; i < N; i++)
    s += x.a[i] * y.a[i];
  return s;
}

using Vec3 = Mat<3, 1>;
using Vec4 = Mat<4, 1>;
using Vec6 = Mat<6, 1>;
using Vec15 = Mat<15, 1>;
using Mat3 = Mat<3, 3>;
using Mat4 = Mat<4, 4>;
using Mat6 = Mat<6, 6>;
using Mat15 = Mat<15, 15>;

// Lower Cholesky A = L L^T.  Mirrors Eigen::LLT semantics closely enough for the
// oracle: on a non-positive pivot the factorisation stops and reports failure,
// leaving the already computed columns in place and the rest zero.
template <int N> bool chol_lower(const Mat<N, N> &A, Mat<N, N> &L) {
  L = Mat<N, N>();
  for (int j = 0; j < N; j++) {
    double d = A(j, j);
    for (int k = 0; k < j; k++)
      d -= L(j, k) * L(j, k);
    if (!(d > 0.0))
      return false;
    const double ljj = std::sqrt(d);
    L(j, j) = ljj;
    for (int i = j + 1; i < N; i++) {
      double s = A(i, j);
      for (int k = 0; k < j; k++)
        s -= L(i, k) * L(j, k);
      L(i, j) = s / ljj;
    }
  }
  return true;
}

// Inverse of a lower-triangular matrix (forward substitution on the identity).
// A zero diagonal yields a zero row/column block (rank-revealing behaviour of the
// colPivHouseholderQr solve the reference uses on a zero matrix gives W = 0).
template <int N> Mat<N, N> inv_lower(const Mat<N, N> &L) {
  Mat<N, N> X;
  for (int c = 0; c < N; c++) {
    for (int r = c; r < N; r++) {
      double s = (r == c) ? 1.0 : 0.0;
      for (int k = c; k < r; k++)
        s -= L(r, k) * X(k, c);
      X(r, c) = (L(r, r) != 0.0) ? s / L(r, r) : 0.0;
    }
  }
  return X;
}

// General inverse by Gauss-Jordan with partial pivoting (stand-in for
// Eigen's PartialPivLU based .inverse()).  Singular input produces inf/NaN
// entries just like the reference's unchecked inverse.
template <int N> Mat<N, N> inverse(const Mat<N, N> &A) {
  Mat<N, N> M = A, I = Mat<N, N>::Identity();
  for (int c = 0; c < N; c++) {
    int piv = c;
    double best = std::fabs(M(c, c));
    for (int r = c + 1; r < N; r++)
      if (std::fabs(M(r, c)) > best) {
        best = std::fabs(M(r, c));
        piv = r;
      }
    if (piv != c)
      for (int k = 0; k < N; k++) {
        double t = M(c, k);
        M(c, k) = M(piv, k);
        M(piv, k) = t;
        t = I(c, k);
        I(c, k) = I(piv, k);
        I(piv, k) = t;
      }
    const double d = 1.0 / M(c, c);
    for (int k = 0; k < N; k++) {
      M(c, k) *= d;
      I(c, k) *= d;
    }
    for (int r = 0; r < N; r++) {
      if (r == c)
        continue;
      const double f = M(r, c);
      if (f == 0.0)
        continue;
      for (int k = 0; k < N; k++) {
        M(r, k) -= f * M(c, k);
        I(r, k) -= f * I(c, k);
      }
    }
  }
  return I;
}

// SPD inverse through Cholesky: A^-1 = L^-T L^-1.
template <int N> bool inverse_spd(const Mat<N, N> &A, Mat<N, N> &Ainv) {
  Mat<N, N> L;
  if (!chol_lower(A, L))
    return false;
  Mat<N, N> Li = inv_lower(L);
  Ainv = Li.transpose() * Li;
  return true;
}

} // namespace orc
