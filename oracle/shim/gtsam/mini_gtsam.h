// ORACLE / TEST INFRASTRUCTURE -- not product code.
//
// Type scaffolding that lets the reference's factor and node sources (src/gtsam/*.h, *.cpp) compile
// without GTSAM 4.2a5 / Boost: the typedefs, Values (a typed key -> node map), the NoiseModelFactorN
// base classes (keys + a noise-model handle + the virtual evaluateError the reference overrides),
// boost::optional<Matrix&>, OptionalJacobian<N,M>, traits / Manifold and gtsam::equal.
// NO GTSAM ARITHMETIC lives here: the noise models only remember how they were constructed (the
// whitening by Gaussian::Covariance, the Huber weights and the Levenberg-Marquardt optimiser are
// GTSAM's; oracle/solver.hpp restates them, and mini_gtsam_lm.h restates them a second time only to
// drive the reference's ViconGraphSolver.cpp end to end -- both unpinned).  What this header buys is
// the reference's own evaluateError / retract / localCoordinates bodies, compiled where they lie.
#pragma once
#include <cmath>
#include <cstdint>
#include <functional>
#include <iostream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../mini_eigen.h"

#define GTSAM_EXPORT

namespace boost {
struct none_t {};
static const none_t none = none_t();
template <class T> class optional; // only the reference-to-matrix form is used by the reference
template <class T> class optional<T &> {
  T *p_;

public:
  optional() : p_(nullptr) {}
  optional(none_t) : p_(nullptr) {}
  optional(T &t) : p_(&t) {}
  explicit operator bool() const { return p_ != nullptr; }
  T &operator*() const { return *p_; }
  T *operator->() const { return p_; }
};
} // namespace boost

namespace gtsam {

typedef Eigen::MatrixXd Matrix;
typedef Eigen::VectorXd Vector;
typedef Eigen::Matrix<double, 1, 1> Vector1;
typedef Eigen::Matrix<double, 2, 1> Vector2;
typedef Eigen::Matrix<double, 3, 1> Vector3;
typedef Eigen::Matrix<double, 4, 1> Vector4;
typedef Eigen::Matrix<double, 6, 1> Vector6;
typedef std::uint64_t Key;
typedef std::function<std::string(Key)> KeyFormatter;
inline std::string mini_default_key_formatter(Key k) { return std::to_string(k); }
static const KeyFormatter DefaultKeyFormatter = &mini_default_key_formatter;

inline bool equal(double a, double b, double tol) { return std::fabs(a - b) <= tol; }
template <class A, class B> bool equal(const Eigen::Base<A> &a, const Eigen::Base<B> &b, double tol) {
  if (a.rows() != b.rows() || a.cols() != b.cols())
    return false;
  for (int i = 0; i < a.rows(); i++)
    for (int j = 0; j < a.cols(); j++)
      if (!(std::fabs(a.coeff(i, j) - b.coeff(i, j)) <= tol))
        return false;
  return true;
}

// traits<T>: dimension and retract of a node type.  The reference declares
// `template <> struct traits<JPLNavState> : internal::Manifold<JPLNavState> {}` for its three node classes
// (T::dimension, T::retract); fixed-size vectors are vector spaces.
template <class T> struct traits {};
namespace internal {
template <class T> struct Manifold {
  static int dim(const T &) { return T::dimension; }
  static T retract(const T &v, const Vector &d) { return v.retract(Eigen::Matrix<double, T::dimension, 1>(d)); }
};
} // namespace internal
template <int N> struct traits<Eigen::Matrix<double, N, 1>> {
  static int dim(const Eigen::Matrix<double, N, 1> &) { return N; }
  static Eigen::Matrix<double, N, 1> retract(const Eigen::Matrix<double, N, 1> &v, const Vector &d) {
    return Eigen::Matrix<double, N, 1>(v + d);
  }
};

// Values: key -> node of any type (insert / at<T> / find / erase / clear as ViconGraphSolver.cpp uses them)
struct ValueBase {
  virtual ~ValueBase() {}
  virtual int dim() const = 0;
  virtual std::shared_ptr<ValueBase> retract(const Vector &d) const = 0;
};
template <class T> struct GenericValue : ValueBase {
  T v;
  explicit GenericValue(const T &t) : v(t) {}
  int dim() const override { return traits<T>::dim(v); }
  std::shared_ptr<ValueBase> retract(const Vector &d) const override {
    return std::make_shared<GenericValue<T>>(traits<T>::retract(v, d));
  }
};
class Values {
  std::map<Key, std::shared_ptr<ValueBase>> m_;

public:
  typedef std::map<Key, std::shared_ptr<ValueBase>>::const_iterator const_iterator;
  template <class T> void insert(Key k, const T &v) {
    if (m_.count(k))
      Eigen::mini_fail("Values::insert: key already exists");
    m_[k] = std::make_shared<GenericValue<T>>(v);
  }
  template <class T> const T &at(Key k) const {
    auto it = m_.find(k);
    if (it == m_.end())
      Eigen::mini_fail("Values::at: key does not exist");
    const GenericValue<T> *g = dynamic_cast<const GenericValue<T> *>(it->second.get());
    if (!g)
      Eigen::mini_fail("Values::at: wrong type");
    return g->v;
  }
  const_iterator find(Key k) const { return m_.find(k); }
  const_iterator begin() const { return m_.begin(); }
  const_iterator end() const { return m_.end(); }
  void erase(Key k) { m_.erase(k); }
  void clear() { m_.clear(); }
  size_t size() const { return m_.size(); }
  void set_raw(Key k, const std::shared_ptr<ValueBase> &v) { m_[k] = v; }
};

template <int N, int M> class OptionalJacobian {
  Matrix m_;

public:
  template <class E> OptionalJacobian(const Eigen::Base<E> &m) : m_(m) {
    if (m_.rows() != N || m_.cols() != M)
      Eigen::mini_fail("OptionalJacobian: size mismatch");
  }
  const Matrix &operator*() const { return m_; }
};

namespace noiseModel {
struct Base {
  virtual ~Base() {}
  virtual void print(const std::string &s) const { std::cout << s << "(shim noise model)" << std::endl; }
};
struct Gaussian : Base {
  Matrix covariance; // remembered, never applied here
  static std::shared_ptr<Gaussian> Covariance(const Matrix &c) {
    auto g = std::make_shared<Gaussian>();
    g->covariance = c;
    return g;
  }
};
namespace mEstimator {
struct Huber {
  double k;
  static std::shared_ptr<Huber> Create(double k) {
    auto h = std::make_shared<Huber>();
    h->k = k;
    return h;
  }
};
} // namespace mEstimator
struct Robust : Base {
  std::shared_ptr<mEstimator::Huber> robust;
  std::shared_ptr<Gaussian> noise;
  static std::shared_ptr<Robust> Create(const std::shared_ptr<mEstimator::Huber> &r, const std::shared_ptr<Gaussian> &n) {
    auto o = std::make_shared<Robust>();
    o->robust = r;
    o->noise = n;
    return o;
  }
};
} // namespace noiseModel
typedef std::shared_ptr<noiseModel::Base> SharedNoiseModel;

class NonlinearFactor {
public:
  virtual ~NonlinearFactor() {}
};
class NoiseModelFactor : public NonlinearFactor {
protected:
  SharedNoiseModel noiseModel_;
  Key keys_[4];
  int nkeys_;

public:
  NoiseModelFactor(const SharedNoiseModel &n, int nkeys, Key k1, Key k2, Key k3, Key k4) : noiseModel_(n), nkeys_(nkeys) {
    keys_[0] = k1, keys_[1] = k2, keys_[2] = k3, keys_[3] = k4;
  }
  int nkeys() const { return nkeys_; }
  Key key(int i) const { return keys_[i]; }
  // error vector as evaluateError returns it, and (H != nullptr) one Jacobian per key
  virtual Vector unwhitenedError(const Values &v, std::vector<Matrix> *H) const = 0;
  Key key1() const { return keys_[0]; }
  Key key2() const { return keys_[1]; }
  Key key3() const { return keys_[2]; }
  Key key4() const { return keys_[3]; }
  const SharedNoiseModel &noiseModel() const { return noiseModel_; }
  bool equals(const NoiseModelFactor &o, double) const {
    return keys_[0] == o.keys_[0] && keys_[1] == o.keys_[1] && keys_[2] == o.keys_[2] && keys_[3] == o.keys_[3];
  }
};
template <class A, class B, class C> class NoiseModelFactor3 : public NoiseModelFactor {
public:
  NoiseModelFactor3(const SharedNoiseModel &n, Key k1, Key k2, Key k3) : NoiseModelFactor(n, 3, k1, k2, k3, 0) {}
  virtual Vector evaluateError(const A &, const B &, const C &, boost::optional<Matrix &> H1 = boost::none,
                               boost::optional<Matrix &> H2 = boost::none, boost::optional<Matrix &> H3 = boost::none) const = 0;
  Vector unwhitenedError(const Values &v, std::vector<Matrix> *H) const override {
    if (!H)
      return evaluateError(v.at<A>(keys_[0]), v.at<B>(keys_[1]), v.at<C>(keys_[2]));
    H->assign(3, Matrix());
    return evaluateError(v.at<A>(keys_[0]), v.at<B>(keys_[1]), v.at<C>(keys_[2]), (*H)[0], (*H)[1], (*H)[2]);
  }
};
template <class A, class B, class C, class D> class NoiseModelFactor4 : public NoiseModelFactor {
public:
  NoiseModelFactor4(const SharedNoiseModel &n, Key k1, Key k2, Key k3, Key k4) : NoiseModelFactor(n, 4, k1, k2, k3, k4) {}
  virtual Vector evaluateError(const A &, const B &, const C &, const D &, boost::optional<Matrix &> H1 = boost::none,
                               boost::optional<Matrix &> H2 = boost::none, boost::optional<Matrix &> H3 = boost::none,
                               boost::optional<Matrix &> H4 = boost::none) const = 0;
  Vector unwhitenedError(const Values &v, std::vector<Matrix> *H) const override {
    if (!H)
      return evaluateError(v.at<A>(keys_[0]), v.at<B>(keys_[1]), v.at<C>(keys_[2]), v.at<D>(keys_[3]));
    H->assign(4, Matrix());
    return evaluateError(v.at<A>(keys_[0]), v.at<B>(keys_[1]), v.at<C>(keys_[2]), v.at<D>(keys_[3]), (*H)[0], (*H)[1], (*H)[2],
                         (*H)[3]);
  }
};

} // namespace gtsam
