// ORACLE / TEST INFRASTRUCTURE -- not product code.
//
// What src/solver/ViconGraphSolver.{h,cpp} needs from GTSAM beyond mini_gtsam.h, so that the reference's
// solver class compiles where it lies and runs end to end: Symbol shorthands, NonlinearFactorGraph,
// LevenbergMarquardtParams / LevenbergMarquardtOptimizer.
//
// UNLIKE mini_gtsam.h THIS HEADER CONTAINS ARITHMETIC THAT IS NOT THE REFERENCE'S: a second, independent
// restatement (the first is oracle/solver.hpp) of the GTSAM 4.2a5 pieces the reference calls --
//   noiseModel::Gaussian::Covariance(P)  -> Information(P^-1): R = chol_upper(P^-1), whiten e -> R e;
//                                           Covariance(I) -> Unit
//   noiseModel::Robust(Huber k, noise)   -> WhitenSystem: whiten, then scale [A b] by sqrt(w(|b|)),
//                                           w = 1 (|b| <= k) else k / |b|; loss = rho_huber(|whitened e|)
//   NonlinearFactorGraph::error          -> sum of 0.5 |R e|^2 resp. the robust loss
//   LevenbergMarquardtOptimizer          -> defaultOptimize / iterate / tryLambda / checkConvergence with the
//                                           fixed lambda factor, as configured at ViconGraphSolver.cpp:547-555
// written generically (any factor, any node type, normal equations in skyline storage under a degree ordering,
// scalar Cholesky), i.e. structurally unlike both the oracle's block-tridiagonal solver and the CUDA path.  An end-to-end run of the
// reference's ViconGraphSolver over this driver therefore pins everything the reference itself owns -- timestamp
// cleaning, state erasure, initial guess, graph construction, relinearisation loop, CSV writer -- and cross-checks
// the GTSAM restatement by a second implementation; it does not replace a run against real GTSAM.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>
#include <set>

#include "mini_gtsam.h"

namespace gtsam {

namespace symbol_shorthand {
inline Key mini_symbol(char c, std::uint64_t j) { return (static_cast<Key>(static_cast<unsigned char>(c)) << 56) | j; }
inline Key C(std::uint64_t j) { return mini_symbol('c', j); }
inline Key G(std::uint64_t j) { return mini_symbol('g', j); }
inline Key T(std::uint64_t j) { return mini_symbol('t', j); }
inline Key X(std::uint64_t j) { return mini_symbol('x', j); }
} // namespace symbol_shorthand

// ---- noise-model arithmetic (GTSAM restated, see header)
struct MiniWhitener {
  bool unit = true;
  Matrix R;        // square-root information (upper), when !unit
  double huber_k = -1; // < 0: not robust
  explicit MiniWhitener(const SharedNoiseModel &nm) {
    std::shared_ptr<noiseModel::Gaussian> g = std::dynamic_pointer_cast<noiseModel::Gaussian>(nm);
    if (std::shared_ptr<noiseModel::Robust> r = std::dynamic_pointer_cast<noiseModel::Robust>(nm)) {
      huber_k = r->robust->k;
      g = r->noise;
    }
    if (!g)
      Eigen::mini_fail("unknown noise model");
    const Matrix &P = g->covariance;
    bool ident = true;
    for (int i = 0; i < P.rows(); i++)
      for (int j = 0; j < P.cols(); j++)
        if (P(i, j) != (i == j ? 1.0 : 0.0))
          ident = false;
    unit = ident;
    if (!unit) {
      Matrix info = P.inverse();
      R = info.llt().matrixL().transpose();
    }
  }
  Vector whiten(const Vector &e) const { return unit ? e : Vector(R * e); }
  Matrix whiten(const Matrix &A) const { return unit ? A : Matrix(R * A); }
  double loss(const Vector &e) const {
    const Vector w = whiten(e);
    const double sq = w.squaredNorm();
    if (huber_k < 0)
      return 0.5 * sq;
    const double d = std::sqrt(sq);
    return d <= huber_k ? 0.5 * d * d : huber_k * (d - 0.5 * huber_k);
  }
  double sqrt_weight(const Vector &whitened) const {
    if (huber_k < 0)
      return 1.0;
    const double d = whitened.norm();
    return std::sqrt(d <= huber_k ? 1.0 : huber_k / d);
  }
};

class NonlinearFactorGraph {
  std::vector<std::shared_ptr<NoiseModelFactor>> f_;

public:
  typedef std::vector<std::shared_ptr<NoiseModelFactor>>::iterator iterator;
  template <class F> void add(const F &f) { f_.push_back(std::make_shared<F>(f)); }
  iterator begin() { return f_.begin(); }
  iterator end() { return f_.end(); }
  void erase(iterator a, iterator b) { f_.erase(a, b); }
  size_t nrFactors() const { return f_.size(); }
  size_t size() const { return f_.size(); }
  const NoiseModelFactor &at(size_t i) const { return *f_[i]; }
  std::set<Key> keys() const {
    std::set<Key> k;
    for (const auto &f : f_)
      for (int i = 0; i < f->nkeys(); i++)
        k.insert(f->key(i));
    return k;
  }
  double error(const Values &v) const {
    double e = 0;
    for (const auto &f : f_)
      e += MiniWhitener(f->noiseModel()).loss(f->unwhitenedError(v, nullptr));
    return e;
  }
};

struct NonlinearOptimizerParams {
  enum Verbosity { SILENT, TERMINATION, ERROR, VALUES, DELTA, LINEAR };
};
struct Ordering {
  enum OrderingType { COLAMD, METIS, NATURAL, CUSTOM };
};
struct LevenbergMarquardtParams {
  NonlinearOptimizerParams::Verbosity verbosity = NonlinearOptimizerParams::SILENT;
  Ordering::OrderingType orderingType = Ordering::COLAMD;
  size_t maxIterations = 100;
  double relativeErrorTol = 1e-5, absoluteErrorTol = 1e-5, errorTol = 0.0;
  double lambdaInitial = 1e-5, lambdaFactor = 10.0, lambdaUpperBound = 1e5, lambdaLowerBound = 0.0;
  double minModelFidelity = 1e-3;
};

// record of the optimiser runs of this process, for the wrapper (ViconGraphSolver keeps its optimiser local)
struct MiniLmLog {
  int iterations = 0, tries = 0, linearizations = 0, runs = 0;
  double initial_error = 0, final_error = 0, lambda = 0;
  int max_iterations_override = -1; // >= 0: replaces params.maxIterations (0 = hand back the initial values)
  std::vector<double> trace;        // per damped solve: lambda, error before, error after (inf if not evaluated), linear change, accepted
  static MiniLmLog &instance() {
    static MiniLmLog l;
    return l;
  }
};

class LevenbergMarquardtOptimizer {
  const NonlinearFactorGraph &graph_;
  Values values_;
  LevenbergMarquardtParams p_;
  double error_ = 0, lambda_ = 0;
  size_t iterations_ = 0;

  // Normal equations in skyline (profile) storage: row i keeps columns first_[i]..i of the lower triangle.  Variables
  // are ordered by how many factors touch them (ties by key), so the states form a band and the few variables every
  // factor touches (calibration, time offset, gravity) come last: the Cholesky factor has the same profile.
  struct Lin {
    std::vector<std::vector<Matrix>> A; // whitened (and robust-reweighted) Jacobians per factor and key
    std::vector<Vector> b;              // whitened -error
    std::vector<double> H, g;           // skyline normal equations, gradient
    double error0 = 0;
  };
  std::map<Key, int> offset_;
  std::vector<int> first_;
  std::vector<size_t> rowptr_;
  int n_ = 0;

  double &Hat(std::vector<double> &H, int i, int j) const { return H[rowptr_[i] + (j - first_[i])]; }
  double Hat(const std::vector<double> &H, int i, int j) const { return H[rowptr_[i] + (j - first_[i])]; }

  void layout() {
    std::map<Key, int> degree;
    for (auto it = values_.begin(); it != values_.end(); ++it)
      degree[it->first] = 0;
    for (size_t f = 0; f < graph_.size(); f++)
      for (int a = 0; a < graph_.at(f).nkeys(); a++)
        degree[graph_.at(f).key(a)]++;
    std::vector<std::pair<int, Key>> order;
    for (const auto &d : degree)
      order.push_back({d.second, d.first});
    std::sort(order.begin(), order.end());
    offset_.clear();
    n_ = 0;
    for (const auto &o : order) {
      offset_[o.second] = n_;
      n_ += values_.find(o.second)->second->dim();
    }
    first_.assign(n_, 0);
    for (int i = 0; i < n_; i++)
      first_[i] = i;
    for (size_t f = 0; f < graph_.size(); f++) {
      const NoiseModelFactor &F = graph_.at(f);
      int lo = n_;
      for (int a = 0; a < F.nkeys(); a++)
        lo = std::min(lo, offset_.at(F.key(a)));
      for (int a = 0; a < F.nkeys(); a++) {
        const int o = offset_.at(F.key(a)), d = values_.find(F.key(a))->second->dim();
        for (int i = o; i < o + d; i++)
          first_[i] = std::min(first_[i], lo);
      }
    }
    rowptr_.assign(n_ + 1, 0);
    for (int i = 0; i < n_; i++)
      rowptr_[i + 1] = rowptr_[i] + (size_t)(i - first_[i] + 1);
  }
  void linearize(Lin &L) const {
    const size_t m = graph_.size();
    L.A.assign(m, {});
    L.b.assign(m, Vector());
    L.H.assign(rowptr_[n_], 0.0);
    L.g.assign(n_, 0.0);
    L.error0 = 0;
    for (size_t f = 0; f < m; f++) {
      const NoiseModelFactor &F = graph_.at(f);
      std::vector<Matrix> H;
      const Vector e = F.unwhitenedError(values_, &H);
      MiniWhitener W(F.noiseModel());
      Vector b = W.whiten(Vector(-e));
      const double sw = W.sqrt_weight(b);
      b = Vector(b * sw);
      for (Matrix &h : H)
        h = Matrix(W.whiten(h) * sw);
      L.error0 += 0.5 * b.squaredNorm();
      for (int a = 0; a < F.nkeys(); a++) {
        const int oa = offset_.at(F.key(a));
        const Matrix &Aa = H[a];
        for (int c = 0; c < Aa.cols(); c++) {
          double s = 0;
          for (int r = 0; r < Aa.rows(); r++)
            s += Aa(r, c) * b(r);
          L.g[oa + c] += s;
        }
        for (int k = 0; k < F.nkeys(); k++) {
          const int ok = offset_.at(F.key(k));
          const Matrix &Ak = H[k];
          for (int c = 0; c < Aa.cols(); c++)
            for (int d = 0; d < Ak.cols(); d++) {
              if (ok + d > oa + c)
                continue; // lower triangle only
              double s = 0;
              for (int r = 0; r < Aa.rows(); r++)
                s += Aa(r, c) * Ak(r, d);
              Hat(L.H, oa + c, ok + d) += s;
            }
        }
      }
      L.A[f] = H;
      L.b[f] = b;
    }
  }
  // (H + lambda I) delta = g by skyline Cholesky; false if not positive definite
  bool solve(const Lin &L, double lambda, std::vector<double> &delta) const {
    const int n = n_;
    std::vector<double> M(L.H);
    for (int i = 0; i < n; i++)
      Hat(M, i, i) += lambda;
    for (int i = 0; i < n; i++) {
      for (int j = first_[i]; j <= i; j++) {
        double s = Hat(M, i, j);
        for (int k = std::max(first_[i], first_[j]); k < j; k++)
          s -= Hat(M, i, k) * Hat(M, j, k);
        if (j < i) {
          Hat(M, i, j) = s / Hat(M, j, j);
        } else {
          if (!(s > 0))
            return false;
          Hat(M, i, i) = std::sqrt(s);
        }
      }
    }
    delta.assign(L.g.begin(), L.g.end());
    for (int i = 0; i < n; i++) {
      double s = delta[i];
      for (int k = first_[i]; k < i; k++)
        s -= Hat(M, i, k) * delta[k];
      delta[i] = s / Hat(M, i, i);
    }
    for (int i = n - 1; i >= 0; i--) {
      delta[i] /= Hat(M, i, i);
      for (int k = first_[i]; k < i; k++)
        delta[k] -= Hat(M, i, k) * delta[i];
    }
    return true;
  }
  // GaussianFactorGraph::error(delta) = sum 0.5 |A delta - b|^2
  double linear_error(const Lin &L, const std::vector<double> &delta) const {
    double e = 0;
    for (size_t f = 0; f < graph_.size(); f++) {
      const NoiseModelFactor &F = graph_.at(f);
      Vector r = Vector(-L.b[f]);
      for (int a = 0; a < F.nkeys(); a++) {
        const int oa = offset_.at(F.key(a));
        const Matrix &A = L.A[f][a];
        for (int i = 0; i < A.rows(); i++) {
          double s = 0;
          for (int c = 0; c < A.cols(); c++)
            s += A(i, c) * delta[oa + c];
          r(i) += s;
        }
      }
      e += 0.5 * r.squaredNorm();
    }
    return e;
  }
  Values retract(const std::vector<double> &delta) const {
    Values out;
    for (auto it = values_.begin(); it != values_.end(); ++it) {
      const int o = offset_.at(it->first), d = it->second->dim();
      Vector dv = Vector::Zero(d, 1);
      for (int i = 0; i < d; i++)
        dv(i) = delta[o + i];
      out.set_raw(it->first, it->second->retract(dv));
    }
    return out;
  }
  // LevenbergMarquardtOptimizer::tryLambda: true = leave the lambda search of this iteration
  bool tryLambda(const Lin &L) {
    MiniLmLog::instance().tries++;
    bool step_ok = false, stop_search = false;
    std::vector<double> delta;
    Values newValues;
    double newError = std::numeric_limits<double>::infinity(), linChange = 0;
    if (solve(L, lambda_, delta)) {
      const double oldLin = L.error0;
      const double newLin = linear_error(L, delta);
      linChange = oldLin - newLin;
      if (linChange >= 0) {
        newValues = retract(delta);
        newError = graph_.error(newValues);
        const double costChange = error_ - newError;
        if (linChange > std::numeric_limits<double>::epsilon() * oldLin)
          step_ok = (costChange / linChange) > p_.minModelFidelity;
        if (std::fabs(costChange) < p_.relativeErrorTol * error_)
          stop_search = true;
      }
    }
    for (double v : {lambda_, error_, newError, linChange, step_ok ? 1.0 : 0.0})
      MiniLmLog::instance().trace.push_back(v);
    if (step_ok) {
      values_ = newValues;
      error_ = newError;
      lambda_ = std::max(p_.lambdaLowerBound, lambda_ / p_.lambdaFactor);
      iterations_++;
      return true;
    } else if (!stop_search) {
      lambda_ *= p_.lambdaFactor;
      return lambda_ >= p_.lambdaUpperBound; // "giving up because cannot decrease error with maximum lambda"
    }
    return true;
  }

public:
  LevenbergMarquardtOptimizer(const NonlinearFactorGraph &graph, const Values &initial, const LevenbergMarquardtParams &p)
      : graph_(graph), values_(initial), p_(p) {
    if (MiniLmLog::instance().max_iterations_override >= 0)
      p_.maxIterations = (size_t)MiniLmLog::instance().max_iterations_override;
    lambda_ = p_.lambdaInitial;
    error_ = graph_.error(values_);
  }
  size_t iterations() const { return iterations_; }
  double error() const { return error_; }
  double lambda() const { return lambda_; }
  // NonlinearOptimizer::defaultOptimize
  const Values &optimize() {
    MiniLmLog &log = MiniLmLog::instance();
    if (log.runs == 0)
      log.initial_error = error_;
    log.runs++;
    layout();
    double currentError = error_;
    if (!(currentError <= p_.errorTol) && iterations_ < p_.maxIterations) {
      bool go;
      do {
        currentError = error_;
        Lin L;
        linearize(L);
        log.linearizations++;
        while (!tryLambda(L)) {
        }
        const double newError = error_;
        bool converged;
        if (p_.errorTol >= newError) {
          converged = true;
        } else {
          const double absDec = currentError - newError, relDec = absDec / currentError;
          converged = (p_.relativeErrorTol != 0 && relDec <= p_.relativeErrorTol) || (absDec <= p_.absoluteErrorTol);
        }
        go = iterations_ < p_.maxIterations && !converged && std::isfinite(currentError);
      } while (go);
    }
    log.iterations += (int)iterations_;
    log.final_error = error_;
    log.lambda = lambda_;
    return values_;
  }
};

} // namespace gtsam
