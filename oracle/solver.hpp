// ORACLE (test infrastructure, NOT product code) -- parity unpinned for this file (GTSAM's LM and the solver glue).
//
// CPU restatement of ViconGraphSolver (src/solver/ViconGraphSolver.{h,cpp}) with the
// GTSAM 4.2a5 Levenberg-Marquardt policy the reference configures in
// ViconGraphSolver.cpp:547-567 restated from GTSAM's published algorithm
// (LevenbergMarquardtOptimizer::iterate / tryLambda, NonlinearOptimizer::defaultOptimize,
// checkConvergence).  GTSAM is a third-party dependency that is NOT vendored by the
// reference (ReadMe.md:93 pins "4.2a5" in prose only) and the reference ships no tests,
// so there are no golden vectors to pin this restatement against: PARITY UNPINNED.
//
// Linear algebra: the damped normal equations (block-tridiagonal 15x15 chain + 9-wide
// arrow) are solved with an exact sequential block Cholesky.  GTSAM uses METIS ordering +
// multifrontal Cholesky: same solution up to round-off.
#pragma once
#include "../include/vicon2gt_b200.h"
#include "factors.hpp"
#include <cstdio>
#include <fstream>
#include <iomanip>
#include <sstream>
#include <string>
#include <sys/stat.h>
#include <vector>

namespace orc {

struct LinearSystem {
  // normal-equation blocks (undamped), unknown order X(0..n-1) x15 then 9 globals
  std::vector<Mat15> D, O;        // O[i] couples i and i+1 (block (i, i+1)); O[n-1] = 0
  std::vector<Mat<15, 9>> B;
  std::vector<Vec15> g;
  Mat<9, 9> Gamma;
  Mat<9, 1> gG;
  double lin_error0 = 0; // 1/2 |b|^2 of the whitened (and Huber-reweighted) system
  // whitened Jacobians per factor (for the GTSAM-style linear.error(delta))
  std::vector<Mat<15, 15>> A1, A2;
  std::vector<Mat<15, 2>> A3;
  std::vector<Vec15> b_imu;
  std::vector<Mat<6, 15>> V1;
  std::vector<Mat<6, 7>> V2;
  std::vector<Vec6> b_vic;
};

struct LmTraceRow {
  double lambda, cost_before, cost_after, lin_cost_change;
  int accepted;
};

class ViconGraphSolver {
public:
  ViconGraphSolver(const v2gt_params &p, const Propagator &prop, const Interpolator &interp, const std::vector<double> &timestamp_cameras_)
      : params(p), propagator(prop), interpolator(interp), timestamp_cameras(timestamp_cameras_) {
    cfg.toff = p.estimate_toff_vicon_to_imu != 0;
    cfg.ori = p.estimate_ori_vicon_to_imu != 0;
    cfg.pos = p.estimate_pos_vicon_to_imu != 0;
    for (int i = 0; i < 9; i++) {
      init_R_GtoV.a[i] = p.R_GtoV[i];
      init_R_BtoI.a[i] = p.R_BtoI[i];
    }
    for (int i = 0; i < 3; i++)
      init_p_BinI(i) = p.p_BinI[i];
  }

  // ---- options (oracle only)
  bool faithful_scan = false;   // O(M) scans of Propagator (reference-faithful timing mode)
  bool faithful_linerr = true;  // linear.error(delta) per factor (GTSAM) instead of the closed form

  v2gt_params params;
  Propagator propagator;
  Interpolator interpolator;
  std::vector<double> timestamp_cameras;
  EstimateFlags cfg;
  Mat3 init_R_GtoV, init_R_BtoI;
  Vec3 init_p_BinI;

  std::vector<NavState> states; // aligned with timestamp_cameras (the reference's Values X(id))
  Globals gl;
  std::vector<ImuFactorData> imu; // imu[i] links states[i] -> states[i+1]
  v2gt_trial_info info{};
  std::vector<LmTraceRow> trace;
  bool have_values = false;
  double time_build = 0, time_opt = 0;

  // src/solver/ViconGraphSolver.cpp:96-192
  int build_and_solve() {
    info = v2gt_trial_info{};
    if (timestamp_cameras.empty())
      return info.status = V2GT_ERR_NO_TIMESTAMPS;
    if (interpolator.size() < 2)
      return info.status = V2GT_ERR_NO_VICON;
    clean_timestamps();
    if (timestamp_cameras.empty())
      return info.status = V2GT_ERR_ALL_OUT_OF_RANGE;
    states.clear();
    have_values = false;
    for (int i = 0; i <= params.num_loop_relin; i++) {
      int st = build_problem(i == 0);
      if (st != V2GT_OK)
        return info.status = st;
      optimize_problem();
    }
    return info.status = V2GT_OK;
  }

  // src/solver/ViconGraphSolver.cpp:114-142
  void clean_timestamps() {
    const double first = interpolator.get_time_min() + params.toff_imu_to_vicon + 2 * params.time_offset_window;
    const double last = interpolator.get_time_max() + params.toff_imu_to_vicon - 2 * params.time_offset_window;
    std::vector<double> keep;
    keep.reserve(timestamp_cameras.size());
    for (double t : timestamp_cameras) {
      if (!propagator.has_bounding_imu(t, faithful_scan))
        info.removed_no_imu++;
      else if (t < first)
        info.removed_before++;
      else if (t > last)
        info.removed_after++;
      else
        keep.push_back(t);
    }
    timestamp_cameras.swap(keep);
  }

  // src/solver/ViconGraphSolver.cpp:390-536
  int build_problem(bool init_states) {
    if (init_states) {
      gl.q_BtoI = rot_2_quat(init_R_BtoI);
      gl.p_BinI = init_p_BinI;
      Vec3 rpy = rot2rpy(init_R_GtoV);
      gl.g.theta_x = rpy(0);
      gl.g.theta_y = rpy(1);
      gl.t_off = params.toff_imu_to_vicon;
      states.assign(timestamp_cameras.size(), NavState());
    }
    imu.clear();
    const double TO = params.time_offset_window;
    size_t k = 0; // index into timestamp_cameras/states (the reference's iterator it1)
    while (k < timestamp_cameras.size()) {
      const double timestamp_inI = timestamp_cameras[k];
      const double timestamp_inV = timestamp_inI - gl.t_off;
      Interpolator::Result r0 = interpolator.get_pose(timestamp_inV - TO);
      Interpolator::Result r2 = interpolator.get_pose(timestamp_inV + TO);
      Interpolator::Result r1 = interpolator.get_pose(timestamp_inV);
      bool drop = (!r0.ok || !r2.ok || !r1.ok);
      if (!drop) {
        // R_vicon holds the third lookup's covariance (.cpp:453)
        if (std::isnan(r1.R.norm()) || std::isnan(inverse(r1.R).norm()))
          drop = true;
      }
      if (drop) {
        timestamp_cameras.erase(timestamp_cameras.begin() + k);
        states.erase(states.begin() + k);
        info.removed_no_vicon++;
        continue;
      }
      if (init_states) {
        const Mat3 RBtoI_T = init_R_BtoI.transpose();
        NavState s;
        s.t = timestamp_inI;
        s.q = quat_multiply(rot_2_quat(init_R_BtoI), r1.q);
        s.p = r1.p - quat_2_Rot(Inv(r1.q)) * RBtoI_T * init_p_BinI;
        const Vec3 p_I0inV = r0.p - quat_2_Rot(Inv(r0.q)) * RBtoI_T * init_p_BinI;
        const Vec3 p_I2inV = r2.p - quat_2_Rot(Inv(r2.q)) * RBtoI_T * init_p_BinI;
        s.v = (p_I2inV - p_I0inV) / (2 * TO);
        states[k] = s;
      }
      if (k == 0) {
        k++;
        continue;
      }
      const double time0 = timestamp_cameras[k - 1];
      const double time1 = timestamp_cameras[k];
      CpiV1 preint(0, 0, 0, 0, true);
      const bool has_imu = propagator.propagate(time0, time1, states[k - 1].bg, states[k - 1].ba, preint, faithful_scan);
      if (!has_imu || preint.DT != (time1 - time0))
        return V2GT_ERR_INVALID_PREINT;
      ImuFactorData f;
      f.P = preint.P_meas;
      f.dt = preint.DT;
      f.grav_mag = params.gravity_magnitude;
      f.alpha = preint.alpha_tau;
      f.beta = preint.beta_tau;
      f.q_KtoK1 = preint.q_k2tau;
      f.ba_lin = preint.b_a_lin;
      f.bg_lin = preint.b_w_lin;
      f.J_q = preint.J_q;
      f.J_beta = preint.J_b;
      f.J_alpha = preint.J_a;
      f.H_beta = preint.H_b;
      f.H_alpha = preint.H_a;
      if (f.P.hasNaN() || !make_imu_noise(f) || f.info.hasNaN())
        return V2GT_ERR_NAN_COVARIANCE;
      imu.push_back(f);
      k++;
    }
    have_values = true;
    info.n_states = (int)states.size();
    if (states.empty())
      return V2GT_ERR_ALL_OUT_OF_RANGE;
    return V2GT_OK;
  }

  // ---- cost: NonlinearFactorGraph::error (sum of noise-model losses)
  double total_error(const std::vector<NavState> &xs, const Globals &g) const {
    double e_imu = 0, e_vic = 0;
    for (size_t i = 0; i + 1 < xs.size(); i++) {
      ImuEval ev = eval_imu_factor(imu[i], xs[i], xs[i + 1], g.g, false);
      Vec15 w = imu[i].sqrtinfo * ev.e;
      e_imu += 0.5 * w.squaredNorm();
    }
    for (size_t i = 0; i < xs.size(); i++) {
      ViconEval ev = eval_vicon_factor(interpolator, cfg, xs[i], g, false);
      e_vic += huber_loss(params.huber_k, ev.e.norm());
    }
    return e_imu + e_vic;
  }

  // ---- linearisation: whitened Jacobians -> normal-equation blocks
  void linearize(const std::vector<NavState> &xs, const Globals &g, LinearSystem &ls) const {
    const size_t n = xs.size();
    ls.D.assign(n, Mat15());
    ls.O.assign(n, Mat15());
    ls.B.assign(n, Mat<15, 9>());
    ls.g.assign(n, Vec15());
    ls.Gamma = Mat<9, 9>();
    ls.gG = Mat<9, 1>();
    ls.lin_error0 = 0;
    ls.A1.assign(n ? n - 1 : 0, Mat15());
    ls.A2.assign(n ? n - 1 : 0, Mat15());
    ls.A3.assign(n ? n - 1 : 0, Mat<15, 2>());
    ls.b_imu.assign(n ? n - 1 : 0, Vec15());
    ls.V1.assign(n, Mat<6, 15>());
    ls.V2.assign(n, Mat<6, 7>());
    ls.b_vic.assign(n, Vec6());
    for (size_t i = 0; i + 1 < n; i++) {
      ImuEval ev = eval_imu_factor(imu[i], xs[i], xs[i + 1], g.g, true);
      const Mat15 &R = imu[i].sqrtinfo;
      const Mat15 A1 = R * ev.H1, A2 = R * ev.H2;
      const Mat<15, 2> A3 = R * ev.H3;
      const Vec15 b = -(R * ev.e);
      ls.A1[i] = A1;
      ls.A2[i] = A2;
      ls.A3[i] = A3;
      ls.b_imu[i] = b;
      ls.lin_error0 += 0.5 * b.squaredNorm();
      ls.D[i] += A1.transpose() * A1;
      ls.D[i + 1] += A2.transpose() * A2;
      ls.O[i] += A1.transpose() * A2;
      const Mat<15, 2> B1 = A1.transpose() * A3, B2 = A2.transpose() * A3;
      const Mat<2, 2> GG = A3.transpose() * A3;
      const Vec15 g1 = A1.transpose() * b, g2 = A2.transpose() * b;
      const Mat<2, 1> gg = A3.transpose() * b;
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < 2; c++) {
          ls.B[i](r, 7 + c) += B1(r, c);
          ls.B[i + 1](r, 7 + c) += B2(r, c);
        }
      for (int r = 0; r < 2; r++) {
        for (int c = 0; c < 2; c++)
          ls.Gamma(7 + r, 7 + c) += GG(r, c);
        ls.gG(7 + r) += gg(r);
      }
      ls.g[i] += g1;
      ls.g[i + 1] += g2;
    }
    for (size_t i = 0; i < n; i++) {
      ViconEval ev = eval_vicon_factor(interpolator, cfg, xs[i], g, true);
      const double sw = std::sqrt(huber_weight(params.huber_k, ev.e.norm()));
      Mat<6, 15> V1 = sw * ev.H1;
      Mat<6, 7> V2;
      for (int r = 0; r < 6; r++) {
        for (int c = 0; c < 3; c++) {
          V2(r, c) = sw * ev.H2(r, c);
          V2(r, 3 + c) = sw * ev.H3(r, c);
        }
        V2(r, 6) = sw * ev.H4(r);
      }
      const Vec6 b = -(sw * ev.e);
      ls.V1[i] = V1;
      ls.V2[i] = V2;
      ls.b_vic[i] = b;
      ls.lin_error0 += 0.5 * b.squaredNorm();
      ls.D[i] += V1.transpose() * V1;
      const Mat<15, 7> Bv = V1.transpose() * V2;
      const Mat<7, 7> Gv = V2.transpose() * V2;
      const Vec15 gv = V1.transpose() * b;
      const Mat<7, 1> ggv = V2.transpose() * b;
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < 7; c++)
          ls.B[i](r, c) += Bv(r, c);
      for (int r = 0; r < 7; r++) {
        for (int c = 0; c < 7; c++)
          ls.Gamma(r, c) += Gv(r, c);
        ls.gG(r) += ggv(r);
      }
      ls.g[i] += gv;
    }
  }

  // ---- (H + lambda I) delta = g : sequential block-tridiagonal + arrow Cholesky
  bool solve(const LinearSystem &ls, double lambda, std::vector<double> &delta) const {
    const size_t n = ls.D.size();
    delta.assign(15 * n + 9, 0.0);
    std::vector<Mat15> L(n), WO(n);
    std::vector<Mat<15, 9>> WB(n);
    std::vector<Vec15> wg(n);
    Mat15 Dc = ls.D.empty() ? Mat15() : ls.D[0];
    Mat<15, 9> Bc = ls.B.empty() ? Mat<15, 9>() : ls.B[0];
    Vec15 gc = ls.g.empty() ? Vec15() : ls.g[0];
    Mat<9, 9> Gam = ls.Gamma;
    Mat<9, 1> gG = ls.gG;
    for (size_t i = 0; i < n; i++) {
      for (int d = 0; d < 15; d++)
        Dc(d, d) += lambda;
      if (!chol_lower(Dc, L[i]))
        return false;
      const Mat15 Li = inv_lower(L[i]);
      WB[i] = Li * Bc;
      wg[i] = Li * gc;
      Gam -= WB[i].transpose() * WB[i];
      gG -= WB[i].transpose() * wg[i];
      if (i + 1 < n) {
        WO[i] = Li * ls.O[i];
        Dc = ls.D[i + 1] - WO[i].transpose() * WO[i];
        Bc = ls.B[i + 1] - WO[i].transpose() * WB[i];
        gc = ls.g[i + 1] - WO[i].transpose() * wg[i];
      }
    }
    for (int d = 0; d < 9; d++)
      Gam(d, d) += lambda;
    Gam = 0.5 * (Gam + Gam.transpose());
    Mat<9, 9> LG;
    if (!chol_lower(Gam, LG))
      return false;
    const Mat<9, 9> LGi = inv_lower(LG);
    const Mat<9, 1> dG = LGi.transpose() * (LGi * gG);
    for (int d = 0; d < 9; d++)
      delta[15 * n + d] = dG(d);
    Vec15 next;
    for (size_t ii = n; ii-- > 0;) {
      Vec15 rhs = wg[ii] - WB[ii] * dG;
      if (ii + 1 < n)
        rhs -= WO[ii] * next;
      // solve L^T x = rhs
      Vec15 x;
      for (int r = 14; r >= 0; r--) {
        double s = rhs(r);
        for (int c = r + 1; c < 15; c++)
          s -= L[ii](c, r) * x(c);
        x(r) = s / L[ii](r, r);
      }
      for (int d = 0; d < 15; d++)
        delta[15 * ii + d] = x(d);
      next = x;
    }
    return true;
  }

  // GaussianFactorGraph::error(delta) = sum 1/2 |A delta - b|^2 (undamped system)
  double linear_error(const LinearSystem &ls, const std::vector<double> &delta) const {
    const size_t n = ls.D.size();
    auto dx = [&](size_t i) {
      Vec15 v;
      for (int d = 0; d < 15; d++)
        v(d) = delta[15 * i + d];
      return v;
    };
    Mat<7, 1> dC;
    Mat<2, 1> dGr;
    for (int d = 0; d < 7; d++)
      dC(d) = delta[15 * n + d];
    dGr(0) = delta[15 * n + 7];
    dGr(1) = delta[15 * n + 8];
    double err = 0;
    for (size_t i = 0; i + 1 < n; i++) {
      Vec15 r = ls.A1[i] * dx(i) + ls.A2[i] * dx(i + 1) + ls.A3[i] * dGr - ls.b_imu[i];
      err += 0.5 * r.squaredNorm();
    }
    for (size_t i = 0; i < n; i++) {
      Vec6 r = ls.V1[i] * dx(i) + ls.V2[i] * dC - ls.b_vic[i];
      err += 0.5 * r.squaredNorm();
    }
    return err;
  }

  void apply_delta(const std::vector<NavState> &xs, const Globals &g, const std::vector<double> &delta, std::vector<NavState> &xn,
                   Globals &gn) const {
    const size_t n = xs.size();
    xn.resize(n);
    for (size_t i = 0; i < n; i++) {
      Vec15 xi;
      for (int d = 0; d < 15; d++)
        xi(d) = delta[15 * i + d];
      xn[i] = retract(xs[i], xi);
    }
    const double *dg = &delta[15 * n];
    gn = g;
    gn.q_BtoI = retract_quat(g.q_BtoI, Vec3{dg[0], dg[1], dg[2]});
    for (int d = 0; d < 3; d++)
      gn.p_BinI(d) = g.p_BinI(d) + dg[3 + d];
    gn.t_off = g.t_off + dg[6];
    gn.g = retract_rotxy(g.g, dg[7], dg[8]);
  }

  // src/solver/ViconGraphSolver.cpp:538-570 + GTSAM LM (see file header)
  void optimize_problem() {
    trace.clear();
    double lambda = params.lm_lambda_initial;
    int iterations = 0, tries = 0, linearizations = 0;
    double error = total_error(states, gl);
    info.initial_error = error;
    info.lm_state = 0;
    LinearSystem ls;
    std::vector<double> delta;
    std::vector<NavState> xn;
    Globals gn;
    const double eps = std::numeric_limits<double>::epsilon();
    if (iterations >= params.lm_max_iterations)
      info.lm_state = 1;
    while (info.lm_state == 0) {
      const double currentError = error;
      // ---- iterate(): linearize once, then tryLambda until it returns true
      linearize(states, gl, ls);
      linearizations++;
      bool gave_up = false;
      while (true) {
        if (params.lm_max_tries > 0 && tries >= params.lm_max_tries) {
          info.lm_state = 4;
          break;
        }
        tries++;
        bool step_ok = false, stop_search = false;
        double newError = std::numeric_limits<double>::infinity(), linChange = 0, costChange = 0;
        const bool solved = solve(ls, lambda, delta);
        if (solved) {
          const double oldLin = ls.lin_error0;
          double newLin;
          if (faithful_linerr) {
            newLin = linear_error(ls, delta);
            linChange = oldLin - newLin;
          } else {
            linChange = 0.5 * (gdot(ls, delta) + lambda * sqnorm(delta));
          }
          if (linChange >= 0) {
            apply_delta(states, gl, delta, xn, gn);
            newError = total_error(xn, gn);
            costChange = error - newError;
            if (linChange > eps * oldLin) {
              const double fidelity = costChange / linChange;
              step_ok = fidelity > params.lm_min_model_fidelity;
            }
            const double minAbsTol = params.lm_relative_error_tol * error;
            if (std::abs(costChange) < minAbsTol)
              stop_search = true;
          }
        }
        trace.push_back(LmTraceRow{lambda, error, newError, linChange, step_ok ? 1 : 0});
        if (step_ok) {
          states.swap(xn);
          gl = gn;
          error = newError;
          lambda = std::max(params.lm_lambda_lower_bound, lambda / params.lm_lambda_factor);
          iterations++;
          break;
        } else if (!stop_search) {
          lambda *= params.lm_lambda_factor;
          if (lambda >= params.lm_lambda_upper_bound) {
            gave_up = true;
            break;
          }
        } else {
          break; // relative cost change tiny: stop searching, state unchanged
        }
      }
      if (info.lm_state == 4)
        break;
      // ---- defaultOptimize loop condition
      const double newErr = error;
      if (iterations >= params.lm_max_iterations) {
        info.lm_state = 1;
        break;
      }
      // checkConvergence(relTol, absTol, errTol=0, currentError, newError)
      if (newErr <= 0.0) {
        info.lm_state = 2;
        break;
      }
      const double absDec = currentError - newErr;
      const double relDec = absDec / currentError;
      const bool converged = (params.lm_relative_error_tol != 0 && relDec <= params.lm_relative_error_tol) ||
                             (absDec <= params.lm_absolute_error_tol);
      if (converged) {
        info.lm_state = gave_up ? 3 : 2;
        break;
      }
      if (!std::isfinite(currentError)) {
        info.lm_state = 2;
        break;
      }
    }
    info.iterations += iterations;
    info.tries += tries;
    info.linearizations += linearizations;
    info.final_error = error;
    info.lambda = lambda;
  }

  static double sqnorm(const std::vector<double> &v) {
    double s = 0;
    for (double x : v)
      s += x * x;
    return s;
  }
  static double gdot(const LinearSystem &ls, const std::vector<double> &delta) {
    const size_t n = ls.D.size();
    double s = 0;
    for (size_t i = 0; i < n; i++)
      for (int d = 0; d < 15; d++)
        s += ls.g[i](d) * delta[15 * i + d];
    for (int d = 0; d < 9; d++)
      s += ls.gG(d) * delta[15 * n + d];
    return s;
  }

  // ---- Eigen-style printing used by the info file (operator<< of a dense matrix with the
  // default IOFormat: stream precision, columns aligned to the widest coefficient).
  template <int R, int C> static std::string eigen_print(const Mat<R, C> &m, int precision = 6) {
    std::string cells[R * C];
    size_t width = 0;
    for (int i = 0; i < R * C; i++) {
      std::ostringstream o;
      o << std::setprecision(precision) << m.a[i];
      cells[i] = o.str();
      width = std::max(width, cells[i].size());
    }
    std::ostringstream out;
    for (int r = 0; r < R; r++) {
      if (r)
        out << "\n";
      for (int c = 0; c < C; c++) {
        if (c)
          out << " ";
        out << std::setw((int)width) << cells[r * C + c];
      }
    }
    return out.str();
  }

  // src/solver/ViconGraphSolver.cpp:194-248
  bool write_to_file(const std::string &csvfilepath, const std::string &infofilepath) const {
    std::remove(csvfilepath.c_str());
    std::remove(infofilepath.c_str());
    std::ofstream of_state(csvfilepath, std::ofstream::out | std::ofstream::app);
    if (!of_state)
      return false;
    of_state << "#time(ns),px,py,pz,qw,qx,qy,qz,vx,vy,vz,bwx,bwy,bwz,bax,bay,baz" << std::endl;
    const Mat3 R_GtoV = gl.g.rot();
    const Vec4 q_GtoV = rot_2_quat(R_GtoV);
    for (size_t i = 0; i < states.size(); i++) {
      const NavState &s = states[i];
      const Vec4 q = quat_multiply(s.q, q_GtoV);
      const Vec3 p = R_GtoV.transpose() * s.p;
      const Vec3 v = R_GtoV.transpose() * s.v;
      of_state << std::setprecision(20) << std::floor(1e9 * timestamp_cameras[i]) << "," << std::setprecision(6) << p(0) << "," << p(1)
               << "," << p(2) << "," << q(3) << "," << q(0) << "," << q(1) << "," << q(2) << "," << v(0) << "," << v(1) << "," << v(2)
               << "," << s.bg(0) << "," << s.bg(1) << "," << s.bg(2) << "," << s.ba(0) << "," << s.ba(1) << "," << s.ba(2) << std::endl;
    }
    of_state.close();
    std::ofstream of_info(infofilepath, std::ofstream::out | std::ofstream::app);
    if (!of_info)
      return false;
    Mat<1, 1> toff;
    toff(0) = gl.t_off;
    of_info << "R_BtoI: " << std::endl << eigen_print(quat_2_Rot(gl.q_BtoI)) << std::endl << std::endl;
    of_info << "q_BtoI: " << std::endl << eigen_print(gl.q_BtoI) << std::endl << std::endl;
    of_info << "p_BinI: " << std::endl << eigen_print(gl.p_BinI) << std::endl << std::endl;
    of_info << "R_GtoV: " << std::endl << eigen_print(R_GtoV) << std::endl << std::endl;
    of_info << "R_GtoV (thetax, thetay): " << std::endl;
    of_info << gl.g.theta_x << " " << gl.g.theta_x << std::endl << std::endl; // thetax twice, as in the reference (.cpp:244)
    of_info << "gravity norm: " << std::endl << params.gravity_magnitude << std::endl << std::endl;
    of_info << "t_off_vicon_to_imu: " << std::endl << eigen_print(toff) << std::endl << std::endl;
    of_info.close();
    return true;
  }
};

} // namespace orc
