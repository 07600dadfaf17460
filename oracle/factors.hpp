// ORACLE (test infrastructure, NOT product code) -- node / factor arithmetic pinned on the reference's own sources (oracle/_ref, tests/test_ref_pin_cpu.py); GTSAM noise models unpinned.
//
// Graph nodes and factors, restated from the reference's src/gtsam/*:
//   JPLNavState (15 DoF), JPLQuaternion (3 DoF), RotationXY (2 DoF), Vector3, Vector1 nodes,
//   ImuFactorCPIv1 (15 rows, X_k X_k+1 G) and MeasBased_ViconPoseTimeoffsetFactor (6 rows, X C0 C1 T).
// GTSAM 4.2a5 (not vendored by the reference) supplies the noise models; their
// published semantics are restated here: Gaussian::Covariance(P) whitens with
// R = chol_upper(P^-1); Robust(Huber(1.345), Unit) reweights [A b] by sqrt(w),
// w = 1 if |e|<=k else k/|e|, and its loss is rho(|e|) = e^2/2 or k(|e|-k/2).
#pragma once
#include "meas.hpp"

namespace orc {

// src/gtsam/JPLNavState.h : t; q(4) bg(3) v(3) ba(3) p(3)
struct NavState {
  double t = 0;
  Vec4 q{0, 0, 0, 1};
  Vec3 bg, v, ba, p;
};

// dq from a 3-vector, shared by JPLNavState::retract (JPLNavState.cpp:25-45) and
// JPLQuaternion::retract (JPLQuaternion.cpp:24-44): NaN (zero step) => identity.
inline Vec4 small_quat(const Vec3 &dth) {
  const double n = dth.norm();
  const Vec3 dq13 = (std::sin(n / 2) / n) * dth;
  const double dq4 = std::cos(n / 2);
  Vec4 dq{dq13(0), dq13(1), dq13(2), dq4};
  dq = dq / dq.norm();
  if (dq(3) < 0)
    dq = -dq;
  if (std::isnan(dq.norm()))
    dq = Vec4{0, 0, 0, 1.0};
  return dq;
}

// src/gtsam/JPLNavState.cpp:25-54
inline NavState retract(const NavState &x, const Vec15 &xi) {
  NavState y;
  y.t = x.t;
  y.q = quat_multiply(small_quat(Vec3{xi(0), xi(1), xi(2)}), x.q);
  for (int i = 0; i < 3; i++) {
    y.bg(i) = x.bg(i) + xi(3 + i);
    y.v(i) = x.v(i) + xi(6 + i);
    y.ba(i) = x.ba(i) + xi(9 + i);
    y.p(i) = x.p(i) + xi(12 + i);
  }
  return y;
}
// src/gtsam/JPLQuaternion.cpp:24-49
inline Vec4 retract_quat(const Vec4 &q, const Vec3 &xi) { return quat_multiply(small_quat(xi), q); }

// src/gtsam/RotationXY.{h,cpp}
struct RotXY {
  double theta_x = 0, theta_y = 0;
  Mat3 rot() const { return rot_y(theta_y) * rot_x(theta_x); } // RotationXY.h:62
};
inline RotXY retract_rotxy(const RotXY &g, double dx, double dy) { // RotationXY.cpp:24-32
  RotXY o;
  o.theta_x = wrap2pi(g.theta_x + wrap2pi(dx));
  o.theta_y = wrap2pi(g.theta_y + wrap2pi(dy));
  return o;
}

// The four calibration/global nodes C(0), C(1), T(0), G(0)
struct Globals {
  Vec4 q_BtoI{0, 0, 0, 1};
  Vec3 p_BinI;
  double t_off = 0;
  RotXY g;
};

// src/gtsam/GtsamConfig.h:35-41
struct EstimateFlags {
  bool toff = true, ori = true, pos = true;
};

// Constants of one ImuFactorCPIv1 (ctor argument mapping of
// ViconGraphSolver.cpp:527-530 vs ImuFactorCPIv1.h:71-73: J_b->J_beta, J_a->J_alpha,
// H_b->H_beta, H_a->H_alpha).
struct ImuFactorData {
  double dt = 0, grav_mag = 9.81;
  Vec3 alpha, beta;
  Vec4 q_KtoK1{0, 0, 0, 1};
  Vec3 ba_lin, bg_lin;
  Mat3 J_q, J_beta, J_alpha, H_beta, H_alpha;
  Mat15 P;       // covariance handed to Gaussian::Covariance
  Mat15 info;    // P^-1
  Mat15 sqrtinfo; // R upper triangular, R^T R = P^-1 (GTSAM Gaussian::Information -> LLT.matrixU)
  bool ok = true;
};

// GTSAM noiseModel::Gaussian::Covariance(P): information = P^-1, R = upper Cholesky of it.
// P^-1 is taken through the Cholesky factor of P (P is SPD); the reference's
// Eigen general inverse gives the same matrix up to round-off.
inline bool make_imu_noise(ImuFactorData &f) {
  Mat15 L;
  if (!inverse_spd(f.P, f.info)) {
    f.ok = false;
    return false;
  }
  // enforce exact symmetry of the information matrix before factoring
  f.info = 0.5 * (f.info + f.info.transpose());
  if (!chol_lower(f.info, L)) {
    f.ok = false;
    return false;
  }
  f.sqrtinfo = L.transpose();
  return true;
}

struct ImuEval {
  Vec15 e;                 // unwhitened residual
  Mat15 H1, H2;            // d e / d X_k, d e / d X_k+1
  Mat<15, 2> H3;           // d e / d (theta_x, theta_y)
};

// src/gtsam/ImuFactorCPIv1.cpp:25-216
inline ImuEval eval_imu_factor(const ImuFactorData &f, const NavState &si, const NavState &sj, const RotXY &rotxy, bool jac) {
  ImuEval o;
  const Mat3 I3 = Mat3::Identity();
  const Vec3 ez{0, 0, 1};
  const Vec3 grav_inG = rotxy.rot() * (f.grav_mag * ez);
  const double dt = f.dt;

  const Mat3 ExpB = exp_so3(-(f.J_q * (si.bg - f.bg_lin)));
  const Vec4 q_b = rot_2_quat(ExpB);
  const Vec4 q_n = quat_multiply(sj.q, Inv(si.q));
  const Vec4 q_rminus = quat_multiply(q_n, Inv(f.q_KtoK1));
  const Vec4 q_r = quat_multiply(q_rminus, q_b);
  const Vec4 q_m = quat_multiply(Inv(q_b), f.q_KtoK1);

  const Mat3 Rk = quat_2_Rot(si.q);
  const Vec3 pterm = sj.p - si.p - si.v * dt + 0.5 * grav_inG * std::pow(dt, 2);
  const Vec3 vterm = sj.v - si.v + grav_inG * dt;
  Vec3 alphahat = Rk * pterm;
  alphahat = alphahat - f.J_alpha * (si.bg - f.bg_lin) - f.H_alpha * (si.ba - f.ba_lin);
  Vec3 betahat = Rk * vterm;
  betahat = betahat - f.J_beta * (si.bg - f.bg_lin) - f.H_beta * (si.ba - f.ba_lin);

  for (int i = 0; i < 3; i++) {
    o.e(i) = 2 * q_r(i);
    o.e(3 + i) = sj.bg(i) - si.bg(i);
    o.e(6 + i) = betahat(i) - f.beta(i);
    o.e(9 + i) = sj.ba(i) - si.ba(i);
    o.e(12 + i) = alphahat(i) - f.alpha(i);
  }
  if (!jac)
    return o;

  const Vec3 qn_v{q_n(0), q_n(1), q_n(2)}, qm_v{q_m(0), q_m(1), q_m(2)};
  const Vec3 qrm_v{q_rminus(0), q_rminus(1), q_rminus(2)}, qr_v{q_r(0), q_r(1), q_r(2)};
  // H1, theta column
  o.H1.setBlock(0, 0, -((q_n(3) * I3 - skew_x(qn_v)) * (q_m(3) * I3 - skew_x(qm_v)) + qn_v * qm_v.transpose()));
  o.H1.setBlock(6, 0, skew_x(Rk * vterm));
  o.H1.setBlock(12, 0, skew_x(Rk * pterm));
  // bg column
  o.H1.setBlock(0, 3, (q_rminus(3) * I3 - skew_x(qrm_v)) * f.J_q);
  o.H1.setBlock(3, 3, -I3);
  o.H1.setBlock(6, 3, -f.J_beta);
  o.H1.setBlock(12, 3, -f.J_alpha);
  // v column
  o.H1.setBlock(6, 6, -Rk);
  o.H1.setBlock(12, 6, -dt * Rk);
  // ba column
  o.H1.setBlock(6, 9, -f.H_beta);
  o.H1.setBlock(9, 9, -I3);
  o.H1.setBlock(12, 9, -f.H_alpha);
  // p column
  o.H1.setBlock(12, 12, -Rk);

  o.H2.setBlock(0, 0, q_r(3) * I3 + skew_x(qr_v));
  o.H2.setBlock(3, 3, I3);
  o.H2.setBlock(6, 6, Rk);
  o.H2.setBlock(9, 9, I3);
  o.H2.setBlock(12, 12, Rk);

  const double sx = std::sin(rotxy.theta_x), cx = std::cos(rotxy.theta_x);
  const double sy = std::sin(rotxy.theta_y), cy = std::cos(rotxy.theta_y);
  Mat<3, 2> Hxy;
  Hxy(0, 0) = -sy * sx;
  Hxy(1, 0) = -cx;
  Hxy(2, 0) = -cy * sx;
  Hxy(0, 1) = cy * cx;
  Hxy(1, 1) = 0.0;
  Hxy(2, 1) = -sy * cx;
  o.H3.setBlock(6, 0, (Rk * dt * f.grav_mag) * Hxy);
  o.H3.setBlock(12, 0, (0.5 * Rk * std::pow(dt, 2) * f.grav_mag) * Hxy);
  return o;
}

struct ViconEval {
  bool has_vicon = false;
  long idx0 = -1, idx1 = -1;
  Vec6 e;              // whitened residual  W [2 vec(q_r); p_BinV - p_interp]   (before Huber)
  Mat<6, 15> H1;       // whitened Jacobians (before Huber)
  Mat<6, 3> H2, H3;
  Vec6 H4;
  Mat6 W;
};

// src/gtsam/MeasBased_ViconPoseTimeoffsetFactor.cpp:24-142.
// A failed lookup contributes nothing: for the begin/end failures the reference gets
// R = 0 => L = 0 => W = 0; for the gap-threshold failure its outputs are
// unassigned (undefined behaviour) and are DEFINED here as zero contribution.
inline ViconEval eval_vicon_factor(const Interpolator &interp, const EstimateFlags &cfg, const NavState &s, const Globals &gl, bool jac) {
  ViconEval o;
  const Vec4 q_VtoB = quat_multiply(Inv(gl.q_BtoI), s.q);
  const Vec3 p_BinV = s.p + quat_2_Rot(s.q).transpose() * gl.p_BinI;
  Interpolator::Result ir = interp.get_pose_with_jacobian(s.t - gl.t_off);
  o.has_vicon = ir.ok;
  o.idx0 = ir.idx0;
  o.idx1 = ir.idx1;
  if (!ir.ok)
    return o; // zero residual, zero Jacobians
  Mat6 L;
  if (!chol_lower(ir.R, L)) {
    // Eigen's LLT on an indefinite matrix stops early; treat as zero weight
    o.has_vicon = false;
    return o;
  }
  o.W = inv_lower(L);
  const Vec4 q_r = quat_multiply(q_VtoB, Inv(ir.q));
  Vec6 err;
  for (int i = 0; i < 3; i++) {
    err(i) = 2 * q_r(i);
    err(3 + i) = p_BinV(i) - ir.p(i);
  }
  o.e = o.W * err;
  if (!jac)
    return o;
  Mat<6, 15> H;
  H.setBlock(0, 0, quat_2_Rot(Inv(gl.q_BtoI)));
  H.setBlock(3, 0, -(quat_2_Rot(Inv(s.q)) * skew_x(gl.p_BinI)));
  H.setBlock(3, 12, Mat3::Identity());
  o.H1 = o.W * H;
  if (cfg.ori) {
    Mat<6, 3> Hc;
    Hc.setBlock(0, 0, -quat_2_Rot(Inv(gl.q_BtoI)));
    o.H2 = o.W * Hc;
  }
  if (cfg.pos) {
    Mat<6, 3> Hc;
    Hc.setBlock(3, 0, quat_2_Rot(Inv(s.q)));
    o.H3 = o.W * Hc;
  }
  if (cfg.toff) {
    o.H4 = o.W * (-ir.H_toff);
  }
  return o;
}

// GTSAM mEstimator::Huber(k): weight and loss on the whitened residual norm
inline double huber_weight(double k, double dist) { return (dist <= k) ? 1.0 : k / dist; }
inline double huber_loss(double k, double dist) { return (dist <= k) ? 0.5 * dist * dist : k * (dist - 0.5 * k); }

} // namespace orc
