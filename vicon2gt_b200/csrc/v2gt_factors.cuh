// vicon2gt_b200 kernel (3), device functions: residuals and Jacobian pieces of the two factors.
//
//   ImuFactorCPIv1::evaluateError                      (src/gtsam/ImuFactorCPIv1.cpp:25-216)
//   MeasBased_ViconPoseTimeoffsetFactor::evaluateError (src/gtsam/MeasBased_ViconPoseTimeoffsetFactor.cpp:24-142)
// plus the GTSAM noise models they are wrapped in (Gaussian::Covariance -> information matrix;
// Robust(Huber(k), Unit) -> block reweighting by sqrt(w) and the Huber loss).
#pragma once
#include "v2gt_interp.cuh"

namespace v2 {

// gravity in the vicon frame and d(rot e_z)/d(theta_x, theta_y)   (ImuFactorCPIv1.cpp:44-45, 191-197)
struct GravTerms {
  double g[3];      // rot_y(thy) rot_x(thx) [0 0 g_mag]^T
  double Hxy[6];    // 3x2 row-major
  double gmag;
};
__device__ __forceinline__ void grav_terms(double thx, double thy, double gmag, GravTerms &o) {
  double R[9];
  rot_xy(thx, thy, R);
  o.g[0] = R[2] * gmag; o.g[1] = R[5] * gmag; o.g[2] = R[8] * gmag;
  double sx, cx, sy, cy;
  sincos(thx, &sx, &cx);
  sincos(thy, &sy, &cy);
  o.Hxy[0] = -sy * sx; o.Hxy[1] = cy * cx;
  o.Hxy[2] = -cx;      o.Hxy[3] = 0.0;
  o.Hxy[4] = -cy * sx; o.Hxy[5] = -sy * cx;
  o.gmag = gmag;
}

// load the 3x3 block (a,b) of a packed-lower symmetric 15x15 matrix
template <class View> __device__ __forceinline__ void ld_sym_blk(const View P, int a, int b, double *o) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const int R = 3 * a + r, C = 3 * b + c;
      o[3 * r + c] = P[R >= C ? tri(R, C) : tri(C, R)];
    }
}

// e^T Lambda e with Lambda packed lower
__device__ __forceinline__ double quad_form15(const SoaR P, const double *e) {
  double s = 0;
  int k = 0;
#pragma unroll
  for (int r = 0; r < 15; r++) {
    double row = 0;
#pragma unroll
    for (int c = 0; c < r; c++)
      row += P[k + c] * e[c];
    s += e[r] * (2.0 * row + P[k + r] * e[r]);
    k += r + 1;
  }
  return s;
}

struct ImuPieces {
  double e[15];
  double Tth[9], Tb[9], sv[3], sp[3], Rk[9], Jth2[9];
};

// residual (always) and Jacobian pieces (JAC) of the IMU factor between xi and xj
template <bool JAC>
__device__ inline void imu_factor(const SoaR pre, const double *xi, const double *xj, const GravTerms &G,
                                  ImuPieces &o) {
  const double dt = pre[PRE_DT];
  double dbg[3], dba[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    dbg[k] = xi[4 + k] - pre[PRE_BGLIN + k];
    dba[k] = xi[10 + k] - pre[PRE_BALIN + k];
  }
  double Jq[9], qmeas[4];
#pragma unroll
  for (int e = 0; e < 9; e++)
    Jq[e] = pre[PRE_JQ + e];
#pragma unroll
  for (int e = 0; e < 4; e++)
    qmeas[e] = pre[PRE_Q + e];
  // q_b = rot_2_quat(exp_so3(-J_q (bg_k - bg_lin)))
  double th[3], Eb[9], q_b[4];
  mv3(Jq, dbg, th);
  th[0] = -th[0]; th[1] = -th[1]; th[2] = -th[2];
  exp_so3(th, Eb);
  rot_2_quat(Eb, q_b);
  double qi_inv[4], qmeas_inv[4], qb_inv[4], q_n[4], q_rminus[4], q_r[4], q_m[4];
  quat_inv(xi, qi_inv);
  quat_inv(qmeas, qmeas_inv);
  quat_inv(q_b, qb_inv);
  quat_mul(xj, qi_inv, q_n);
  quat_mul(q_n, qmeas_inv, q_rminus);
  quat_mul(q_rminus, q_b, q_r);
  quat_mul(qb_inv, qmeas, q_m);
  double Rk[9];
  quat_2_Rot(xi, Rk);
  const double dt2 = dt * dt;
  double pterm[3], vterm[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    pterm[k] = xj[13 + k] - xi[13 + k] - xi[7 + k] * dt + 0.5 * G.g[k] * dt2;
    vterm[k] = xj[7 + k] - xi[7 + k] + G.g[k] * dt;
  }
  double sp[3], sv[3];
  mv3(Rk, pterm, sp);
  mv3(Rk, vterm, sv);
  double Ja[9], Jb[9], Ha[9], Hb[9];
#pragma unroll
  for (int e = 0; e < 9; e++) {
    Jb[e] = pre[PRE_JB + e];
    Ja[e] = pre[PRE_JA + e];
    Hb[e] = pre[PRE_HB + e];
    Ha[e] = pre[PRE_HA + e];
  }
  double t1[3], t2[3], t3[3], t4[3];
  mv3(Ja, dbg, t1);
  mv3(Ha, dba, t2);
  mv3(Jb, dbg, t3);
  mv3(Hb, dba, t4);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.e[k] = 2 * q_r[k];
    o.e[3 + k] = xj[4 + k] - xi[4 + k];
    o.e[6 + k] = ((sv[k] - t3[k]) - t4[k]) - pre[PRE_BETA + k];
    o.e[9 + k] = xj[10 + k] - xi[10 + k];
    o.e[12 + k] = ((sp[k] - t1[k]) - t2[k]) - pre[PRE_ALPHA + k];
  }
  if (JAC) {
    // Tth = -((qn4 I - [qn]x)(qm4 I - [qm]x) + qn qm^T)
    double Sn[9], Sm[9], An[9], Am[9], Pm[9];
    skew(q_n, Sn);
    skew(q_m, Sm);
#pragma unroll
    for (int e = 0; e < 9; e++) {
      An[e] = ((e % 4 == 0) ? q_n[3] : 0.0) - Sn[e];
      Am[e] = ((e % 4 == 0) ? q_m[3] : 0.0) - Sm[e];
    }
    mm3(An, Am, Pm);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        o.Tth[3 * r + c] = -(Pm[3 * r + c] + q_n[r] * q_m[c]);
    double Sr[9], Ar[9];
    skew(q_rminus, Sr);
#pragma unroll
    for (int e = 0; e < 9; e++)
      Ar[e] = ((e % 4 == 0) ? q_rminus[3] : 0.0) - Sr[e];
    mm3(Ar, Jq, o.Tb);
    double Sq[9];
    skew(q_r, Sq);
#pragma unroll
    for (int e = 0; e < 9; e++) {
      o.Jth2[e] = ((e % 4 == 0) ? q_r[3] : 0.0) + Sq[e];
      o.Rk[e] = Rk[e];
    }
#pragma unroll
    for (int k = 0; k < 3; k++) {
      o.sv[k] = sv[k];
      o.sp[k] = sp[k];
    }
  }
}

struct ViconPieces {
  double r[6];      // sqrt(w) W e
  double A[36];     // sqrt(w) W H1(:, [theta p])  6x6 row-major
  double G[42];     // sqrt(w) W [H2 H3 H4]        6x7 row-major
  double w;         // Huber weight
  double dist;      // |W e|
  int has, idx0, idx1;
};

// vicon factor of state x (time t) against the trial's pose store
template <bool JAC>
__device__ inline void vicon_factor(const TrialDev &tr, const double *__restrict__ vt, const double *__restrict__ vs,
                                    const double *__restrict__ vc, bool cov_const, double t, const double *x, const double *gl,
                                    ViconPieces &o, int *bracket_hint = nullptr) {
  InterpResult ir;
  // the state's bracket at its previous evaluation is the guess for this one (lower bound = idx0 + 1)
  vicon_interp(vt, vs, vc, cov_const, tr.n_vic, t - gl[GL_TOFF], tr.prm.interp_gap_factor, true, ir,
               bracket_hint ? (long long)*bracket_hint : -1);
  if (bracket_hint && ir.idx0 >= 0)
    *bracket_hint = ir.idx0 + 1;
  o.has = ir.ok;
  o.idx0 = ir.idx0;
  o.idx1 = ir.idx1;
  o.w = 1.0;
  o.dist = 0.0;
#pragma unroll
  for (int k = 0; k < 6; k++)
    o.r[k] = 0.0;
  if (JAC) {
#pragma unroll
    for (int k = 0; k < 36; k++)
      o.A[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 42; k++)
      o.G[k] = 0.0;
  }
  if (!ir.ok)
    return;
  // L = chol(R) (6x6 lower), W = L^-1
  double R36[36], L[36], W[36];
  interp_dense_cov(ir, R36);
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 36; k++) {
    L[k] = 0.0;
    W[k] = 0.0;
  }
#pragma unroll
  for (int j = 0; j < 6; j++) {
    double d = R36[6 * j + j];
#pragma unroll
    for (int k = 0; k < j; k++)
      d -= L[6 * j + k] * L[6 * j + k];
    if (!(d > 0.0))
      bad = true;
    const double ljj = sqrt(d);
    L[6 * j + j] = ljj;
#pragma unroll
    for (int i = j + 1; i < 6; i++) {
      double s = R36[6 * i + j];
#pragma unroll
      for (int k = 0; k < j; k++)
        s -= L[6 * i + k] * L[6 * j + k];
      L[6 * i + j] = s / ljj;
    }
  }
  if (bad) {
    o.has = 0;
    return;
  }
#pragma unroll
  for (int c = 0; c < 6; c++)
#pragma unroll
    for (int r = c; r < 6; r++) {
      double s = (r == c) ? 1.0 : 0.0;
#pragma unroll
      for (int k = c; k < r; k++)
        s -= L[6 * r + k] * W[6 * k + c];
      W[6 * r + c] = s / L[6 * r + r];
    }
  // expected measurement
  double qBI_inv[4], q_VtoB[4], Rq[9], pB[3], p_BinV[3];
  quat_inv(gl + GL_Q, qBI_inv);
  quat_mul(qBI_inv, x, q_VtoB);
  quat_2_Rot(x, Rq);
  mv3_t(Rq, gl + GL_P, pB);
#pragma unroll
  for (int k = 0; k < 3; k++)
    p_BinV[k] = x[13 + k] + pB[k];
  double qi_inv[4], q_r[4], err[6];
  quat_inv(ir.q, qi_inv);
  quat_mul(q_VtoB, qi_inv, q_r);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    err[k] = 2 * q_r[k];
    err[3 + k] = p_BinV[k] - ir.p[k];
  }
  double we[6];
  double n2 = 0;
#pragma unroll
  for (int r = 0; r < 6; r++) {
    double s = 0;
#pragma unroll
    for (int c = 0; c <= r; c++)
      s += W[6 * r + c] * err[c];
    we[r] = s;
    n2 += s * s;
  }
  const double dist = sqrt(n2);
  o.dist = dist;
  const double k_h = tr.prm.huber_k;
  o.w = (dist <= k_h) ? 1.0 : k_h / dist;
  const double sw = sqrt(o.w);
#pragma unroll
  for (int r = 0; r < 6; r++)
    o.r[r] = sw * we[r];
  if (!JAC)
    return;
  // raw Jacobians:  H1 = [Ra 0 ; Sb I] on columns [theta | p], Ra = R(q_BtoI^-1), Sb = -R(q_VtoI^-1) [p_BinI]x
  double Ra[9], RqT[9], Sx[9], Sb[9];
  quat_2_Rot(qBI_inv, Ra);
  quat_inv(x, qi_inv);
  quat_2_Rot(qi_inv, RqT);
  skew(gl + GL_P, Sx);
  mm3(RqT, Sx, Sb);
  double H1[36];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) {
      H1[6 * r + c] = Ra[3 * r + c];
      H1[6 * r + 3 + c] = 0.0;
      H1[6 * (3 + r) + c] = -Sb[3 * r + c];
      H1[6 * (3 + r) + 3 + c] = (r == c) ? 1.0 : 0.0;
    }
  double Hg[42];
#pragma unroll
  for (int k = 0; k < 42; k++)
    Hg[k] = 0.0;
  if (tr.prm.estimate_ori_vicon_to_imu) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        Hg[7 * r + c] = -Ra[3 * r + c];
  }
  if (tr.prm.estimate_pos_vicon_to_imu) {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        Hg[7 * (3 + r) + 3 + c] = RqT[3 * r + c];
  }
  if (tr.prm.estimate_toff_vicon_to_imu) {
#pragma unroll
    for (int r = 0; r < 6; r++)
      Hg[7 * r + 6] = -ir.Ht[r];
  }
#pragma unroll
  for (int r = 0; r < 6; r++) {
#pragma unroll
    for (int c = 0; c < 6; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k <= r; k++)
        s += W[6 * r + k] * H1[6 * k + c];
      o.A[6 * r + c] = sw * s;
    }
#pragma unroll
    for (int c = 0; c < 7; c++) {
      double s = 0;
#pragma unroll
      for (int k = 0; k <= r; k++)
        s += W[6 * r + k] * Hg[7 * k + c];
      o.G[7 * r + c] = sw * s;
    }
  }
}

__device__ __forceinline__ double huber_loss(double k, double dist) {
  return (dist <= k) ? 0.5 * dist * dist : k * (dist - 0.5 * k);
}

} // namespace v2
