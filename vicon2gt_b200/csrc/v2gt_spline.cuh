// vicon2gt_b200: SE(3) helpers and the cumulative cubic B-spline of the reference's simulator, on plain fp64 arrays.
//
// Shared by the host simulator (vicon2gt_b200/host/sim.cpp, g++) and by the GPU simulator (csrc/k_sim.cu): the same
// statements evaluate a pose / velocity / acceleration on either side.  Behaviour follows src/utils/quat_ops.h
// (exp_se3 :311-340, log_se3 :353-385, hat_se3 :397-402, Inv_se3 :432) and src/sim/BsplineSE3.cpp (feed_trajectory
// :21-83, get_pose / get_velocity / get_acceleration :85-240; its two lookups :242-351 are one binary search over
// sorted times here, spline_bracket below).  The host side is pinned on the reference's own BsplineSE3 / Simulator
// (tests/test_ref_pin_cpu.py::test_host_simulator_against_the_reference_simulator).
#pragma once
#include "v2gt_math.cuh"

namespace v2sp {

struct M4 {
  double a[16];
};
V2_HD M4 zero4() {
  M4 m;
  for (int i = 0; i < 16; i++)
    m.a[i] = 0.0;
  return m;
}
V2_HD M4 eye4() {
  M4 m = zero4();
  m.a[0] = m.a[5] = m.a[10] = m.a[15] = 1;
  return m;
}
V2_HD M4 mul(const M4 &A, const M4 &B) {
  M4 C;
  for (int r = 0; r < 4; r++)
    for (int c = 0; c < 4; c++) {
      double s = 0;
      for (int k = 0; k < 4; k++)
        s += A.a[4 * r + k] * B.a[4 * k + c];
      C.a[4 * r + c] = s;
    }
  return C;
}
V2_HD M4 add(const M4 &A, const M4 &B) {
  M4 C;
  for (int i = 0; i < 16; i++)
    C.a[i] = A.a[i] + B.a[i];
  return C;
}
V2_HD M4 scale(double s, const M4 &A) {
  M4 C;
  for (int i = 0; i < 16; i++)
    C.a[i] = s * A.a[i];
  return C;
}
V2_HD void getR(const M4 &T, double *R) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      R[3 * r + c] = T.a[4 * r + c];
}
V2_HD void setR(M4 &T, const double *R) {
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      T.a[4 * r + c] = R[3 * r + c];
}
// quat_ops.h:432 Inv_se3
V2_HD M4 inv_se3(const M4 &T) {
  M4 Ti = eye4();
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      Ti.a[4 * r + c] = T.a[4 * c + r];
  for (int r = 0; r < 3; r++) {
    double s = 0;
    for (int c = 0; c < 3; c++)
      s += Ti.a[4 * r + c] * T.a[4 * c + 3];
    Ti.a[4 * r + 3] = -s;
  }
  return Ti;
}
// quat_ops.h:311-340 exp_se3
V2_HD M4 exp_se3(const double *vec) {
  const double *w = vec, *u = vec + 3;
  const double theta = sqrt(v2::dot3(w, w));
  double S[9], S2[9];
  v2::skew(w, S);
  v2::mm3(S, S, S2);
  double A, B, C;
  if (theta < 1e-12) {
    A = 1;
    B = 0.5;
    C = 1.0 / 6.0;
  } else {
    A = sin(theta) / theta;
    B = (1 - cos(theta)) / (theta * theta);
    C = (1 - A) / (theta * theta);
  }
  M4 m = zero4();
  double V[9], R[9];
  for (int i = 0; i < 9; i++) {
    const double I = (i % 4 == 0) ? 1.0 : 0.0;
    V[i] = I + B * S[i] + C * S2[i];
    R[i] = I + A * S[i] + B * S2[i];
  }
  setR(m, R);
  double t[3];
  v2::mv3(V, u, t);
  m.a[3] = t[0];
  m.a[7] = t[1];
  m.a[11] = t[2];
  m.a[15] = 1;
  return m;
}
// quat_ops.h:353-385 log_se3
V2_HD void log_se3(const M4 &mat, double *vec) {
  double R[9], t[3] = {mat.a[3], mat.a[7], mat.a[11]};
  getR(mat, R);
  const double a = 0.5 * ((R[0] + R[4] + R[8]) - 1);
  const double theta = (a > 1) ? acos(1.0) : ((a < -1) ? acos(-1.0) : acos(a));
  double A, B, D, E;
  if (theta < 1e-12) {
    A = 1;
    B = 0.5;
    D = 0.5;
    E = 1.0 / 12.0;
  } else {
    A = sin(theta) / theta;
    B = (1 - cos(theta)) / (theta * theta);
    D = theta / (2 * sin(theta));
    E = 1 / (theta * theta) * (1 - 0.5 * A / B);
  }
  double W[9], W2[9], Vinv[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      W[3 * r + c] = D * (R[3 * r + c] - R[3 * c + r]);
  v2::mm3(W, W, W2);
  for (int i = 0; i < 9; i++)
    Vinv[i] = ((i % 4 == 0) ? 1.0 : 0.0) - 0.5 * W[i] + E * W2[i];
  vec[0] = W[7];
  vec[1] = W[2];
  vec[2] = W[3];
  v2::mv3(Vinv, t, vec + 3);
}
// quat_ops.h:397-402 hat_se3
V2_HD M4 hat_se3(const double *vec) {
  M4 m = zero4();
  double S[9];
  v2::skew(vec, S);
  setR(m, S);
  m.a[3] = vec[3];
  m.a[7] = vec[4];
  m.a[11] = vec[5];
  return m;
}

// Bracket of `timestamp` in the sorted, unique control-point times `t[0..n)`, as the std::map lookups of
// find_bounding_poses + find_bounding_control_points give it: i1 = the last time <= timestamp, i2 = the first time
// > timestamp, and both i1 - 1 and i2 + 1 must exist.  Returns false where the reference's lookup fails.
V2_HD bool spline_bracket(const double *t, long long n, double timestamp, long long &i1) {
  long long lo = 0, hi = n; // first index with t[idx] > timestamp (upper_bound)
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (t[mid] > timestamp)
      hi = mid;
    else
      lo = mid + 1;
  }
  const long long ub = lo;
  if (ub >= n)
    return false; // no newer control point
  // older: lower_bound == the key itself, or the one before the first key greater than the timestamp
  if (ub == 0)
    return false; // timestamp before the first control point
  i1 = ub - 1;    // t[i1] <= timestamp < t[ub]
  if (i1 == 0)
    return false; // no control point before the bracket
  if (ub + 1 >= n)
    return false; // no control point after the bracket
  return true;
}

// BsplineSE3.cpp:85-240 with the four bounding control points given: order 0 = pose, 1 = +velocity, 2 = +acceleration
V2_HD void spline_eval_core(const M4 &pose0, const M4 &pose1, const M4 &pose2, const M4 &pose3, double t1, double t2, double timestamp,
                            int order, double *R_GtoI, double *p_IinG, double *w_IinI, double *v_IinG, double *alpha_IinI,
                            double *a_IinG) {
  const double DT = (t2 - t1);
  const double u = (timestamp - t1) / DT;
  const double b0 = 1.0 / 6.0 * (5 + 3 * u - 3 * u * u + u * u * u);
  const double b1 = 1.0 / 6.0 * (1 + 3 * u + 3 * u * u - 2 * u * u * u);
  const double b2 = 1.0 / 6.0 * (u * u * u);
  const double b0dot = 1.0 / (6.0 * DT) * (3 - 6 * u + 3 * u * u);
  const double b1dot = 1.0 / (6.0 * DT) * (3 + 6 * u - 6 * u * u);
  const double b2dot = 1.0 / (6.0 * DT) * (3 * u * u);
  const double b0dd = 1.0 / (6.0 * DT * DT) * (-6 + 6 * u);
  const double b1dd = 1.0 / (6.0 * DT * DT) * (6 - 12 * u);
  const double b2dd = 1.0 / (6.0 * DT * DT) * (6 * u);
  double o10[6], o21[6], o32[6], s[6];
  log_se3(mul(inv_se3(pose0), pose1), o10);
  log_se3(mul(inv_se3(pose1), pose2), o21);
  log_se3(mul(inv_se3(pose2), pose3), o32);
  for (int i = 0; i < 6; i++)
    s[i] = b0 * o10[i];
  const M4 A0 = exp_se3(s);
  for (int i = 0; i < 6; i++)
    s[i] = b1 * o21[i];
  const M4 A1 = exp_se3(s);
  for (int i = 0; i < 6; i++)
    s[i] = b2 * o32[i];
  const M4 A2 = exp_se3(s);
  const M4 pose_interp = mul(mul(mul(pose0, A0), A1), A2);
  double Rt[9];
  getR(pose_interp, Rt); // R_ItoG
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++)
      R_GtoI[3 * r + c] = Rt[3 * c + r];
  p_IinG[0] = pose_interp.a[3];
  p_IinG[1] = pose_interp.a[7];
  p_IinG[2] = pose_interp.a[11];
  if (order < 1)
    return;
  const M4 h10 = hat_se3(o10), h21 = hat_se3(o21), h32 = hat_se3(o32);
  const M4 A0dot = scale(b0dot, mul(h10, A0));
  const M4 A1dot = scale(b1dot, mul(h21, A1));
  const M4 A2dot = scale(b2dot, mul(h32, A2));
  const M4 vel_interp = mul(pose0, add(add(mul(mul(A0dot, A1), A2), mul(mul(A0, A1dot), A2)), mul(mul(A0, A1), A2dot)));
  double Rv[9], om[9];
  getR(vel_interp, Rv);
  v2::mm3_tn(Rt, Rv, om); // R^T Rdot = skew(omega)
  w_IinI[0] = om[7];
  w_IinI[1] = om[2];
  w_IinI[2] = om[3];
  v_IinG[0] = vel_interp.a[3];
  v_IinG[1] = vel_interp.a[7];
  v_IinG[2] = vel_interp.a[11];
  if (order < 2)
    return;
  const M4 A0dd = add(scale(b0dot, mul(h10, A0dot)), scale(b0dd, mul(h10, A0)));
  const M4 A1dd = add(scale(b1dot, mul(h21, A1dot)), scale(b1dd, mul(h21, A1)));
  const M4 A2dd = add(scale(b2dot, mul(h32, A2dot)), scale(b2dd, mul(h32, A2)));
  M4 acc = mul(mul(A0dd, A1), A2);
  acc = add(acc, mul(mul(A0, A1dd), A2));
  acc = add(acc, mul(mul(A0, A1), A2dd));
  acc = add(acc, scale(2, mul(mul(A0dot, A1dot), A2)));
  acc = add(acc, scale(2, mul(mul(A0, A1dot), A2dot)));
  acc = add(acc, scale(2, mul(mul(A0dot, A1), A2dot)));
  const M4 acc_interp = mul(pose0, acc);
  double Ra[9], tmp[9], al[9];
  getR(acc_interp, Ra);
  v2::mm3(Rv, om, tmp);
  for (int i = 0; i < 9; i++)
    tmp[i] = Ra[i] - tmp[i];
  v2::mm3_tn(Rt, tmp, al);
  alpha_IinI[0] = al[7];
  alpha_IinI[1] = al[2];
  alpha_IinI[2] = al[3];
  a_IinG[0] = acc_interp.a[3];
  a_IinG[1] = acc_interp.a[7];
  a_IinG[2] = acc_interp.a[11];
}

} // namespace v2sp

// ---- host side: control points from trajectory rows.  The reference keeps both the trajectory poses and the control
// points in ordered associative containers keyed by time (BsplineSE3.cpp:21-83, lookups :242-351); here both are plain
// arrays sorted by time with repeated stamps dropped (the first one fed wins), the uniform walk over the trajectory
// keeps a forward cursor, and a stamp is bracketed by spline_bracket() above -- the same routine the device simulator
// uses.  Plain host code, also parsed (and dropped) by nvcc's device pass of csrc/k_sim.cu.
#if 1
#include <algorithm>
#include <array>
#include <cmath>
#include <numeric>
#include <vector>
namespace v2sp {

struct Spline {
  std::vector<double> cp_t; // control-point times: strictly increasing, uniform spacing dt
  std::vector<M4> cp;       // control-point poses
  double dt = 0, timestamp_start = 0;

  void feed_trajectory(const std::vector<std::array<double, 8>> &traj) {
    cp_t.clear();
    cp.clear();
    const size_t rows = traj.size();
    double span = 0;
    for (size_t i = 0; i + 1 < rows; i++)
      span += traj[i + 1][0] - traj[i][0];
    dt = std::max(0.05, span / (double)(rows - 1));
    // trajectory poses ordered by time; the last row is not used (the reference stops one short)
    const size_t used = rows ? rows - 1 : 0;
    std::vector<size_t> order(used);
    std::iota(order.begin(), order.end(), (size_t)0);
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return traj[a][0] < traj[b][0]; });
    std::vector<double> pt;
    std::vector<M4> pp;
    pt.reserve(used);
    pp.reserve(used);
    for (size_t k : order) {
      if (!pt.empty() && pt.back() == traj[k][0])
        continue;
      double R[9], Rt[9];
      v2::quat_2_Rot(&traj[k][4], R);
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++)
          Rt[3 * r + c] = R[3 * c + r];
      M4 T = eye4();
      setR(T, Rt);
      T.a[3] = traj[k][1];
      T.a[7] = traj[k][2];
      T.a[11] = traj[k][3];
      pt.push_back(traj[k][0]);
      pp.push_back(T);
    }
    if (pt.empty())
      return;
    // uniform walk from the first stamp: every control point is the SE(3) interpolation of the two trajectory poses
    // around it; the walk ends at the first stamp with no later pose.  `nxt` = first pose later than the stamp.
    size_t nxt = 0;
    for (double ts = pt.front();; ts += dt) {
      while (nxt < pt.size() && !(pt[nxt] > ts))
        nxt++;
      if (nxt == pt.size() || nxt == 0)
        break;
      const size_t prv = nxt - 1;
      const double lambda = (ts - pt[prv]) / (pt[nxt] - pt[prv]);
      double lg[6];
      log_se3(mul(pp[nxt], inv_se3(pp[prv])), lg);
      for (int i = 0; i < 6; i++)
        lg[i] *= lambda;
      cp_t.push_back(ts);
      cp.push_back(mul(exp_se3(lg), pp[prv]));
    }
    timestamp_start = pt.front() + 2 * dt;
  }

  // BsplineSE3.cpp:85-240 : order 0 = pose, 1 = +velocity, 2 = +acceleration
  bool eval(double timestamp, int order, double *R_GtoI, double *p_IinG, double *w_IinI, double *v_IinG, double *alpha_IinI,
            double *a_IinG) const {
    long long i1 = 0;
    if (!spline_bracket(cp_t.data(), (long long)cp_t.size(), timestamp, i1))
      return false;
    const size_t k = (size_t)i1;
    spline_eval_core(cp[k - 1], cp[k], cp[k + 1], cp[k + 2], cp_t[k], cp_t[k + 1], timestamp, order, R_GtoI, p_IinG, w_IinI, v_IinG,
                     alpha_IinI, a_IinG);
    return true;
  }
};

} // namespace v2sp
#endif
