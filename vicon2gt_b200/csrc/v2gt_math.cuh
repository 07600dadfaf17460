// vicon2gt_b200: JPL quaternion / SO(3) / roll-pitch primitives on plain fp64 arrays.
//
// Shared by every CUDA kernel (device) and by the host-side simulator (host), written for
// registers: fixed-size row-major 3x3 blocks, no heap, no dynamic indexing in the hot
// functions.  Semantics follow the reference's src/utils/quat_ops.h and
// src/utils/rpy_ops.h (citations per function) including its branch thresholds and
// sign conventions; quaternions are JPL [x y z w], R = quat_2_Rot(q) maps global->local.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define V2_HD __host__ __device__ __forceinline__
#else
#define V2_HD inline
#endif

namespace v2 {

// ---------------------------------------------------------------- 3-vectors / 3x3 blocks
V2_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
V2_HD double norm3(const double *a) { return sqrt(dot3(a, a)); }

// S = [w]x  (quat_ops.h:136)
V2_HD void skew(const double *w, double *S) {
  S[0] = 0.0;   S[1] = -w[2]; S[2] = w[1];
  S[3] = w[2];  S[4] = 0.0;   S[5] = -w[0];
  S[6] = -w[1]; S[7] = w[0];  S[8] = 0.0;
}
// C = A B
V2_HD void mm3(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
// C = A B^T
V2_HD void mm3_nt(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] = A[3 * r] * B[3 * c] + A[3 * r + 1] * B[3 * c + 1] + A[3 * r + 2] * B[3 * c + 2];
}
// C = A^T B
V2_HD void mm3_tn(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] = A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
// y = A x ;  y = A^T x
V2_HD void mv3(const double *A, const double *x, double *y) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    y[r] = A[3 * r] * x[0] + A[3 * r + 1] * x[1] + A[3 * r + 2] * x[2];
}
V2_HD void mv3_t(const double *A, const double *x, double *y) {
#pragma unroll
  for (int r = 0; r < 3; r++)
    y[r] = A[r] * x[0] + A[3 + r] * x[1] + A[6 + r] * x[2];
}
V2_HD void eye3(double *A) {
  A[0] = 1; A[1] = 0; A[2] = 0; A[3] = 0; A[4] = 1; A[5] = 0; A[6] = 0; A[7] = 0; A[8] = 1;
}
// S2 = [w]x [w]x = w w^T - |w|^2 I, formed as the explicit product (as the reference does)
V2_HD void skew_sq(const double *w, double *S2) {
  double S[9];
  skew(w, S);
  mm3(S, S, S2);
}

// ---------------------------------------------------------------- quaternions
// quat_ops.h:85-120 : four-branch largest pivot, force w >= 0, normalise
V2_HD void rot_2_quat(const double *R, double *q) {
  const double T = R[0] + R[4] + R[8];
  if ((R[0] >= T) && (R[0] >= R[4]) && (R[0] >= R[8])) {
    q[0] = sqrt((1 + (2 * R[0]) - T) / 4);
    const double s = 1 / (4 * q[0]);
    q[1] = s * (R[1] + R[3]);
    q[2] = s * (R[2] + R[6]);
    q[3] = s * (R[5] - R[7]);
  } else if ((R[4] >= T) && (R[4] >= R[0]) && (R[4] >= R[8])) {
    q[1] = sqrt((1 + (2 * R[4]) - T) / 4);
    const double s = 1 / (4 * q[1]);
    q[0] = s * (R[1] + R[3]);
    q[2] = s * (R[5] + R[7]);
    q[3] = s * (R[6] - R[2]);
  } else if ((R[8] >= T) && (R[8] >= R[0]) && (R[8] >= R[4])) {
    q[2] = sqrt((1 + (2 * R[8]) - T) / 4);
    const double s = 1 / (4 * q[2]);
    q[0] = s * (R[2] + R[6]);
    q[1] = s * (R[5] + R[7]);
    q[3] = s * (R[1] - R[3]);
  } else {
    q[3] = sqrt((1 + T) / 4);
    const double s = 1 / (4 * q[3]);
    q[0] = s * (R[5] - R[7]);
    q[1] = s * (R[6] - R[2]);
    q[2] = s * (R[1] - R[3]);
  }
  if (q[3] < 0) {
    q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3];
  }
  const double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  q[0] /= n; q[1] /= n; q[2] /= n; q[3] /= n;
}

// quat_ops.h:153-158 : R = (2w^2-1) I - 2w [v]x + 2 v v^T
V2_HD void quat_2_Rot(const double *q, double *R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  const double d = 2 * w * w - 1, tw = 2 * w;
  R[0] = d + 2 * (x * x);
  R[1] = tw * z + 2 * (x * y);
  R[2] = -(tw * y) + 2 * (x * z);
  R[3] = -(tw * z) + 2 * (y * x);
  R[4] = d + 2 * (y * y);
  R[5] = tw * x + 2 * (y * z);
  R[6] = tw * y + 2 * (z * x);
  R[7] = -(tw * x) + 2 * (z * y);
  R[8] = d + 2 * (z * z);
}

// quat_ops.h:181-196 : o = q (x) p = L(q) p ; force w >= 0 ; normalise.  o may alias neither input.
V2_HD void quat_mul(const double *q, const double *p, double *o) {
  // L(q) = [ q4 I - [qv]x , qv ; -qv^T , q4 ]
  o[0] = q[3] * p[0] + q[2] * p[1] - q[1] * p[2] + q[0] * p[3];
  o[1] = -q[2] * p[0] + q[3] * p[1] + q[0] * p[2] + q[1] * p[3];
  o[2] = q[1] * p[0] - q[0] * p[1] + q[3] * p[2] + q[2] * p[3];
  o[3] = -q[0] * p[0] - q[1] * p[1] - q[2] * p[2] + q[3] * p[3];
  if (o[3] < 0) {
    o[0] = -o[0]; o[1] = -o[1]; o[2] = -o[2]; o[3] = -o[3];
  }
  const double n = sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
  o[0] /= n; o[1] /= n; o[2] /= n; o[3] /= n;
}
// quat_ops.h:445
V2_HD void quat_inv(const double *q, double *o) {
  o[0] = -q[0]; o[1] = -q[1]; o[2] = -q[2]; o[3] = q[3];
}

// Correction quaternion of JPLNavState::retract / JPLQuaternion::retract
// (src/gtsam/JPLNavState.cpp:28-45, JPLQuaternion.cpp:27-44): NaN (zero step) => identity.
V2_HD void small_quat(const double *dth, double *dq) {
  const double n = norm3(dth);
  const double s = sin(n / 2) / n;
  dq[0] = s * dth[0]; dq[1] = s * dth[1]; dq[2] = s * dth[2];
  dq[3] = cos(n / 2);
  const double m = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
  dq[0] /= m; dq[1] /= m; dq[2] /= m; dq[3] /= m;
  if (dq[3] < 0) {
    dq[0] = -dq[0]; dq[1] = -dq[1]; dq[2] = -dq[2]; dq[3] = -dq[3];
  }
  const double m2 = sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2] + dq[3] * dq[3]);
  if (m2 != m2) {
    dq[0] = 0; dq[1] = 0; dq[2] = 0; dq[3] = 1.0;
  }
}

// ---------------------------------------------------------------- SO(3)
// quat_ops.h:232-253
V2_HD void exp_so3(const double *w, double *R) {
  const double theta = norm3(w);
  if (theta == 0) {
    eye3(R);
    return;
  }
  double A, B;
  if (theta < 1e-12) {
    A = 1;
    B = 0.5;
  } else {
    A = sin(theta) / theta;
    B = (1 - cos(theta)) / (theta * theta);
  }
  double S[9], S2[9];
  skew(w, S);
  mm3(S, S, S2);
#pragma unroll
  for (int i = 0; i < 9; i++)
    R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + A * S[i] + B * S2[i];
}

// quat_ops.h:268-289 (acos form; exact-identity test kept)
V2_HD void log_so3(const double *R, double *w) {
  const double a = 0.5 * ((R[0] + R[4] + R[8]) - 1);
  const double theta = (a > 1) ? 0.0 : ((a < -1) ? acos(-1.0) : acos(a));
  const double D = (theta < 1e-12) ? 0.5 : theta / (2 * sin(theta));
  const bool is_eye = (R[0] == 1.0 && R[1] == 0.0 && R[2] == 0.0 && R[3] == 0.0 && R[4] == 1.0 && R[5] == 0.0 && R[6] == 0.0 &&
                       R[7] == 0.0 && R[8] == 1.0);
  if (is_eye) {
    w[0] = 0; w[1] = 0; w[2] = 0;
    return;
  }
  w[0] = D * (R[7] - R[5]);
  w[1] = D * (R[2] - R[6]);
  w[2] = D * (R[3] - R[1]);
}

// quat_ops.h:492-502 ; Jr_so3(w) = Jl_so3(-w) (quat_ops.h:514)
V2_HD void Jl_so3(const double *w, double *J) {
  const double theta = norm3(w);
  if (theta < 1e-12) {
    eye3(J);
    return;
  }
  const double a[3] = {w[0] / theta, w[1] / theta, w[2] / theta};
  const double s = sin(theta) / theta;
  const double c = (1 - cos(theta)) / theta;
  double S[9];
  skew(a, S);
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int cc = 0; cc < 3; cc++)
      J[3 * r + cc] = ((r == cc) ? s : 0.0) + (1 - s) * (a[r] * a[cc]) + c * S[3 * r + cc];
}
V2_HD void Jr_so3(const double *w, double *J) {
  const double m[3] = {-w[0], -w[1], -w[2]};
  Jl_so3(m, J);
}

// ---------------------------------------------------------------- roll / pitch (rpy_ops.h)
V2_HD void rot_x(double t, double *R) { // rpy_ops.h:28-34
  const double ct = cos(t), st = sin(t);
  R[0] = 1; R[1] = 0; R[2] = 0; R[3] = 0; R[4] = ct; R[5] = -st; R[6] = 0; R[7] = st; R[8] = ct;
}
V2_HD void rot_y(double t, double *R) { // rpy_ops.h:36-42
  const double ct = cos(t), st = sin(t);
  R[0] = ct; R[1] = 0; R[2] = st; R[3] = 0; R[4] = 1; R[5] = 0; R[6] = -st; R[7] = 0; R[8] = ct;
}
// RotationXY::rot() = rot_y(thetay) rot_x(thetax)  (src/gtsam/RotationXY.h:62)
V2_HD void rot_xy(double thx, double thy, double *R) {
  double Rx[9], Ry[9];
  rot_x(thx, Rx);
  rot_y(thy, Ry);
  mm3(Ry, Rx, R);
}
V2_HD void rot2rpy(const double *R, double *rpy) { // rpy_ops.h:68-74
  rpy[2] = atan2(R[3], R[0]);
  rpy[1] = atan2(-R[6], sqrt(R[7] * R[7] + R[8] * R[8]));
  rpy[0] = atan2(R[7], R[8]);
}
V2_HD double wrap2pi(double theta) { // rpy_ops.h:81-87
  const double PI = 3.14159265358979323846;
  while (theta > PI)
    theta -= 2 * PI;
  while (theta < -PI)
    theta += 2 * PI;
  return theta;
}

} // namespace v2
