// ORACLE (test infrastructure, NOT product code) -- pinned on the reference's own sources (oracle/_ref, tests/test_ref_pin_cpu.py).
//
// JPL quaternion / SO(3) / roll-pitch helpers, restated from the reference's
// src/utils/quat_ops.h and src/utils/rpy_ops.h (citations per function).
// Quaternions are JPL, vector first [x y z w]; R = quat_2_Rot(q) maps global->local.
#pragma once
#include "linalg.hpp"

namespace orc {

// src/utils/quat_ops.h:136
inline Mat3 skew_x(const Vec3 &w) {
  Mat3 m;
  m(0, 1) = -w(2);
  m(0, 2) = w(1);
  m(1, 0) = w(2);
  m(1, 2) = -w(0);
  m(2, 0) = -w(1);
  m(2, 1) = w(0);
  return m;
}

// src/utils/quat_ops.h:85-120 : four-branch largest pivot, force w>=0, normalise
inline Vec4 rot_2_quat(const Mat3 &rot) {
  Vec4 q;
  const double T = rot.trace();
  if ((rot(0, 0) >= T) && (rot(0, 0) >= rot(1, 1)) && (rot(0, 0) >= rot(2, 2))) {
    q(0) = std::sqrt((1 + (2 * rot(0, 0)) - T) / 4);
    q(1) = (1 / (4 * q(0))) * (rot(0, 1) + rot(1, 0));
    q(2) = (1 / (4 * q(0))) * (rot(0, 2) + rot(2, 0));
    q(3) = (1 / (4 * q(0))) * (rot(1, 2) - rot(2, 1));
  } else if ((rot(1, 1) >= T) && (rot(1, 1) >= rot(0, 0)) && (rot(1, 1) >= rot(2, 2))) {
    q(1) = std::sqrt((1 + (2 * rot(1, 1)) - T) / 4);
    q(0) = (1 / (4 * q(1))) * (rot(0, 1) + rot(1, 0));
    q(2) = (1 / (4 * q(1))) * (rot(1, 2) + rot(2, 1));
    q(3) = (1 / (4 * q(1))) * (rot(2, 0) - rot(0, 2));
  } else if ((rot(2, 2) >= T) && (rot(2, 2) >= rot(0, 0)) && (rot(2, 2) >= rot(1, 1))) {
    q(2) = std::sqrt((1 + (2 * rot(2, 2)) - T) / 4);
    q(0) = (1 / (4 * q(2))) * (rot(0, 2) + rot(2, 0));
    q(1) = (1 / (4 * q(2))) * (rot(1, 2) + rot(2, 1));
    q(3) = (1 / (4 * q(2))) * (rot(0, 1) - rot(1, 0));
  } else {
    q(3) = std::sqrt((1 + T) / 4);
    q(0) = (1 / (4 * q(3))) * (rot(1, 2) - rot(2, 1));
    q(1) = (1 / (4 * q(3))) * (rot(2, 0) - rot(0, 2));
    q(2) = (1 / (4 * q(3))) * (rot(0, 1) - rot(1, 0));
  }
  if (q(3) < 0)
    q = -q;
  return q / q.norm();
}

// src/utils/quat_ops.h:153-158 : (2w^2-1) I - 2w [v]x + 2 v v^T
inline Mat3 quat_2_Rot(const Vec4 &q) {
  Vec3 v{q(0), q(1), q(2)};
  Mat3 qx = skew_x(v);
  Mat3 R = (2 * q(3) * q(3) - 1) * Mat3::Identity() - (2 * q(3)) * qx + 2 * (v * v.transpose());
  return R;
}

// src/utils/quat_ops.h:181-196 : q (x) p = L(q) p, force w>=0, normalise
inline Vec4 quat_multiply(const Vec4 &q, const Vec4 &p) {
  Vec3 qv{q(0), q(1), q(2)};
  Mat3 blk = q(3) * Mat3::Identity() - skew_x(qv);
  Vec4 r;
  for (int i = 0; i < 3; i++)
    r(i) = blk(i, 0) * p(0) + blk(i, 1) * p(1) + blk(i, 2) * p(2) + q(i) * p(3);
  r(3) = -q(0) * p(0) - q(1) * p(1) - q(2) * p(2) + q(3) * p(3);
  if (r(3) < 0)
    r *= -1.0;
  return r / r.norm();
}

// src/utils/quat_ops.h:445
inline Vec4 Inv(const Vec4 &q) { return Vec4{-q(0), -q(1), -q(2), q(3)}; }

// src/utils/quat_ops.h:232-253
inline Mat3 exp_so3(const Vec3 &w) {
  Mat3 wx = skew_x(w);
  const double theta = w.norm();
  double A, B;
  if (theta < 1e-12) {
    A = 1;
    B = 0.5;
  } else {
    A = std::sin(theta) / theta;
    B = (1 - std::cos(theta)) / (theta * theta);
  }
  if (theta == 0)
    return Mat3::Identity();
  return Mat3::Identity() + A * wx + B * (wx * wx);
}

// src/utils/quat_ops.h:268-289 (acos form, exact-identity test kept)
inline Vec3 log_so3(const Mat3 &R) {
  const double a = 0.5 * (R.trace() - 1);
  const double theta = (a > 1) ? std::acos(1.0) : ((a < -1) ? std::acos(-1.0) : std::acos(a));
  double D;
  if (theta < 1e-12)
    D = 0.5;
  else
    D = theta / (2 * std::sin(theta));
  Mat3 wx = D * (R - R.transpose());
  if (R != Mat3::Identity())
    return Vec3{wx(2, 1), wx(0, 2), wx(1, 0)};
  return Vec3();
}

// src/utils/quat_ops.h:492-502
inline Mat3 Jl_so3(const Vec3 &w) {
  const double theta = w.norm();
  if (theta < 1e-12)
    return Mat3::Identity();
  Vec3 a = w / theta;
  return (std::sin(theta) / theta) * Mat3::Identity() + (1 - std::sin(theta) / theta) * (a * a.transpose()) +
         ((1 - std::cos(theta)) / theta) * skew_x(a);
}
// src/utils/quat_ops.h:514
inline Mat3 Jr_so3(const Vec3 &w) { return Jl_so3(-w); }

// src/utils/rpy_ops.h:28-46
inline Mat3 rot_x(double t) {
  const double ct = std::cos(t), st = std::sin(t);
  return Mat3{1.0, 0.0, 0.0, 0.0, ct, -st, 0.0, st, ct};
}
inline Mat3 rot_y(double t) {
  const double ct = std::cos(t), st = std::sin(t);
  return Mat3{ct, 0.0, st, 0.0, 1.0, 0.0, -st, 0.0, ct};
}
inline Mat3 rot_z(double t) {
  const double ct = std::cos(t), st = std::sin(t);
  return Mat3{ct, -st, 0.0, st, ct, 0.0, 0.0, 0.0, 1.0};
}
// src/utils/rpy_ops.h:68-74
inline Vec3 rot2rpy(const Mat3 &rot) {
  Vec3 rpy;
  rpy(2) = std::atan2(rot(1, 0), rot(0, 0));
  rpy(1) = std::atan2(-rot(2, 0), std::sqrt(rot(2, 1) * rot(2, 1) + rot(2, 2) * rot(2, 2)));
  rpy(0) = std::atan2(rot(2, 1), rot(2, 2));
  return rpy;
}
// src/utils/rpy_ops.h:81-87
inline double wrap2pi(double theta) {
  while (theta > M_PI)
    theta -= 2 * M_PI;
  while (theta < -M_PI)
    theta += 2 * M_PI;
  return theta;
}

} // namespace orc
