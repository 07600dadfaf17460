// ORACLE (test infrastructure, NOT product code) -- pinned on the reference's own sources (oracle/_ref, tests/test_ref_pin_cpu.py).
//
// Continuous preintegration, model 1 (piecewise-constant measurement), restated
// from the reference's src/cpi/CpiBase.h:49-116 and src/cpi/CpiV1.h:58-341.
// Error-state / covariance ordering: [theta(0:3) bg(3:6) v(6:9) ba(9:12) p(12:15)].
#pragma once
#include "quat_ops.hpp"

namespace orc {

struct CpiV1 {
  // src/cpi/CpiBase.h:49-63 : continuous noise matrix, averaging flag
  CpiV1(double sigma_w, double sigma_wb, double sigma_a, double sigma_ab, bool imu_avg_ = false) {
    for (int i = 0; i < 3; i++) {
      Q_c(i, i) = sigma_w * sigma_w;
      Q_c(3 + i, 3 + i) = sigma_wb * sigma_wb;
      Q_c(6 + i, 6 + i) = sigma_a * sigma_a;
      Q_c(9 + i, 9 + i) = sigma_ab * sigma_ab;
    }
    imu_avg = imu_avg_;
    R_k2tau = Mat3::Identity();
  }
  // src/cpi/CpiBase.h:68-76
  void setLinearizationPoints(const Vec3 &b_w_lin_, const Vec3 &b_a_lin_) {
    b_w_lin = b_w_lin_;
    b_a_lin = b_a_lin_;
  }

  bool imu_avg = false;
  double DT = 0;
  Vec3 alpha_tau, beta_tau;
  Vec4 q_k2tau;
  Mat3 R_k2tau;
  Mat3 J_q, J_a, J_b, H_a, H_b;
  Vec3 b_w_lin, b_a_lin;
  Mat<12, 12> Q_c;
  Mat15 P_meas;

  // src/cpi/CpiV1.h:58-341
  void feed_IMU(double t_0, double t_1, const Vec3 &w_m_0, const Vec3 &a_m_0, const Vec3 &w_m_1, const Vec3 &a_m_1) {
    const Mat3 eye3 = Mat3::Identity();
    const double delta_t = t_1 - t_0;
    DT += delta_t;
    if (delta_t == 0)
      return;

    Vec3 w_hat = w_m_0 - b_w_lin;
    Vec3 a_hat = a_m_0 - b_a_lin;
    if (imu_avg) {
      w_hat += w_m_1 - b_w_lin;
      w_hat = 0.5 * w_hat;
      a_hat += a_m_1 - b_a_lin;
      a_hat = .5 * a_hat;
    }
    const Vec3 w_hatdt = w_hat * delta_t;
    const double w_1 = w_hat(0), w_2 = w_hat(1), w_3 = w_hat(2);
    const double mag_w = w_hat.norm();
    const double w_dt = mag_w * delta_t;
    const bool small_w = (mag_w < 0.008726646);
    const double dt_2 = delta_t * delta_t;
    const double cos_wt = std::cos(w_dt);
    const double sin_wt = std::sin(w_dt);

    const Mat3 w_x = skew_x(w_hat);
    const Mat3 a_x = skew_x(a_hat);
    const Mat3 w_tx = skew_x(w_hatdt);
    const Mat3 w_x_2 = w_x * w_x;

    // ---- measurement means (CpiV1.h:108-148)
    const Mat3 R_tau2tau1 = small_w ? eye3 - delta_t * w_x + (std::pow(delta_t, 2) / 2) * w_x_2
                                    : eye3 - (sin_wt / mag_w) * w_x + ((1.0 - cos_wt) / (std::pow(mag_w, 2.0))) * w_x_2;
    const Mat3 R_k2tau1 = R_tau2tau1 * R_k2tau;
    const Mat3 R_tau12k = R_k2tau1.transpose();

    double f_1, f_2, f_3, f_4;
    if (small_w) {
      f_1 = -(std::pow(delta_t, 3) / 3);
      f_2 = (std::pow(delta_t, 4) / 8);
      f_3 = -(std::pow(delta_t, 2) / 2);
      f_4 = (std::pow(delta_t, 3) / 6);
    } else {
      f_1 = (w_dt * cos_wt - sin_wt) / (std::pow(mag_w, 3));
      f_2 = (std::pow(w_dt, 2) - 2 * cos_wt - 2 * w_dt * sin_wt + 2) / (2 * std::pow(mag_w, 4));
      f_3 = -(1 - cos_wt) / std::pow(mag_w, 2);
      f_4 = (w_dt - sin_wt) / std::pow(mag_w, 3);
    }
    const Mat3 alpha_arg = ((dt_2 / 2.0) * eye3 + f_1 * w_x + f_2 * w_x_2);
    const Mat3 Beta_arg = (delta_t * eye3 + f_3 * w_x + f_4 * w_x_2);
    const Mat3 H_al = R_tau12k * alpha_arg;
    const Mat3 H_be = R_tau12k * Beta_arg;
    alpha_tau += beta_tau * delta_t + H_al * a_hat;
    beta_tau += H_be * a_hat;

    // ---- bias Jacobians (CpiV1.h:150-244)
    const Mat3 J_r_tau1 =
        small_w ? eye3 - .5 * w_tx + (1.0 / 6.0) * (w_tx * w_tx)
                : eye3 - ((1 - cos_wt) / (std::pow((w_dt), 2.0))) * w_tx + ((w_dt - sin_wt) / (std::pow(w_dt, 3.0))) * (w_tx * w_tx);
    J_q = R_tau2tau1 * J_q + J_r_tau1 * delta_t;
    H_a -= H_al;
    H_a += delta_t * H_b;
    H_b -= H_be;

    const Vec3 e[3] = {Vec3{1, 0, 0}, Vec3{0, 1, 0}, Vec3{0, 0, 1}};
    Mat3 ex[3], d_R_bw[3];
    for (int j = 0; j < 3; j++) {
      ex[j] = skew_x(e[j]);
      d_R_bw[j] = -(R_tau12k * skew_x(J_q * e[j])); // uses the already-updated J_q (CpiV1.h:168-170)
    }
    double df_1_dw_mag, df_2_dw_mag, df_3_dw_mag, df_4_dw_mag;
    if (small_w) {
      df_1_dw_mag = -(std::pow(delta_t, 5) / 15);
      df_2_dw_mag = (std::pow(delta_t, 6) / 72);
      df_3_dw_mag = -(std::pow(delta_t, 4) / 12);
      df_4_dw_mag = (std::pow(delta_t, 5) / 60);
    } else {
      df_1_dw_mag = (std::pow(w_dt, 2) * sin_wt - 3 * sin_wt + 3 * w_dt * cos_wt) / std::pow(mag_w, 5);
      df_2_dw_mag = (std::pow(w_dt, 2) - 4 * cos_wt - 4 * w_dt * sin_wt + std::pow(w_dt, 2) * cos_wt + 4) / (std::pow(mag_w, 6));
      df_3_dw_mag = (2 * (cos_wt - 1) + w_dt * sin_wt) / (std::pow(mag_w, 4));
      df_4_dw_mag = (2 * w_dt + w_dt * cos_wt - 3 * sin_wt) / (std::pow(mag_w, 5));
    }
    const double wv[3] = {w_1, w_2, w_3};
    J_a += J_b * delta_t; // old J_b (CpiV1.h:232)
    for (int j = 0; j < 3; j++) {
      const double df1 = wv[j] * df_1_dw_mag, df2 = wv[j] * df_2_dw_mag;
      const double df3 = wv[j] * df_3_dw_mag, df4 = wv[j] * df_4_dw_mag;
      const Mat3 sym = ex[j] * w_x + w_x * ex[j];
      const Vec3 ca = (d_R_bw[j] * alpha_arg + R_tau12k * (df1 * w_x - f_1 * ex[j] + df2 * w_x_2 - f_2 * sym)) * a_hat;
      const Vec3 cb = (d_R_bw[j] * Beta_arg + R_tau12k * (df3 * w_x - f_3 * ex[j] + df4 * w_x_2 - f_4 * sym)) * a_hat;
      for (int r = 0; r < 3; r++) {
        J_a(r, j) += ca(r);
        J_b(r, j) += cb(r);
      }
    }

    // ---- measurement covariance, RK4 (CpiV1.h:246-335)
    Mat3 R_mid = small_w ? eye3 - .5 * delta_t * w_x + (std::pow(.5 * delta_t, 2) / 2) * w_x_2
                         : eye3 - (std::sin(mag_w * .5 * delta_t) / mag_w) * w_x +
                               ((1.0 - std::cos(mag_w * .5 * delta_t)) / (std::pow(mag_w, 2.0))) * w_x_2;
    R_mid = R_mid * R_k2tau;

    auto P_dot = [&](const Mat3 &Rk, const Mat15 &P) {
      Mat15 F;
      F.setBlock(0, 0, -w_x);
      F.setBlock(0, 3, -eye3);
      F.setBlock(6, 0, -(Rk.transpose() * a_x));
      F.setBlock(6, 9, -Rk.transpose());
      F.setBlock(12, 6, eye3);
      Mat<15, 12> G;
      G.setBlock(0, 0, -eye3);
      G.setBlock(3, 3, eye3);
      G.setBlock(6, 6, -Rk.transpose());
      G.setBlock(9, 9, eye3);
      return F * P + P * F.transpose() + G * Q_c * G.transpose();
    };
    const Mat15 P_dot_k1 = P_dot(R_k2tau, P_meas);
    const Mat15 P_k2 = P_meas + P_dot_k1 * delta_t / 2.0;
    const Mat15 P_dot_k2 = P_dot(R_mid, P_k2);
    const Mat15 P_k3 = P_meas + P_dot_k2 * delta_t / 2.0;
    const Mat15 P_dot_k3 = P_dot(R_mid, P_k3);
    const Mat15 P_k4 = P_meas + P_dot_k3 * delta_t;
    const Mat15 P_dot_k4 = P_dot(R_k2tau1, P_k4);
    P_meas += (delta_t / 6.0) * (P_dot_k1 + 2.0 * P_dot_k2 + 2.0 * P_dot_k3 + P_dot_k4);
    P_meas = 0.5 * (P_meas + P_meas.transpose());

    // rotation mean last (old orientation is used by the covariance above)
    R_k2tau = R_k2tau1;
    q_k2tau = rot_2_quat(R_k2tau);
  }
};

} // namespace orc
