// vicon2gt_b200 host library: synthetic input generator (SE(3) B-spline simulator).
//
// Input generator for the benchmark / test configurations of BASELINE.json ("synthetic
// trajectories from src/sim").  CPU, untimed, never on the solve path.  Behaviour follows
// the reference's src/sim/BsplineSE3.cpp, src/sim/Simulator.cpp, src/sim/SimulatorParams.h
// and the measurement loop + state grid of src/run_simulation.cpp:108-158:
//   * control points every dt = max(mean dt, 0.05) by SE(3) interpolation of the trajectory
//     rows (last row unused), spline starts at t_min + 2 dt        (BsplineSE3.cpp:21-83)
//   * cumulative cubic B-spline pose / velocity / acceleration      (BsplineSE3.cpp:85-240)
//   * three clocks (imu, cam, vicon); a sensor fires only when it is not later than the
//     other two; vicon stamp = clock - viconimu_dt                  (Simulator.cpp:135-262)
//   * std::mt19937(seed) + a FRESH std::normal_distribution per call (libstdc++ polar
//     method) so seeds mean what they mean in a GCC build of the reference
//   * random truth per seed: R_BtoI = exp(0.1 n), p_BinI = 0.2 n, dt = 0.08 n,
//     R_GtoV = rot_y(0.1 pi n) rot_x(0.1 pi n)                      (Simulator.cpp:29-38)
// The trajectory files of the reference (data/*.txt) are not shipped; besides loading a
// "t x y z qx qy qz qw" text file this generator can synthesise a smooth trajectory of a
// requested shape (start time, rows, dt), see v2gt_sim_make_trajectory.
#include "../csrc/v2gt_math.cuh"
#include "../csrc/v2gt_spline.cuh"
#include "v2gt_sim.h"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <random>
#include <sstream>
#include <string>
#include <vector>

using v2sp::M4;
using v2sp::Spline;

struct v2gt_sim {
  v2gt_sim_params prm;
  std::vector<std::array<double, 8>> traj;
  Spline spline;
  // truth (Simulator.cpp:29-38)
  double R_BtoI[9], p_BinI[3], viconimu_dt, R_GtoV[9];
  // outputs
  std::vector<double> imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, cam_t, state_t;
  std::vector<double> bias_t, bias_g, bias_a; // true bias history (Simulator.cpp:72-77,188-190)
  std::string error;
};

extern "C" {

void v2gt_sim_params_default(v2gt_sim_params *p) {
  std::memset(p, 0, sizeof(*p));
  p->sigma_w = 1.6968e-04; // SimulatorParams.h:41-47
  p->sigma_wb = 1.9393e-05;
  p->sigma_a = 2.0000e-3;
  p->sigma_ab = 3.0000e-03;
  const double sv[6] = {1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2}; // SimulatorParams.h:53
  std::memcpy(p->sigma_vicon_pose, sv, sizeof(sv));
  p->gravity_magnitude = 9.81;
  p->sim_freq_imu = 400.0; // SimulatorParams.h:106-112
  p->sim_freq_cam = 10.0;
  p->sim_freq_vicon = 100.0;
  p->seed = 0;
  p->state_freq = 100; // run_simulation.cpp:54
}

v2gt_sim *v2gt_sim_create(const v2gt_sim_params *p) {
  v2gt_sim *s = new v2gt_sim();
  s->prm = *p;
  return s;
}
void v2gt_sim_destroy(v2gt_sim *s) { delete s; }
const char *v2gt_sim_error(v2gt_sim *s) { return s->error.c_str(); }

// Simulator::load_data (Simulator.cpp:264-319): '#' comment lines skipped, fields split on
// single spaces, first 8 numeric fields "t x y z qx qy qz qw" parsed with atof.
int v2gt_sim_load_trajectory_file(v2gt_sim *s, const char *path) {
  std::ifstream file(path);
  if (!file) {
    s->error = std::string("unable to open trajectory file ") + path;
    return 1;
  }
  s->traj.clear();
  std::string line;
  while (std::getline(file, line)) {
    if (!line.find("#"))
      continue;
    int i = 0;
    std::istringstream ss(line);
    std::string field;
    std::array<double, 8> data{};
    while (std::getline(ss, field, ' ')) {
      if (field.empty() || i >= 8)
        continue;
      data[i] = std::atof(field.c_str());
      i++;
    }
    if (i > 7)
      s->traj.push_back(data);
  }
  if (s->traj.empty()) {
    s->error = "could not parse any data from the trajectory file";
    return 1;
  }
  return 0;
}

int v2gt_sim_set_trajectory(v2gt_sim *s, int64_t n_rows, const double *rows8) {
  s->traj.resize((size_t)n_rows);
  for (int64_t i = 0; i < n_rows; i++)
    for (int d = 0; d < 8; d++)
      s->traj[(size_t)i][d] = rows8[8 * i + d];
  return n_rows >= 4 ? 0 : 1;
}

// Procedural stand-in for the reference's data/*.txt files (which are not shipped): a smooth
// hand-held-like motion of the requested shape.  rows8 layout "t x y z qx qy qz qw" (q = JPL q_GtoI).
int v2gt_sim_make_trajectory(double t_start, int64_t n_rows, double dt, double extent_m, double *rows8) {
  for (int64_t i = 0; i < n_rows; i++) {
    const double t = dt * (double)i;
    double *o = rows8 + 8 * i;
    o[0] = t_start + t;
    // position: slow corridor walk + body sway (speed ~1 m/s for extent 20 m)
    o[1] = extent_m * std::sin(2 * M_PI * t / 140.0) + 0.15 * std::sin(2 * M_PI * t / 3.1);
    o[2] = 0.25 * extent_m * std::sin(2 * M_PI * t / 70.0 + 0.7) + 0.10 * std::cos(2 * M_PI * t / 2.3);
    o[3] = 1.5 + 0.35 * std::sin(2 * M_PI * t / 11.0) + 0.05 * std::sin(2 * M_PI * t / 1.7);
    // orientation: yaw follows a slow turn, roll/pitch wobble (angular rate ~0.5-1 rad/s)
    const double yaw = 1.8 * std::sin(2 * M_PI * t / 37.0) + 0.4 * std::sin(2 * M_PI * t / 5.3);
    const double pitch = 0.25 * std::sin(2 * M_PI * t / 4.1 + 0.3);
    const double roll = 0.20 * std::sin(2 * M_PI * t / 3.3 + 1.1);
    double Rz[9], Ry[9], Rx[9], T1[9], R_ItoG[9], R_GtoI[9];
    const double cz = std::cos(yaw), sz = std::sin(yaw);
    const double Rzv[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    std::memcpy(Rz, Rzv, sizeof(Rzv));
    v2::rot_y(pitch, Ry);
    v2::rot_x(roll, Rx);
    v2::mm3(Rz, Ry, T1);
    v2::mm3(T1, Rx, R_ItoG);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++)
        R_GtoI[3 * r + c] = R_ItoG[3 * c + r];
    v2::rot_2_quat(R_GtoI, o + 4);
  }
  return 0;
}

// Simulator ctor + run_simulation.cpp:108-158 measurement loop and state grid.
int v2gt_sim_run(v2gt_sim *s) {
  const v2gt_sim_params &P = s->prm;
  if (s->traj.size() < 4) {
    s->error = "trajectory not set";
    return 1;
  }
  // ---- random truth (Simulator.cpp:29-38)
  {
    std::normal_distribution<double> w(0, 1);
    std::mt19937 r((unsigned)P.seed);
    double wv[3];
    wv[0] = 0.1 * w(r);
    wv[1] = 0.1 * w(r);
    wv[2] = 0.1 * w(r);
    v2::exp_so3(wv, s->R_BtoI);
    s->p_BinI[0] = 0.2 * w(r);
    s->p_BinI[1] = 0.2 * w(r);
    s->p_BinI[2] = 0.2 * w(r);
    s->viconimu_dt = 0.08 * w(r);
    // Simulator.cpp:38 leaves the order of the two draws unspecified; GCC evaluates the
    // operands of the overloaded operator* right to left, i.e. rot_x's angle is drawn first.
    const double ax = 0.1 * M_PI * w(r);
    const double ay = 0.1 * M_PI * w(r);
    v2::rot_xy(ax, ay, s->R_GtoV);
  }
  if (P.use_fixed_truth) {
    std::memcpy(s->R_BtoI, P.fixed_R_BtoI, sizeof(s->R_BtoI));
    std::memcpy(s->p_BinI, P.fixed_p_BinI, sizeof(s->p_BinI));
    std::memcpy(s->R_GtoV, P.fixed_R_GtoV, sizeof(s->R_GtoV));
    s->viconimu_dt = P.fixed_viconimu_dt;
  }
  s->spline = Spline();
  s->spline.feed_trajectory(s->traj);
  double timestamp_last_imu = s->spline.timestamp_start;
  double timestamp_last_cam = s->spline.timestamp_start;
  double timestamp_last_vicon = s->spline.timestamp_start;
  double Rtmp[9], ptmp[3];
  if (!s->spline.eval(s->spline.timestamp_start, 0, Rtmp, ptmp, nullptr, nullptr, nullptr, nullptr)) {
    s->error = "unable to find the first pose in the spline";
    return 1;
  }
  double true_bias_accel[3] = {0, 0, 0}, true_bias_gyro[3] = {0, 0, 0};
  s->imu_t.clear(); s->imu_w.clear(); s->imu_a.clear();
  s->vic_t.clear(); s->vic_q.clear(); s->vic_p.clear();
  s->cam_t.clear(); s->state_t.clear();
  s->bias_t.clear(); s->bias_g.clear(); s->bias_a.clear();
  auto push_bias = [&](double t) {
    s->bias_t.push_back(t);
    for (int d = 0; d < 3; d++) {
      s->bias_g.push_back(true_bias_gyro[d]);
      s->bias_a.push_back(true_bias_accel[d]);
    }
  };
  push_bias(timestamp_last_imu - 1.0 / P.sim_freq_imu);
  push_bias(timestamp_last_imu);
  std::mt19937 gen_meas_imu((unsigned)P.seed), gen_meas_vicon((unsigned)P.seed);
  bool is_running = true;
  double start_time = -1, end_time = -1;
  const int64_t cap = P.max_imu > 0 ? P.max_imu : INT64_MAX;

  while (is_running) {
    // ---- IMU (Simulator.cpp:135-194)
    if (!(timestamp_last_cam + 1.0 / P.sim_freq_cam < timestamp_last_imu + 1.0 / P.sim_freq_imu ||
          timestamp_last_vicon + 1.0 / P.sim_freq_vicon < timestamp_last_imu + 1.0 / P.sim_freq_imu)) {
      timestamp_last_imu += 1.0 / P.sim_freq_imu;
      double R_GtoI[9], p_IinG[3], w_IinI[3], v_IinG[3], alpha_IinI[3], a_IinG[3];
      if (!s->spline.eval(timestamp_last_imu, 2, R_GtoI, p_IinG, w_IinI, v_IinG, alpha_IinI, a_IinG) ||
          (int64_t)s->imu_t.size() >= cap) {
        is_running = false;
      } else {
        const double ag[3] = {a_IinG[0], a_IinG[1], a_IinG[2] + P.gravity_magnitude};
        double accel_inI[3];
        v2::mv3(R_GtoI, ag, accel_inI);
        const double dt = 1.0 / P.sim_freq_imu;
        std::normal_distribution<double> w(0, 1);
        double wm[3], am[3];
        for (int d = 0; d < 3; d++)
          wm[d] = w_IinI[d] + true_bias_gyro[d] + P.sigma_w / std::sqrt(dt) * w(gen_meas_imu);
        for (int d = 0; d < 3; d++)
          am[d] = accel_inI[d] + true_bias_accel[d] + P.sigma_a / std::sqrt(dt) * w(gen_meas_imu);
        for (int d = 0; d < 3; d++)
          true_bias_gyro[d] += P.sigma_wb * std::sqrt(dt) * w(gen_meas_imu);
        for (int d = 0; d < 3; d++)
          true_bias_accel[d] += P.sigma_ab * std::sqrt(dt) * w(gen_meas_imu);
        push_bias(timestamp_last_imu);
        s->imu_t.push_back(timestamp_last_imu);
        for (int d = 0; d < 3; d++) {
          s->imu_w.push_back(wm[d]);
          s->imu_a.push_back(am[d]);
        }
      }
    }
    // ---- camera clock (Simulator.cpp:196-210); only the clock matters
    if (!(timestamp_last_imu + 1.0 / P.sim_freq_imu < timestamp_last_cam + 1.0 / P.sim_freq_cam ||
          timestamp_last_vicon + 1.0 / P.sim_freq_vicon < timestamp_last_cam + 1.0 / P.sim_freq_cam)) {
      timestamp_last_cam += 1.0 / P.sim_freq_cam;
      s->cam_t.push_back(timestamp_last_cam);
    }
    // ---- vicon (Simulator.cpp:212-262)
    if (!(timestamp_last_imu + 1.0 / P.sim_freq_imu < timestamp_last_vicon + 1.0 / P.sim_freq_vicon ||
          timestamp_last_cam + 1.0 / P.sim_freq_cam < timestamp_last_vicon + 1.0 / P.sim_freq_vicon)) {
      timestamp_last_vicon += 1.0 / P.sim_freq_vicon;
      const double time_vicon = timestamp_last_vicon - s->viconimu_dt;
      double R_GtoI[9], p_IinG[3];
      if (!s->spline.eval(timestamp_last_vicon, 0, R_GtoI, p_IinG, nullptr, nullptr, nullptr, nullptr)) {
        is_running = false;
      } else {
        double R_GtoB[9], pb[3], p_BinG[3], R_VtoB[9], p_BinV[3];
        v2::mm3_tn(s->R_BtoI, R_GtoI, R_GtoB);
        v2::mv3_t(R_GtoI, s->p_BinI, pb);
        for (int d = 0; d < 3; d++)
          p_BinG[d] = p_IinG[d] + pb[d];
        v2::mm3_nt(R_GtoB, s->R_GtoV, R_VtoB);
        v2::mv3(s->R_GtoV, p_BinG, p_BinV);
        std::normal_distribution<double> w(0, 1);
        double w_vec[3], p_vec[3];
        for (int d = 0; d < 3; d++)
          w_vec[d] = P.sigma_vicon_pose[d] * w(gen_meas_vicon);
        for (int d = 0; d < 3; d++)
          p_vec[d] = P.sigma_vicon_pose[3 + d] * w(gen_meas_vicon);
        double E[9], Rn[9], q[4];
        v2::exp_so3(w_vec, E);
        v2::mm3(E, R_VtoB, Rn);
        v2::rot_2_quat(Rn, q);
        s->vic_t.push_back(time_vicon);
        for (int d = 0; d < 4; d++)
          s->vic_q.push_back(q[d]);
        for (int d = 0; d < 3; d++)
          s->vic_p.push_back(p_BinV[d] + p_vec[d]);
        if (start_time == -1)
          start_time = time_vicon;
        if (start_time != -1)
          end_time = time_vicon;
      }
    }
  }
  // ---- state time grid (run_simulation.cpp:149-158) or external state times (configs C2/C3)
  if (P.state_freq > 0 && start_time != -1 && end_time != -1 && start_time < end_time) {
    double temp_time = start_time;
    while (temp_time < end_time) {
      s->state_t.push_back(temp_time);
      temp_time += 1.0 / (double)P.state_freq;
    }
  } else if (P.state_freq <= 0) {
    // states at the trajectory rows' own timestamps ("states at data/<file> timestamps")
    for (const auto &row : s->traj)
      s->state_t.push_back(row[0]);
  }
  return 0;
}

int64_t v2gt_sim_n_imu(v2gt_sim *s) { return (int64_t)s->imu_t.size(); }
int64_t v2gt_sim_n_vicon(v2gt_sim *s) { return (int64_t)s->vic_t.size(); }
int64_t v2gt_sim_n_states(v2gt_sim *s) { return (int64_t)s->state_t.size(); }

int v2gt_sim_get(v2gt_sim *s, double *imu_t, double *imu_w, double *imu_a, double *vic_t, double *vic_q, double *vic_p,
                 double *state_t) {
  auto cp = [](double *dst, const std::vector<double> &v) {
    if (dst && !v.empty())
      std::memcpy(dst, v.data(), v.size() * sizeof(double));
  };
  cp(imu_t, s->imu_t);
  cp(imu_w, s->imu_w);
  cp(imu_a, s->imu_a);
  cp(vic_t, s->vic_t);
  cp(vic_q, s->vic_q);
  cp(vic_p, s->vic_p);
  cp(state_t, s->state_t);
  return 0;
}

int v2gt_sim_get_truth(v2gt_sim *s, double *R_BtoI, double *p_BinI, double *viconimu_dt, double *R_GtoV) {
  std::memcpy(R_BtoI, s->R_BtoI, sizeof(s->R_BtoI));
  std::memcpy(p_BinI, s->p_BinI, sizeof(s->p_BinI));
  *viconimu_dt = s->viconimu_dt;
  std::memcpy(R_GtoV, s->R_GtoV, sizeof(s->R_GtoV));
  return 0;
}

// Simulator::get_state_in_vicon (Simulator.cpp:91-133): out17 = [t, q_VtoI(4), p_IinV(3), v_IinV(3), bg(3), ba(3)].
// The bias bracket is found by bisection instead of the reference's linear scan.
int v2gt_sim_get_state_in_vicon(v2gt_sim *s, int64_t n, const double *times, double *out17, int32_t *ok) {
  for (int64_t k = 0; k < n; k++) {
    const double t = times[k];
    double *o = out17 + 17 * k;
    std::memset(o, 0, 17 * sizeof(double));
    o[4] = 1;
    double R_GtoI[9], p_IinG[3], w_IinI[3], v_IinG[3];
    const bool success_vel = s->spline.eval(t, 1, R_GtoI, p_IinG, w_IinI, v_IinG, nullptr, nullptr);
    // first i with bias_t[i] < t && bias_t[i+1] >= t
    bool success_bias = false;
    size_t id = 0;
    auto it = std::lower_bound(s->bias_t.begin(), s->bias_t.end(), t); // first >= t
    if (it != s->bias_t.begin() && it != s->bias_t.end()) {
      id = (size_t)(it - s->bias_t.begin()) - 1;
      success_bias = true;
    }
    if (ok)
      ok[k] = (success_vel && success_bias) ? 1 : 0;
    if (!success_vel || !success_bias)
      continue;
    const double lambda = (t - s->bias_t[id]) / (s->bias_t[id + 1] - s->bias_t[id]);
    double Rq[9];
    v2::mm3_nt(R_GtoI, s->R_GtoV, Rq);
    o[0] = t;
    v2::rot_2_quat(Rq, o + 1);
    v2::mv3(s->R_GtoV, p_IinG, o + 5);
    v2::mv3(s->R_GtoV, v_IinG, o + 8);
    for (int d = 0; d < 3; d++) {
      o[11 + d] = (1 - lambda) * s->bias_g[3 * id + d] + lambda * s->bias_g[3 * (id + 1) + d];
      o[14 + d] = (1 - lambda) * s->bias_a[3 * id + d] + lambda * s->bias_a[3 * (id + 1) + d];
    }
  }
  return 0;
}

} // extern "C"
