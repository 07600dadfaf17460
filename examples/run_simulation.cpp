// Non-ROS counterpart of the reference's run_simulation node (src/run_simulation.cpp:37-322): simulate one
// trajectory, feed the Propagator / Interpolator stores, solve with ViconGraphSolver (CUDA, behind the C ABI),
// print the trajectory errors and the converged calibration, optionally write the CSV + info files.
//
//   run_simulation [--seed N] [--duration S] [--freq_imu Hz] [--freq_vicon Hz] [--state_freq Hz]
//                  [--vicon_sigmas so so so sp sp sp] [--states out.csv --info out.txt] [--device N] [--device-sim]
// --device-sim draws the measurements and the truth with the batched device simulator (SimBatch, csrc/k_sim.cu) instead
// of the host one: the same seed gives the same run to round-off.
// Defaults are launch/sim.launch (400 Hz IMU, 100 Hz vicon, state_freq 100, sigmas 0.0174 / 0.05, seed 1) on a
// synthesised trajectory of the tum_corridor1 shape (the reference's data files are not shipped).
#include "v2gt_sim.h"
#include "vicon2gt_b200.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <string>

using namespace vicon2gt_b200;

static void quat_mul(const double *q, const double *p, double *o) { // quat_ops.h:181-196 (JPL)
  o[0] = q[3] * p[0] + q[2] * p[1] - q[1] * p[2] + q[0] * p[3];
  o[1] = -q[2] * p[0] + q[3] * p[1] + q[0] * p[2] + q[1] * p[3];
  o[2] = q[1] * p[0] - q[0] * p[1] + q[3] * p[2] + q[2] * p[3];
  o[3] = -q[0] * p[0] - q[1] * p[1] - q[2] * p[2] + q[3] * p[3];
  double n = std::sqrt(o[0] * o[0] + o[1] * o[1] + o[2] * o[2] + o[3] * o[3]);
  if (o[3] < 0)
    n = -n;
  for (int k = 0; k < 4; k++)
    o[k] /= n;
}

struct Stats { // src/utils/stats.h:64-112
  std::vector<double> v;
  double rmse = 0, mean = 0, mn = 0, mx = 0, sd = 0;
  void calculate() {
    if (v.empty())
      return;
    mn = mx = v[0];
    for (double x : v) {
      mean += x;
      rmse += x * x;
      mn = std::min(mn, x);
      mx = std::max(mx, x);
    }
    mean /= v.size();
    rmse = std::sqrt(rmse / v.size());
    for (double x : v)
      sd += (x - mean) * (x - mean);
    sd = std::sqrt(sd / std::max<size_t>(1, v.size() - 1));
  }
};

int main(int argc, char **argv) {
  int seed = 1, device = 0, state_freq = 100;
  double duration = 30.0, freq_imu = 400, freq_vicon = 100;
  double sig[6] = {0.0174, 0.0174, 0.0174, 0.05, 0.05, 0.05};
  std::string path_states, path_info;
  bool device_sim = false;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() { return (i + 1 < argc) ? argv[++i] : (char *)"0"; };
    if (a == "--seed") seed = std::atoi(next());
    else if (a == "--duration") duration = std::atof(next());
    else if (a == "--freq_imu") freq_imu = std::atof(next());
    else if (a == "--freq_vicon") freq_vicon = std::atof(next());
    else if (a == "--state_freq") state_freq = std::atoi(next());
    else if (a == "--device") device = std::atoi(next());
    else if (a == "--states") path_states = next();
    else if (a == "--info") path_info = next();
    else if (a == "--device-sim") device_sim = true;
    else if (a == "--vicon_sigmas") for (int k = 0; k < 6; k++) sig[k] = std::atof(next());
  }
  // ---- simulate (run_simulation.cpp:100-158)
  v2gt_sim_params sp;
  v2gt_sim_params_default(&sp);
  sp.seed = seed;
  sp.sim_freq_imu = freq_imu;
  sp.sim_freq_cam = 20;
  sp.sim_freq_vicon = freq_vicon;
  sp.state_freq = state_freq;
  for (int k = 0; k < 6; k++)
    sp.sigma_vicon_pose[k] = sig[k];
  v2gt_sim *sim = v2gt_sim_create(&sp);
  const int64_t rows = std::max<int64_t>(8, (int64_t)std::llround(duration / 0.05) + 1);
  std::vector<double> traj((size_t)rows * 8);
  v2gt_sim_make_trajectory(1520531829.301144123, rows, 0.05, 20.0, traj.data());
  if (v2gt_sim_set_trajectory(sim, rows, traj.data()) != 0 || v2gt_sim_run(sim) != 0) {
    std::fprintf(stderr, "simulation failed: %s\n", v2gt_sim_error(sim));
    return EXIT_FAILURE;
  }
  int64_t M = v2gt_sim_n_imu(sim), V = v2gt_sim_n_vicon(sim), N = v2gt_sim_n_states(sim);
  std::vector<double> imu_t(M), imu_w(3 * M), imu_a(3 * M), vic_t(V), vic_q(4 * V), vic_p(3 * V), state_t(N);
  v2gt_sim_get(sim, imu_t.data(), imu_w.data(), imu_a.data(), vic_t.data(), vic_q.data(), vic_p.data(), state_t.data());
  std::unique_ptr<SimBatch> dsim;
  if (device_sim) {
    try {
      dsim.reset(new SimBatch(1, traj, device));
      dsim->set_trial(0, sp);
      dsim->run();
      dsim->get(0, imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, state_t);
    } catch (const SolverError &e) {
      std::fprintf(stderr, "device simulation failed: %s (status %d)\n", e.what(), e.code());
      return EXIT_FAILURE;
    }
    M = (int64_t)imu_t.size();
    V = (int64_t)vic_t.size();
    N = (int64_t)state_t.size();
  }

  auto propagator = std::make_shared<Propagator>(sp.sigma_w, sp.sigma_wb, sp.sigma_a, sp.sigma_ab);
  auto interpolator = std::make_shared<Interpolator>();
  for (int64_t i = 0; i < M; i++)
    propagator->feed_imu(imu_t[i], {imu_w[3 * i], imu_w[3 * i + 1], imu_w[3 * i + 2]}, {imu_a[3 * i], imu_a[3 * i + 1], imu_a[3 * i + 2]});
  Mat3 R_q{}, R_p{}; // run_simulation.cpp:84-96
  for (int k = 0; k < 3; k++) {
    R_q[4 * k] = sig[k] * sig[k];
    R_p[4 * k] = sig[3 + k] * sig[3 + k];
  }
  for (int64_t i = 0; i < V; i++)
    interpolator->feed_pose(vic_t[i], {vic_q[4 * i], vic_q[4 * i + 1], vic_q[4 * i + 2], vic_q[4 * i + 3]},
                            {vic_p[3 * i], vic_p[3 * i + 1], vic_p[3 * i + 2]}, R_q, R_p);
  std::printf("[SIM]: %lld IMU, %lld vicon, %lld state times\n", (long long)M, (long long)V, (long long)N);

  // ---- solve (run_simulation.cpp:167-168)
  ViconGraphSolverParams params; // sim.launch: identity extrinsics, zero offset, estimate everything
  ViconGraphSolver solver(params, propagator, interpolator, state_t, device, /*exit_on_failure=*/true);
  const auto t0 = std::chrono::steady_clock::now();
  solver.build_and_solve();
  const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const v2gt_trial_info &inf = solver.info();
  std::printf("[TIME]: %.4f seconds total (stage + build + optimize)\n", secs);
  std::printf("[LM]: %d states, %d iterations, %d damped solves, error %.6e -> %.6e\n", inf.n_states, inf.iterations, inf.tries,
              inf.initial_error, inf.final_error);
  if (!path_states.empty() && !path_info.empty())
    solver.write_to_file(path_states, path_info);

  // ---- trajectory errors vs truth (run_simulation.cpp:198-297)
  std::vector<double> times;
  std::vector<Pose10> poses;
  solver.get_imu_poses(times, poses);
  std::vector<double> gt(17 * times.size());
  std::vector<int32_t> ok(times.size());
  if (dsim)
    dsim->get_state_in_vicon(0, times, gt, ok);
  else
    v2gt_sim_get_state_in_vicon(sim, (int64_t)times.size(), times.data(), gt.data(), ok.data());
  Stats e_ori, e_pos, e_vel;
  for (size_t i = 0; i < times.size(); i++) {
    if (!ok[i])
      continue;
    const double *g = &gt[17 * i];
    const Pose10 &e = poses[i];
    const double qinv[4] = {-e[0], -e[1], -e[2], e[3]};
    double dq[4];
    quat_mul(g + 1, qinv, dq);
    e_ori.v.push_back(180.0 / M_PI * 2.0 * std::sqrt(dq[0] * dq[0] + dq[1] * dq[1] + dq[2] * dq[2]));
    e_pos.v.push_back(std::sqrt(std::pow(e[4] - g[5], 2) + std::pow(e[5] - g[6], 2) + std::pow(e[6] - g[7], 2)));
    e_vel.v.push_back(std::sqrt(std::pow(e[7] - g[8], 2) + std::pow(e[8] - g[9], 2) + std::pow(e[9] - g[10], 2)));
  }
  e_ori.calculate();
  e_pos.calculate();
  e_vel.calculate();
  std::printf("======================================\nTrajectory Errors (deg,m,m/s)\n======================================\n");
  std::printf("rmse_ori = %.5f | rmse_pos = %.5f | rmse_vel = %.5f\n", e_ori.rmse, e_pos.rmse, e_vel.rmse);
  std::printf("mean_ori = %.5f | mean_pos = %.5f | mean_vel = %.5f\n", e_ori.mean, e_pos.mean, e_vel.mean);
  std::printf("min_ori  = %.5f | min_pos  = %.5f | min_vel  = %.5f\n", e_ori.mn, e_pos.mn, e_vel.mn);
  std::printf("max_ori  = %.5f | max_pos  = %.5f | max_vel  = %.5f\n", e_ori.mx, e_pos.mx, e_vel.mx);
  std::printf("std_ori  = %.5f | std_pos  = %.5f | std_vel  = %.5f\n\n", e_ori.sd, e_pos.sd, e_vel.sd);

  // ---- converged calibration vs truth (run_simulation.cpp:299-318)
  double toff, gt_toff;
  Mat3 R_BtoI, R_GtoV, gt_R_BtoI, gt_R_GtoV;
  Vec3 p_BinI, gt_p;
  solver.get_calibration(toff, R_BtoI, p_BinI, R_GtoV);
  if (dsim)
    dsim->get_truth(0, gt_R_BtoI, gt_p, gt_toff, gt_R_GtoV);
  else
    v2gt_sim_get_truth(sim, gt_R_BtoI.data(), gt_p.data(), &gt_toff, gt_R_GtoV.data());
  auto rpy = [](const Mat3 &R, double *o) { // rpy_ops.h:68-74
    o[0] = std::atan2(R[7], R[8]);
    o[1] = std::atan2(-R[6], std::sqrt(R[7] * R[7] + R[8] * R[8]));
    o[2] = std::atan2(R[3], R[0]);
  };
  double a[3], b[3];
  rpy(gt_R_GtoV, a);
  rpy(R_GtoV, b);
  std::printf("======================================\nConverged Calib vs Groundtruth\n======================================\n");
  std::printf("GT  toff: %.5f\nEST toff: %.5f\n\n", gt_toff, toff);
  std::printf("GT  rpy R_GtoV: %.3f, %.3f, %.3f\nEST rpy R_GtoV: %.3f, %.3f, %.3f\n\n", a[0], a[1], a[2], b[0], b[1], b[2]);
  std::printf("GT  R_BtoI row0: %.4f, %.4f, %.4f\nEST R_BtoI row0: %.4f, %.4f, %.4f\n\n", gt_R_BtoI[0], gt_R_BtoI[1], gt_R_BtoI[2],
              R_BtoI[0], R_BtoI[1], R_BtoI[2]);
  std::printf("GT  p_BinI: %.3f, %.3f, %.3f\nEST p_BinI: %.3f, %.3f, %.3f\n", gt_p[0], gt_p[1], gt_p[2], p_BinI[0], p_BinI[1], p_BinI[2]);
  v2gt_sim_destroy(sim);
  return EXIT_SUCCESS;
}
