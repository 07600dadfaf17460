// Non-ROS counterpart of the reference's Monte-Carlo sweep (scripts/run_monte_carlo.sh: one `roslaunch vicon2gt sim.launch`
// per seed and noise level) on the C++ host classes of include/vicon2gt_b200.hpp: every run is simulated on the GPU
// (SimBatch), handed to the solver device to device (BatchSolver::set_trial_sim), solved in ONE batch, and reported with
// the per-run error statistics of run_simulation.cpp:198-297 computed on the device.
//
//   run_monte_carlo [--seeds N] [--duration S] [--ori_noise rad] [--pos_noise m] [--device N] [--out-dir DIR] [--wave W]
// --wave W: the runs are fed from HOST stores instead and solved W at a time by WaveRunner (a sweep larger than the GPU's
// memory: staging of the next wave and read-back of the previous one overlap the solve).
// Prints one line per run: seed, states, LM iterations, final cost, mean ATE (deg / m), time-offset error (ms).
#include "v2gt_sim.h"
#include "vicon2gt_b200.hpp"

#include <chrono>
#include <cmath>
#include <cstring>
#include <string>

using namespace vicon2gt_b200;

int main(int argc, char **argv) {
  int seeds = 8, device = 0, wave = 0;
  double duration = 30.0, so = 0.0174, sp = 0.05;
  std::string out_dir;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() { return (i + 1 < argc) ? argv[++i] : (char *)"0"; };
    if (a == "--seeds") seeds = std::atoi(next());
    else if (a == "--duration") duration = std::atof(next());
    else if (a == "--ori_noise") so = std::atof(next());
    else if (a == "--pos_noise") sp = std::atof(next());
    else if (a == "--device") device = std::atoi(next());
    else if (a == "--out-dir") out_dir = next();
    else if (a == "--wave") wave = std::atoi(next());
  }
  if (seeds < 1)
    return EXIT_FAILURE;
  const int64_t rows = std::max<int64_t>(8, (int64_t)std::llround(duration / 0.05) + 1);
  std::vector<double> traj((size_t)rows * 8);
  v2gt_sim_make_trajectory(1520531829.301144123, rows, 0.05, 20.0, traj.data());
  try {
    const auto t0 = std::chrono::steady_clock::now();
    // ---- simulate every run on the device (sim.launch rates; seed = run index)
    SimBatch sim(seeds, traj, device);
    for (int k = 0; k < seeds; k++) {
      v2gt_sim_params p;
      v2gt_sim_params_default(&p);
      p.seed = k;
      p.sim_freq_imu = 400;
      p.sim_freq_cam = 20;
      p.sim_freq_vicon = 100;
      p.state_freq = 100;
      for (int d = 0; d < 3; d++) {
        p.sigma_vicon_pose[d] = so;
        p.sigma_vicon_pose[3 + d] = sp;
      }
      sim.set_trial(k, p);
    }
    sim.run();
    const auto t1 = std::chrono::steady_clock::now();
    if (wave > 0) {
      // ---- host stores per run (feed_imu / feed_pose loop of run_simulation.cpp:118-145), solved in waves
      std::vector<Propagator> imus;
      std::vector<Interpolator> vics((size_t)seeds);
      std::vector<std::vector<double>> grids((size_t)seeds);
      Mat3 R_q{}, R_p{};
      for (int d = 0; d < 3; d++) {
        R_q[4 * d] = so * so;
        R_p[4 * d] = sp * sp;
      }
      v2gt_sim_params dp;
      v2gt_sim_params_default(&dp);
      for (int k = 0; k < seeds; k++) {
        std::vector<double> it, iw, ia, vt, vq, vp;
        sim.get(k, it, iw, ia, vt, vq, vp, grids[(size_t)k]);
        imus.emplace_back(dp.sigma_w, dp.sigma_wb, dp.sigma_a, dp.sigma_ab);
        for (size_t i = 0; i < it.size(); i++)
          imus.back().feed_imu(it[i], {iw[3 * i], iw[3 * i + 1], iw[3 * i + 2]}, {ia[3 * i], ia[3 * i + 1], ia[3 * i + 2]});
        for (size_t i = 0; i < vt.size(); i++)
          vics[(size_t)k].feed_pose(vt[i], {vq[4 * i], vq[4 * i + 1], vq[4 * i + 2], vq[4 * i + 3]}, {vp[3 * i], vp[3 * i + 1], vp[3 * i + 2]},
                                    R_q, R_p);
      }
      ViconGraphSolverParams params;
      std::vector<WaveRunner::Trial> trials;
      for (int k = 0; k < seeds; k++)
        trials.push_back({&params, &imus[(size_t)k], &vics[(size_t)k], &grids[(size_t)k]});
      std::vector<v2gt_trial_info> infos((size_t)seeds);
      WaveRunner runner(wave, device);
      const auto w0 = std::chrono::steady_clock::now();
      runner.run(trials, [&](int first, BatchSolver &h) {
        for (int i = 0; i < h.size(); i++)
          infos[(size_t)(first + i)] = h.info(i);
      });
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count();
      std::printf("[MC]: %d runs in waves of %d, %.3f s wall, %.3f s on the device\n", seeds, wave, secs, runner.device_ms() * 1e-3);
      for (int k = 0; k < seeds; k++)
        std::printf("run %3d: status %d, %d states, %d iterations, cost %.6e\n", k, infos[(size_t)k].status, infos[(size_t)k].n_states,
                    infos[(size_t)k].iterations, infos[(size_t)k].final_error);
      return EXIT_SUCCESS;
    }
    // ---- solve them as one batch
    std::array<double, 18> R{}; // run_simulation.cpp:84-96
    for (int d = 0; d < 3; d++) {
      R[4 * d] = so * so;
      R[9 + 4 * d] = sp * sp;
    }
    ViconGraphSolverParams params;
    BatchSolver solver(seeds, device);
    for (int k = 0; k < seeds; k++)
      solver.set_trial_sim(k, params, sim, k, R);
    solver.build_and_solve();
    const auto t2 = std::chrono::steady_clock::now();
    // ---- results + truth at the surviving state times -> error report on the device
    std::vector<int64_t> off;
    std::vector<double> times, states, globals;
    solver.get_all_states(off, times, states, globals);
    for (int k = 0; k < seeds; k++) {
      std::vector<double> tk(times.begin() + off[(size_t)k], times.begin() + off[(size_t)k + 1]), gt, truth10;
      std::vector<int32_t> ok;
      sim.get_state_in_vicon(k, tk, gt, ok);
      truth10.resize(tk.size() * 10);
      for (size_t i = 0; i < tk.size(); i++)
        std::memcpy(&truth10[10 * i], &gt[17 * i + 1], 10 * sizeof(double));
      solver.set_truth(k, truth10);
    }
    const std::vector<double> st = solver.error_stats();
    const auto t3 = std::chrono::steady_clock::now();
    auto secs = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double>(b - a).count();
    };
    std::printf("[MC]: %d runs, simulate %.3f s, build_and_solve %.3f s, read back + error report %.3f s\n", seeds, secs(t0, t1), secs(t1, t2),
                secs(t2, t3));
    for (int k = 0; k < seeds; k++) {
      const v2gt_trial_info inf = solver.info(k);
      Mat3 Rb, Rg;
      Vec3 pb;
      double gt_toff = 0;
      sim.get_truth(k, Rb, pb, gt_toff, Rg);
      std::printf("run %3d: status %d, %d states, %d iterations, cost %.6e, mean ATE %.5f deg / %.5f m, toff error %.3f ms\n", k, inf.status,
                  inf.n_states, inf.iterations, inf.final_error, st[(size_t)k * 24 + 1], st[(size_t)k * 24 + 8 + 1],
                  1e3 * std::fabs(globals[(size_t)k * 10 + 7] - gt_toff));
    }
    if (!out_dir.empty())
      std::printf("[MC]: wrote %d runs under %s\n", solver.write_all(out_dir, true), out_dir.c_str());
  } catch (const SolverError &e) {
    std::fprintf(stderr, "[vicon2gt_b200] %s (status %d)%s\n", e.what(), e.code(), e.code() == V2GT_ERR_NO_DEVICE ? ": no CUDA device" : "");
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
