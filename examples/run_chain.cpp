// One long sequence chain-sharded over the GPUs of a node (BASELINE config 5), host side in C++ over the C ABI:
// one process per GPU.  Every rank simulates the same sequence (deterministic generator), builds its piece of the chain
// (ChainSolver), rank 0 writes the NCCL id to a file the other ranks poll, and all ranks run the Levenberg-Marquardt
// loop inside the library (ncclAllGather on the solver stream, one damped solve = one CUDA graph launch).
//
//   run_chain --rank R --world N --id-file PATH [--duration S] [--seed K] [--segs-per-rank P] [--device D] [--no-graph]
//             [--states out.csv]
// e.g.   for r in 0 1; do examples/run_chain --rank $r --world 2 --id-file /tmp/v2gt.id --device $r & done; wait
// world 1 needs neither NCCL nor the id file.  --states (rank 0, world 1 only) writes "t, q4 bg3 v3 ba3 p3" rows.
#include "v2gt_sim.h"
#include "vicon2gt_b200.hpp"

#include <chrono>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>

using namespace vicon2gt_b200;

int main(int argc, char **argv) {
  int rank = 0, world = 1, device = -1, seed = 5, spr = 0;
  double duration = 600.0;
  bool graph = true;
  std::string id_file, path_states;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto next = [&]() { return (i + 1 < argc) ? argv[++i] : (char *)"0"; };
    if (a == "--rank") rank = std::atoi(next());
    else if (a == "--world") world = std::atoi(next());
    else if (a == "--device") device = std::atoi(next());
    else if (a == "--seed") seed = std::atoi(next());
    else if (a == "--duration") duration = std::atof(next());
    else if (a == "--segs-per-rank") spr = std::atoi(next());
    else if (a == "--id-file") id_file = next();
    else if (a == "--states") path_states = next();
    else if (a == "--no-graph") graph = false;
  }
  if (device < 0)
    device = rank;
  if (world > 1 && id_file.empty()) {
    std::fprintf(stderr, "run_chain: --id-file is needed for more than one rank\n");
    return EXIT_FAILURE;
  }
  // ---- the sequence: 20 Hz states at the trajectory stamps, 200 Hz IMU, 100 Hz vicon (identical on every rank)
  v2gt_sim_params sp;
  v2gt_sim_params_default(&sp);
  sp.seed = seed;
  sp.sim_freq_imu = 200;
  sp.sim_freq_cam = 20;
  sp.sim_freq_vicon = 100;
  sp.state_freq = 0;
  for (int k = 0; k < 6; k++) // launch/sim.launch:16
    sp.sigma_vicon_pose[k] = k < 3 ? 0.0174 : 0.05;
  v2gt_sim *sim = v2gt_sim_create(&sp);
  const int64_t rows = std::max<int64_t>(8, (int64_t)std::llround(duration / 0.05) + 1);
  std::vector<double> traj((size_t)rows * 8);
  v2gt_sim_make_trajectory(1520531829.301144123, rows, 0.05, 20.0, traj.data());
  if (v2gt_sim_set_trajectory(sim, rows, traj.data()) != 0 || v2gt_sim_run(sim) != 0) {
    std::fprintf(stderr, "simulation failed: %s\n", v2gt_sim_error(sim));
    return EXIT_FAILURE;
  }
  const int64_t M = v2gt_sim_n_imu(sim), V = v2gt_sim_n_vicon(sim), N = v2gt_sim_n_states(sim);
  std::vector<double> imu_t(M), imu_w(3 * M), imu_a(3 * M), vic_t(V), vic_q(4 * V), vic_p(3 * V), state_t(N);
  v2gt_sim_get(sim, imu_t.data(), imu_w.data(), imu_a.data(), vic_t.data(), vic_q.data(), vic_p.data(), state_t.data());
  v2gt_sim_destroy(sim);
  Propagator imu(sp.sigma_w, sp.sigma_wb, sp.sigma_a, sp.sigma_ab);
  Interpolator vicon;
  for (int64_t i = 0; i < M; i++)
    imu.feed_imu(imu_t[i], {imu_w[3 * i], imu_w[3 * i + 1], imu_w[3 * i + 2]}, {imu_a[3 * i], imu_a[3 * i + 1], imu_a[3 * i + 2]});
  Mat3 R_q{}, R_p{}; // run_simulation.cpp:84-96
  for (int k = 0; k < 3; k++) {
    R_q[4 * k] = sp.sigma_vicon_pose[k] * sp.sigma_vicon_pose[k];
    R_p[4 * k] = sp.sigma_vicon_pose[3 + k] * sp.sigma_vicon_pose[3 + k];
  }
  for (int64_t i = 0; i < V; i++)
    vicon.feed_pose(vic_t[i], {vic_q[4 * i], vic_q[4 * i + 1], vic_q[4 * i + 2], vic_q[4 * i + 3]},
                    {vic_p[3 * i], vic_p[3 * i + 1], vic_p[3 * i + 2]}, R_q, R_p);
  try {
    ViconGraphSolverParams params;
    if (spr <= 0)
      spr = std::max(16, (int)std::lround(4.0 * std::sqrt((double)N) / world));
    ChainSolver cs(params, imu, vicon, state_t, rank, world, spr, device);
    if (world > 1) {
      ChainSolver::NcclId id{};
      if (rank == 0) {
        id = ChainSolver::unique_id();
        std::ofstream f(id_file + ".tmp", std::ios::binary);
        f.write(reinterpret_cast<const char *>(id.data()), (std::streamsize)id.size());
        f.close();
        std::rename((id_file + ".tmp").c_str(), id_file.c_str());
      } else {
        for (int tries = 0;; tries++) {
          std::ifstream f(id_file, std::ios::binary);
          if (f && f.read(reinterpret_cast<char *>(id.data()), (std::streamsize)id.size()))
            break;
          if (tries > 600) {
            std::fprintf(stderr, "run_chain: rank %d gave up waiting for %s\n", rank, id_file.c_str());
            return EXIT_FAILURE;
          }
          std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
      }
      cs.comm_init(id);
    }
    cs.optimize(0, graph); // warm (captures the graph)
    cs.build();
    const auto t0 = std::chrono::steady_clock::now();
    const v2gt_trial_info inf = cs.optimize(0, graph);
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const auto ms = cs.timing_ms();
    Vec4 q;
    Vec3 p;
    double toff;
    std::array<double, 2> th;
    cs.get_calibration(q, p, toff, th);
    const v2gt_chain_plan_t &pl = cs.plan();
    std::printf("[CHAIN rank %d/%d]: %lld of %lld states (global %lld...), segments %d..%d of %d, %d LM iterations, %d damped solves, "
                "error %.9e, %.3f ms per damped solve (%.1f ms on the device, %.4f s wall), t_off %.9f, p_BinI %.6f %.6f %.6f\n",
                rank, world, (long long)(pl.own_hi - pl.own_lo), (long long)pl.n_total, (long long)(pl.g_off + pl.own_lo), pl.j0, pl.j1,
                pl.p_total, inf.iterations, inf.tries, inf.final_error, ms[7] / std::max(1, inf.tries), ms[7], secs, toff, p[0], p[1], p[2]);
    if (rank == 0 && world == 1 && !path_states.empty()) {
      std::vector<double> t, x;
      cs.states(t, x);
      std::ofstream f(path_states);
      f.precision(17);
      for (size_t i = 0; i < t.size(); i++) {
        f << t[i];
        for (int k = 0; k < 16; k++)
          f << "," << x[16 * i + k];
        f << "\n";
      }
    }
    if (rank == 0 && world > 1)
      std::remove(id_file.c_str());
  } catch (const SolverError &e) {
    std::fprintf(stderr, "run_chain rank %d: %s (status %d)\n", rank, e.what(), e.code());
    return EXIT_FAILURE;
  }
  return EXIT_SUCCESS;
}
