/*
 * vicon2gt_b200 -- C++ host mirror of the reference's solver interface, header-only, over the C ABI
 * (vicon2gt_b200.h).  A user of rpng/vicon2gt replaces
 *
 *     Propagator / Interpolator / ViconGraphSolver            (src/meas/Propagator.h:41-55,
 *                                                              src/meas/Interpolator.h:55-81,
 *                                                              src/solver/ViconGraphSolver.h:61-168)
 *
 * by the classes of the same names below: same method names, argument meaning and order.  Differences,
 * all forced by dropping ROS / Eigen / GTSAM from the boundary:
 *   - the ros::NodeHandle argument of the solver becomes a ViconGraphSolverParams struct (same parameter
 *     names and defaults, ViconGraphSolver.cpp:34-88); Eigen vectors / matrices become std::array
 *     (row-major 3x3);
 *   - visualize() (ROS publishers) does not exist;
 *   - failures: the reference logs and calls std::exit(EXIT_FAILURE) (ViconGraphSolver.cpp:99-150,
 *     508-524).  Here they throw vicon2gt_b200::SolverError carrying the V2GT_ERR_* code; construct the
 *     solver with exit_on_failure = true to get the reference's behaviour.
 * Everything numeric runs in libvicon2gt_b200.so (CUDA, sm_100a); there is no CPU path.
 * BatchSolver is the batch form (independent Monte-Carlo trials on one GPU, scripts/run_monte_carlo.sh shape);
 * ChainSolver shards ONE long sequence over the GPUs of a node (NCCL inside the library).
 */
#ifndef VICON2GT_B200_HPP
#define VICON2GT_B200_HPP

#include "vicon2gt_b200.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <memory>
#include <exception>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace vicon2gt_b200 {

using Vec3 = std::array<double, 3>;
using Vec4 = std::array<double, 4>;  // JPL quaternion [x y z w]
using Mat3 = std::array<double, 9>;  // row-major
using Pose10 = std::array<double, 10>; // [q_VtoI(4) p_IinV(3) v_IinV(3)], as get_imu_poses of the reference

class SolverError : public std::runtime_error {
public:
  SolverError(int code, const std::string &where)
      : std::runtime_error("vicon2gt_b200: " + where + " failed with status " + std::to_string(code)), code_(code) {}
  int code() const { return code_; }

private:
  int code_;
};

inline void check(int status, const char *where) {
  if (status != V2GT_OK)
    throw SolverError(status, where);
}

/* Parameters the reference reads from its ros::NodeHandle (names / defaults: ViconGraphSolver.cpp:34-88). */
struct ViconGraphSolverParams : v2gt_params {
  ViconGraphSolverParams() { v2gt_params_default(this); }
};

/* IMU store (src/meas/Propagator.h:41-55): noise densities + samples in feed order. */
class Propagator {
public:
  Propagator(double sigma_w, double sigma_wb, double sigma_a, double sigma_ab)
      : sigma_w_(sigma_w), sigma_wb_(sigma_wb), sigma_a_(sigma_a), sigma_ab_(sigma_ab) {}

  void feed_imu(double timestamp, const Vec3 &wm, const Vec3 &am) {
    t_.push_back(timestamp);
    w_.insert(w_.end(), wm.begin(), wm.end());
    a_.insert(a_.end(), am.begin(), am.end());
  }
  /* Propagator::has_bounding_imu (Propagator.cpp:165-193) for the time-ordered store */
  bool has_bounding_imu(double timestamp) const { return !t_.empty() && t_.front() <= timestamp && t_.back() >= timestamp; }

  const std::vector<double> &times() const { return t_; }
  const std::vector<double> &gyro() const { return w_; }
  const std::vector<double> &accel() const { return a_; }
  double sigma_w() const { return sigma_w_; }
  double sigma_wb() const { return sigma_wb_; }
  double sigma_a() const { return sigma_a_; }
  double sigma_ab() const { return sigma_ab_; }

private:
  double sigma_w_, sigma_wb_, sigma_a_, sigma_ab_;
  std::vector<double> t_, w_, a_;
};

/* Vicon pose store (src/meas/Interpolator.h:55-81).  Sorting and the std::set de-duplication (first pose of a
 * timestamp wins) are done by the library when the trial is staged. */
class Interpolator {
public:
  void feed_pose(double timestamp, const Vec4 &q, const Vec3 &p, const Mat3 &R_q, const Mat3 &R_p) {
    t_.push_back(timestamp);
    q_.insert(q_.end(), q.begin(), q.end());
    p_.insert(p_.end(), p.begin(), p.end());
    Rq_.insert(Rq_.end(), R_q.begin(), R_q.end());
    Rp_.insert(Rp_.end(), R_p.begin(), R_p.end());
    if (t_.size() == 1 || timestamp < tmin_)
      tmin_ = timestamp;
    if (t_.size() == 1 || timestamp > tmax_)
      tmax_ = timestamp;
  }
  double get_time_min() const { return tmin_; }
  double get_time_max() const { return tmax_; }
  size_t size() const { return t_.size(); }

  const std::vector<double> &times() const { return t_; }
  const std::vector<double> &quats() const { return q_; }
  const std::vector<double> &positions() const { return p_; }
  const std::vector<double> &cov_q() const { return Rq_; }
  const std::vector<double> &cov_p() const { return Rp_; }

private:
  std::vector<double> t_, q_, p_, Rq_, Rp_;
  double tmin_ = 0, tmax_ = 0;
};

/* The reference's Simulator (src/sim/Simulator.h) for a whole Monte-Carlo batch on the GPU: one trajectory ("t x y z qx qy
 * qz qw" rows, JPL q_GtoI), one seed (+ vicon sigmas) per trial.  BatchSolver::set_trial_sim hands a simulated trial to the
 * solver on the device.  Parameters: v2gt_sim_params of vicon2gt_b200/host/v2gt_sim.h -- include that header before this one
 * to set them; without it every trial keeps the SimulatorParams defaults with seed = trial index. */
class SimBatch {
public:
  SimBatch(int n_trials, const std::vector<double> &rows8, int device = 0) : n_(n_trials) {
    check(v2gt_simbatch_create(device, n_trials, (int64_t)(rows8.size() / 8), rows8.data(), &h_), "v2gt_simbatch_create");
  }
  ~SimBatch() { v2gt_simbatch_destroy(h_); }
  SimBatch(const SimBatch &) = delete;
  SimBatch &operator=(const SimBatch &) = delete;
  int size() const { return n_; }
  v2gt_simbatch *handle() { return h_; }
  void set_trial(int trial, const struct v2gt_sim_params &p) { check(v2gt_simbatch_set_trial(h_, trial, &p), "v2gt_simbatch_set_trial"); }
  void plan() { check(v2gt_simbatch_plan(h_), "v2gt_simbatch_plan"); }
  void run() { check(v2gt_simbatch_run(h_), "v2gt_simbatch_run"); }
  /* Simulator::get_next_imu / get_next_vicon results of one trial, in feed order, plus the state-time grid */
  void get(int trial, std::vector<double> &imu_t, std::vector<double> &imu_w, std::vector<double> &imu_a, std::vector<double> &vic_t,
           std::vector<double> &vic_q, std::vector<double> &vic_p, std::vector<double> &state_t) {
    int64_t M = 0, V = 0, N = 0;
    check(v2gt_simbatch_sizes(h_, trial, &M, &V, &N), "v2gt_simbatch_sizes");
    imu_t.resize((size_t)M); imu_w.resize((size_t)M * 3); imu_a.resize((size_t)M * 3);
    vic_t.resize((size_t)V); vic_q.resize((size_t)V * 4); vic_p.resize((size_t)V * 3);
    state_t.resize((size_t)N);
    check(v2gt_simbatch_get(h_, trial, imu_t.data(), imu_w.data(), imu_a.data(), vic_t.data(), vic_q.data(), vic_p.data(), state_t.data()),
          "v2gt_simbatch_get");
  }
  void get_truth(int trial, Mat3 &R_BtoI, Vec3 &p_BinI, double &viconimu_dt, Mat3 &R_GtoV) {
    check(v2gt_simbatch_get_truth(h_, trial, R_BtoI.data(), p_BinI.data(), &viconimu_dt, R_GtoV.data()), "v2gt_simbatch_get_truth");
  }
  /* Simulator::get_state_in_vicon: rows of 17 = [t, q_VtoI(4), p_IinV(3), v_IinV(3), bg(3), ba(3)] */
  void get_state_in_vicon(int trial, const std::vector<double> &times, std::vector<double> &out17, std::vector<int32_t> &ok) {
    out17.resize(times.size() * 17);
    ok.resize(times.size());
    check(v2gt_simbatch_get_state_in_vicon(h_, trial, (int64_t)times.size(), times.data(), out17.data(), ok.data()),
          "v2gt_simbatch_get_state_in_vicon");
  }

private:
  int n_;
  v2gt_simbatch *h_ = nullptr;
};

/* A batch of independent build-and-solve problems resident on one GPU (owns the C handle). */
class BatchSolver {
public:
  explicit BatchSolver(int n_trials, int device = 0) : n_(n_trials) { check(v2gt_batch_create(device, n_trials, &h_), "v2gt_batch_create"); }
  ~BatchSolver() { v2gt_batch_destroy(h_); }
  BatchSolver(const BatchSolver &) = delete;
  BatchSolver &operator=(const BatchSolver &) = delete;

  int size() const { return n_; }
  v2gt_batch *handle() { return h_; }
  void set_option(const char *name, double value) { check(v2gt_batch_set_option(h_, name, value), "v2gt_batch_set_option"); }

  void set_trial(int trial, const v2gt_params &params, const Propagator &imu, const Interpolator &vicon,
                 const std::vector<double> &timestamp_cameras) {
    v2gt_params p = params;
    p.sigma_w = imu.sigma_w();
    p.sigma_wb = imu.sigma_wb();
    p.sigma_a = imu.sigma_a();
    p.sigma_ab = imu.sigma_ab();
    check(v2gt_batch_set_trial(h_, trial, &p, (int64_t)imu.times().size(), imu.times().data(), imu.gyro().data(), imu.accel().data(),
                               (int64_t)vicon.times().size(), vicon.times().data(), vicon.quats().data(), vicon.positions().data(),
                               vicon.cov_q().data(), vicon.cov_p().data(), nullptr, (int64_t)timestamp_cameras.size(),
                               timestamp_cameras.data()),
          "v2gt_batch_set_trial");
  }
  /* The same without the host copy: the stores and the timestamp vector are BORROWED until build_and_solve() (or an
   * explicit upload) has returned -- the lifetime the reference's shared_ptr stores already have. */
  void set_trial_view(int trial, const v2gt_params &params, const Propagator &imu, const Interpolator &vicon,
                      const std::vector<double> &timestamp_cameras) {
    v2gt_params p = params;
    p.sigma_w = imu.sigma_w();
    p.sigma_wb = imu.sigma_wb();
    p.sigma_a = imu.sigma_a();
    p.sigma_ab = imu.sigma_ab();
    check(v2gt_batch_set_trial_view(h_, trial, &p, (int64_t)imu.times().size(), imu.times().data(), imu.gyro().data(),
                                    imu.accel().data(), (int64_t)vicon.times().size(), vicon.times().data(), vicon.quats().data(),
                                    vicon.positions().data(), vicon.cov_q().data(), vicon.cov_p().data(), nullptr,
                                    (int64_t)timestamp_cameras.size(), timestamp_cameras.data()),
          "v2gt_batch_set_trial_view");
  }
  /* Trial `sim_trial` of a simulated batch on the same GPU (measurements copied device -> device at upload); vic_R_const =
   * (R_q, R_p) of run_simulation.cpp:84-96, 18 doubles.  `sim` must outlive build_and_solve(). */
  void set_trial_sim(int trial, const v2gt_params &params, SimBatch &sim, int sim_trial, const std::array<double, 18> &vic_R_const) {
    check(v2gt_batch_set_trial_sim(h_, trial, &params, sim.handle(), sim_trial, vic_R_const.data()), "v2gt_batch_set_trial_sim");
  }
  void build_and_solve() { check(v2gt_batch_build_and_solve(h_), "v2gt_batch_build_and_solve"); }
  /* CSV + info (+ pose text) files of every solved trial under `dir`; returns the number of trials written. */
  int write_all(const std::string &dir, bool with_txt = true) {
    int32_t n = 0;
    check(v2gt_batch_write_all(h_, dir.c_str(), with_txt ? 1 : 0, &n), "v2gt_batch_write_all");
    return (int)n;
  }
  /* Error report of a simulated batch (run_simulation.cpp:198-297, stats.h:64-112) for all trials at once:
   * truth10 = [n][10] (q_VtoI, p_IinV, v_IinV at the trial's surviving state times); stats = [n_trials][3][8] with rows
   * (ori deg, pos m, vel m/s) and columns rmse, mean, median, min, max, std, mean + 2.326 std, count. */
  void set_truth(int trial, const std::vector<double> &truth10) {
    check(v2gt_batch_set_truth(h_, trial, (int64_t)(truth10.size() / 10), truth10.data()), "v2gt_batch_set_truth");
  }
  std::vector<double> error_stats() {
    std::vector<double> out((size_t)n_ * 24, 0.0);
    check(v2gt_batch_error_stats(h_, out.data()), "v2gt_batch_error_stats");
    return out;
  }
  v2gt_trial_info info(int trial) {
    v2gt_trial_info inf;
    check(v2gt_batch_get_info(h_, trial, &inf), "v2gt_batch_get_info");
    return inf;
  }
  void write_to_file(int trial, const std::string &csv, const std::string &info) {
    check(v2gt_batch_write_to_file(h_, trial, csv.c_str(), info.c_str()), "v2gt_batch_write_to_file");
  }
  /* times[n], states[n][16] = [q4 bg3 v3 ba3 p3] */
  void get_states(int trial, std::vector<double> &times, std::vector<double> &states) {
    int64_t n = 0;
    check(v2gt_batch_get_states(h_, trial, 0, nullptr, nullptr, &n), "v2gt_batch_get_states");
    times.assign((size_t)n, 0.0);
    states.assign((size_t)n * V2GT_STATE_DIM, 0.0);
    check(v2gt_batch_get_states(h_, trial, n, times.data(), states.data(), &n), "v2gt_batch_get_states");
  }
  /* every trial in one read-back: trial t owns rows offsets[t] .. offsets[t+1] of times / states (16 per row); globals10 = 10 per trial */
  void get_all_states(std::vector<int64_t> &offsets, std::vector<double> &times, std::vector<double> &states, std::vector<double> &globals10) {
    offsets.assign((size_t)n_ + 1, 0);
    check(v2gt_batch_get_all_states(h_, 0, offsets.data(), nullptr, nullptr, nullptr), "v2gt_batch_get_all_states");
    const size_t n = (size_t)offsets.back();
    times.resize(n);
    states.resize(n * 16);
    globals10.resize((size_t)n_ * 10);
    check(v2gt_batch_get_all_states(h_, (int64_t)n, offsets.data(), times.data(), states.data(), globals10.data()), "v2gt_batch_get_all_states");
  }
  void get_calibration(int trial, Vec4 &q_BtoI, Vec3 &p_BinI, double &toff, std::array<double, 2> &theta_xy) {
    check(v2gt_batch_get_calibration(h_, trial, q_BtoI.data(), p_BinI.data(), &toff, theta_xy.data()), "v2gt_batch_get_calibration");
  }

private:
  int n_;
  v2gt_batch *h_ = nullptr;
};

/* A Monte-Carlo batch larger than one GPU's memory, solved in waves of `wave` trials (scripts/run_monte_carlo.sh runs its
 * trials one roslaunch after the other).  Two resident batch handles: while wave k is being solved, one host thread stages
 * and uploads wave k+1 into the other handle (pinned fill + H2D copies on that handle's stream run under wave k's kernels)
 * and another hands wave k-1 to `on_wave(first_trial, solver)` (read the results there), so the GPU never waits for the
 * host.  The stores and timestamp vectors of `trials` are borrowed until run() returns. */
class WaveRunner {
public:
  struct Trial {
    const v2gt_params *params;
    const Propagator *imu;
    const Interpolator *vicon;
    const std::vector<double> *timestamp_cameras;
  };
  explicit WaveRunner(int wave, int device = 0) : wave_(wave), device_(device) {}
  double device_ms() const { return device_ms_; } /* CUDA-event time of build + optimise, summed over the waves */

  template <class F> void run(const std::vector<Trial> &trials, F on_wave) {
    const int n = (int)trials.size(), n_waves = (n + wave_ - 1) / wave_;
    device_ms_ = 0;
    std::exception_ptr err;
    auto stage = [&](int w) {
      try {
        const int a = w * wave_, cnt = std::min(wave_, n - a);
        std::unique_ptr<BatchSolver> &h = h_[w % 2];
        if (!h || h->size() != cnt)
          h.reset(new BatchSolver(cnt, device_));
        for (int i = 0; i < cnt; i++)
          h->set_trial_view(i, *trials[a + i].params, *trials[a + i].imu, *trials[a + i].vicon, *trials[a + i].timestamp_cameras);
        check(v2gt_batch_upload(h->handle()), "v2gt_batch_upload");
      } catch (...) {
        err = std::current_exception();
      }
    };
    std::thread up(stage, 0), rb;
    for (int w = 0; w < n_waves; w++) {
      up.join();
      if (err)
        break;
      if (w + 1 < n_waves) {
        if (rb.joinable()) /* wave w-1 lives in the handle wave w+1 is about to overwrite */
          rb.join();
        up = std::thread(stage, w + 1);
      }
      BatchSolver &h = *h_[w % 2];
      h.build_and_solve();
      double ms[8];
      check(v2gt_batch_get_timing(h.handle(), ms, nullptr), "v2gt_batch_get_timing");
      device_ms_ += ms[0] + ms[7];
      if (rb.joinable())
        rb.join();
      rb = std::thread([&, w]() {
        try {
          on_wave(w * wave_, *h_[w % 2]);
        } catch (...) {
          err = std::current_exception();
        }
      });
    }
    if (up.joinable())
      up.join();
    if (rb.joinable())
      rb.join();
    if (err)
      std::rethrow_exception(err);
  }

private:
  int wave_, device_;
  double device_ms_ = 0;
  std::unique_ptr<BatchSolver> h_[2];
};

/* Single-trajectory solver with the reference's method names (src/solver/ViconGraphSolver.h:71-110). */
class ViconGraphSolver {
public:
  ViconGraphSolver(const ViconGraphSolverParams &params, std::shared_ptr<Propagator> propagator,
                   std::shared_ptr<Interpolator> interpolator, std::vector<double> timestamp_cameras, int device = 0,
                   bool exit_on_failure = false)
      : params_(params), propagator_(std::move(propagator)), interpolator_(std::move(interpolator)),
        timestamp_cameras_(std::move(timestamp_cameras)), exit_on_failure_(exit_on_failure) {
    try {
      batch_.reset(new BatchSolver(1, device));
    } catch (const SolverError &e) {
      fail(e);
    }
  }

  /* ViconGraphSolver::build_and_solve (.cpp:96-192) */
  void build_and_solve() {
    try {
      batch_->set_trial(0, params_, *propagator_, *interpolator_, timestamp_cameras_);
      batch_->build_and_solve();
      const v2gt_trial_info inf = batch_->info(0);
      if (inf.status != V2GT_OK)
        throw SolverError(inf.status, "build_and_solve");
      info_ = inf;
    } catch (const SolverError &e) {
      fail(e);
    }
  }
  /* ViconGraphSolver::write_to_file (.cpp:194-248): byte-compatible CSV + info text */
  void write_to_file(const std::string &csvfilepath, const std::string &infofilepath) {
    batch_->write_to_file(0, csvfilepath, infofilepath);
  }
  /* ViconGraphSolver::get_imu_poses (.cpp:357-375) */
  void get_imu_poses(std::vector<double> &times, std::vector<Pose10> &poses) {
    std::vector<double> x;
    batch_->get_states(0, times, x);
    poses.resize(times.size());
    for (size_t i = 0; i < times.size(); i++) {
      const double *s = &x[i * V2GT_STATE_DIM];
      poses[i] = {s[0], s[1], s[2], s[3], s[13], s[14], s[15], s[7], s[8], s[9]};
    }
  }
  /* full 16-double state rows [q4 bg3 v3 ba3 p3] (the CSV writer's inputs) */
  void get_states(std::vector<double> &times, std::vector<double> &states) { batch_->get_states(0, times, states); }
  /* ViconGraphSolver::get_calibration (.cpp:377-388) */
  void get_calibration(double &toff, Mat3 &R_BtoI, Vec3 &p_BinI, Mat3 &R_GtoV) {
    Vec4 q;
    std::array<double, 2> th;
    batch_->get_calibration(0, q, p_BinI, toff, th);
    quat_2_Rot(q, R_BtoI);
    rot_xy(th[0], th[1], R_GtoV);
  }
  const v2gt_trial_info &info() const { return info_; }

  /* quat_2_Rot (src/utils/quat_ops.h:153-158) and rot_y(ty) rot_x(tx) (src/gtsam/RotationXY.h:62) on plain arrays */
  static void quat_2_Rot(const Vec4 &q, Mat3 &R) {
    const double x = q[0], y = q[1], z = q[2], w = q[3], a = 2 * w * w - 1;
    R = {a + 2 * x * x,         2 * w * z + 2 * x * y, -2 * w * y + 2 * x * z,
         -2 * w * z + 2 * y * x, a + 2 * y * y,         2 * w * x + 2 * y * z,
         2 * w * y + 2 * z * x, -2 * w * x + 2 * z * y, a + 2 * z * z};
  }
  static void rot_xy(double tx, double ty, Mat3 &R);

private:
  void fail(const SolverError &e) const {
    if (exit_on_failure_) {
      std::fprintf(stderr, "[vicon2gt_b200] %s\n", e.what());
      std::exit(EXIT_FAILURE);
    }
    throw e;
  }
  ViconGraphSolverParams params_;
  std::shared_ptr<Propagator> propagator_;
  std::shared_ptr<Interpolator> interpolator_;
  std::vector<double> timestamp_cameras_;
  bool exit_on_failure_;
  std::unique_ptr<BatchSolver> batch_;
  v2gt_trial_info info_{};
};

/* One long sequence sharded over the GPUs of a node as contiguous chain segments (BASELINE config 5): the C++ face of the
 * chain mode of the C ABI.  One process (or host thread) per GPU constructs a ChainSolver with its rank; the constructor
 * does what ViconGraphSolver::build_and_solve does up to the optimiser for THIS rank's piece of the chain (timestamp
 * cleaning of the whole sequence, src/solver/ViconGraphSolver.cpp:114-142; the rank's slice of the measurements; state
 * initialisation and preintegration on its GPU).  Rank 0 then obtains an NCCL id (unique_id()), hands it to the other
 * ranks by any means, every rank calls comm_init(id) and optimize(): the Levenberg-Marquardt loop runs inside the
 * library, two ncclAllGather calls per damped solve on the solver's stream, one damped solve replayed as a CUDA graph.
 * All ranks end with the same calibration; states() returns the rank's own states.  world == 1 needs no NCCL. */
class ChainSolver {
public:
  using NcclId = std::array<uint8_t, V2GT_NCCL_ID_BYTES>;

  ChainSolver(const ViconGraphSolverParams &params, const Propagator &imu, const Interpolator &vicon,
              const std::vector<double> &timestamp_cameras, int rank, int world, int segs_per_rank = 16, int device = 0)
      : batch_(1, device) {
    v2gt_params p = params;
    p.sigma_w = imu.sigma_w();
    p.sigma_wb = imu.sigma_wb();
    p.sigma_a = imu.sigma_a();
    p.sigma_ab = imu.sigma_ab();
    const std::vector<double> &it = imu.times(), &vt = vicon.times();
    if (it.empty() || vt.size() < 2)
      throw SolverError(V2GT_ERR_NO_VICON, "ChainSolver");
    // ---- timestamp cleaning of the whole sequence (time-ordered stores)
    const double first = vicon.get_time_min() + p.toff_imu_to_vicon + 2 * p.time_offset_window;
    const double last = vicon.get_time_max() + p.toff_imu_to_vicon - 2 * p.time_offset_window;
    std::vector<double> ts;
    for (double t : timestamp_cameras)
      if (t >= it.front() && t <= it.back() && t >= first && t <= last)
        ts.push_back(t);
    int32_t spr = 0;
    check(v2gt_chain_fit_segments((int64_t)ts.size(), world, segs_per_rank, &spr), "v2gt_chain_fit_segments");
    check(v2gt_chain_plan((int64_t)ts.size(), world, spr, rank, &plan_), "v2gt_chain_plan");
    // ---- this rank's states (own + one halo state on each inner side) and the measurements around them
    const std::vector<double> local(ts.begin() + plan_.g_off, ts.begin() + plan_.g_off + plan_.n_local);
    const double margin = std::abs(p.toff_imu_to_vicon) + 2 * p.time_offset_window + p.interp_gap_init + 1.0;
    const size_t v0 = lower_index(vt, local.front() - margin, 1), v1 = upper_index(vt, local.back() + margin, 1);
    const size_t i0 = lower_index(it, local.front() - 0.05, 2), i1 = upper_index(it, local.back() + 0.05, 2);
    check(v2gt_batch_set_trial(batch_.handle(), 0, &p, (int64_t)(i1 - i0), it.data() + i0, imu.gyro().data() + 3 * i0,
                               imu.accel().data() + 3 * i0, (int64_t)(v1 - v0), vt.data() + v0, vicon.quats().data() + 4 * v0,
                               vicon.positions().data() + 3 * v0, vicon.cov_q().data() + 9 * v0, vicon.cov_p().data() + 9 * v0, nullptr,
                               (int64_t)local.size(), local.data()),
          "v2gt_batch_set_trial");
    check(v2gt_chain_configure(batch_.handle(), rank, world, plan_.n_total, plan_.g_off, plan_.p_total, plan_.j0, plan_.j1, plan_.own_lo,
                               plan_.own_hi),
          "v2gt_chain_configure");
    check(v2gt_batch_upload(batch_.handle()), "v2gt_batch_upload");
    build();
  }
  /* state initialisation + preintegration from the uploaded inputs (a fresh solve) */
  void build() {
    check(v2gt_batch_build(batch_.handle(), 1), "v2gt_batch_build");
    const v2gt_trial_info inf = batch_.info(0);
    if (inf.status != V2GT_OK || inf.n_states != plan_.n_local)
      throw SolverError(inf.status ? inf.status : V2GT_ERR_INVALID_ARG, "chain build: local states lost in the rank's slice");
  }
  static NcclId unique_id() {
    NcclId id;
    check(v2gt_chain_nccl_unique_id(id.data()), "v2gt_chain_nccl_unique_id");
    return id;
  }
  void comm_init(const NcclId &id) { check(v2gt_chain_comm_init(batch_.handle(), id.data()), "v2gt_chain_comm_init"); }
  /* Levenberg-Marquardt over the whole chain (collective when world > 1); returns the replicated optimiser summary */
  v2gt_trial_info optimize(int max_tries = 0, bool use_graph = true) {
    check(v2gt_chain_optimize(batch_.handle(), max_tries, use_graph ? 1 : 0), "v2gt_chain_optimize");
    return batch_.info(0);
  }
  const v2gt_chain_plan_t &plan() const { return plan_; }
  /* the rank's OWN states: global index of the first one, times[n], states[n][16] */
  int64_t states(std::vector<double> &times, std::vector<double> &states16) {
    std::vector<double> t, x;
    batch_.get_states(0, t, x);
    times.assign(t.begin() + plan_.own_lo, t.begin() + plan_.own_hi);
    states16.assign(x.begin() + 16 * (size_t)plan_.own_lo, x.begin() + 16 * (size_t)plan_.own_hi);
    return plan_.g_off + plan_.own_lo;
  }
  void get_calibration(Vec4 &q_BtoI, Vec3 &p_BinI, double &toff, std::array<double, 2> &theta_xy) {
    batch_.get_calibration(0, q_BtoI, p_BinI, toff, theta_xy);
  }
  /* [build, linearize, eliminate + pack, exchange A + reduced, backsub, cost, exchange B + decide, optimise total] ms */
  std::array<double, 8> timing_ms() {
    std::array<double, 8> ms{};
    check(v2gt_batch_get_timing(batch_.handle(), ms.data(), nullptr), "v2gt_batch_get_timing");
    return ms;
  }
  BatchSolver &batch() { return batch_; }

private:
  // first index whose time is >= t, moved `back` samples towards the start / one past the last index whose time is <= t,
  // moved `fwd` samples towards the end (sorted times)
  static size_t lower_index(const std::vector<double> &v, double t, size_t back) {
    const size_t i = (size_t)(std::lower_bound(v.begin(), v.end(), t) - v.begin());
    return i > back ? i - back : 0;
  }
  static size_t upper_index(const std::vector<double> &v, double t, size_t fwd) {
    const size_t i = (size_t)(std::upper_bound(v.begin(), v.end(), t) - v.begin());
    return std::min(v.size(), i + fwd);
  }
  BatchSolver batch_;
  v2gt_chain_plan_t plan_{};
};

} // namespace vicon2gt_b200

#include <cmath>
inline void vicon2gt_b200::ViconGraphSolver::rot_xy(double tx, double ty, Mat3 &R) {
  const double cx = std::cos(tx), sx = std::sin(tx), cy = std::cos(ty), sy = std::sin(ty);
  // rot_y(ty) * rot_x(tx)   (src/utils/rpy_ops.h:28-46)
  R = {cy, sy * sx, sy * cx, 0.0, cx, -sx, -sy, cy * sx, cy * cx};
}

#endif /* VICON2GT_B200_HPP */
