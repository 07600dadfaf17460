// ORACLE / TEST INFRASTRUCTURE -- not product code, never linked into libvicon2gt_b200.so.
//
// C entry points over the REFERENCE's own sources, compiled where they lie under /root/reference/src
// (oracle/Makefile.ref; output oracle/_ref/libv2gt_ref.so, git-ignored) against the interface
// stand-ins of oracle/shim/ (Eigen and the GTSAM / Boost type scaffolding are not in this image).
// The functions mirror the factor-level entry points of oracle/oracle_c.cpp (orc_*) so that
// tests/test_ref_pin_cpu.py can compare the restatement with the reference call by call, and
// tools/make_ref_golden.py can freeze the reference's outputs as tests/golden/ref_pin.npz.
//
// Row layouts (shared with oracle_c.cpp):  state x16 = q4 | bg3 | v3 | ba3 | p3,
// globals10 = q_BtoI4 | p_BinI3 | t_off | theta_x | theta_y,  matrices row-major.
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <vector>

#include "cpi/CpiV1.h"
#include "gtsam/GtsamConfig.h"
#include "gtsam/ImuFactorCPIv1.h"
#include "gtsam/JPLNavState.h"
#include "gtsam/JPLQuaternion.h"
#include "gtsam/MeasBased_ViconPoseTimeoffsetFactor.h"
#include "gtsam/RotationXY.h"
#include "meas/Interpolator.h"
#include "meas/Propagator.h"
#include "sim/Simulator.h"
#include "solver/ViconGraphSolver.h"
#include "utils/quat_ops.h"
#include "utils/rpy_ops.h"
#include "utils/stats.h"

namespace {

Eigen::Vector3d v3(const double *p) { return Eigen::Vector3d(p[0], p[1], p[2]); }
Eigen::Vector4d v4(const double *p) { return Eigen::Vector4d(p[0], p[1], p[2], p[3]); }
Eigen::Matrix3d m3(const double *p) {
  Eigen::Matrix3d m;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++)
      m(i, j) = p[3 * i + j];
  return m;
}
template <class M> void put(const M &m, double *o) {
  if (!o)
    return;
  for (int i = 0; i < m.rows(); i++)
    for (int j = 0; j < m.cols(); j++)
      o[i * m.cols() + j] = m(i, j);
}
gtsam::JPLNavState nav(double t, const double *x) {
  return gtsam::JPLNavState(t, v4(x), v3(x + 4), v3(x + 7), v3(x + 10), v3(x + 13));
}
void put_nav(const gtsam::JPLNavState &s, double *o) {
  put(s.q(), o);
  put(s.bg(), o + 4);
  put(s.v(), o + 7);
  put(s.ba(), o + 10);
  put(s.p(), o + 13);
}

struct InterpHandle {
  std::shared_ptr<Interpolator> interp = std::make_shared<Interpolator>();
  std::vector<double> stamps; // sorted unique stamps, to turn get_bounds() times into bracket indices
};

} // namespace

extern "C" {

// ---- src/utils/quat_ops.h, src/utils/rpy_ops.h
void ref_rot_2_quat(const double *R, double *q) { put(rot_2_quat(m3(R)), q); }
void ref_quat_2_Rot(const double *q, double *R) { put(quat_2_Rot(v4(q)), R); }
void ref_quat_multiply(const double *q, const double *p, double *o) { put(quat_multiply(v4(q), v4(p)), o); }
void ref_exp_so3(const double *w, double *R) { put(exp_so3(v3(w)), R); }
void ref_log_so3(const double *R, double *w) { put(log_so3(m3(R)), w); }
void ref_Jl_so3(const double *w, double *J) { put(Jl_so3(v3(w)), J); }
void ref_rot2rpy(const double *R, double *rpy) { put(rot2rpy(m3(R)), rpy); }

// ---- src/meas/Propagator.cpp (feed_imu, has_bounding_imu, propagate) + src/cpi/CpiV1.h
// out[288] as pack_preint of oracle_c.cpp (the information matrix, out[63..], is GTSAM's and left 0);
// cov[225] = P_meas.  returns 1 if propagate() succeeded, 0 if it returned false.
int ref_propagate(const double *sigmas, int64_t n_imu, const double *imu_t, const double *imu_w, const double *imu_a, double time0,
                  double time1, const double *bg_lin, const double *ba_lin, double *out, double *cov) {
  Propagator prop(sigmas[0], sigmas[1], sigmas[2], sigmas[3]);
  for (int64_t i = 0; i < n_imu; i++)
    prop.feed_imu(imu_t[i], v3(imu_w + 3 * i), v3(imu_a + 3 * i));
  CpiV1 pre(0, 0, 0, 0, true);
  if (!prop.propagate(time0, time1, v3(bg_lin), v3(ba_lin), pre))
    return 0;
  if (out) {
    std::memset(out, 0, sizeof(double) * 288);
    out[0] = pre.DT;
    put(pre.alpha_tau, out + 1);
    put(pre.beta_tau, out + 4);
    put(pre.q_k2tau, out + 7);
    put(pre.b_w_lin, out + 11);
    put(pre.b_a_lin, out + 14);
    put(pre.J_q, out + 17);
    put(pre.J_b, out + 26);
    put(pre.J_a, out + 35);
    put(pre.H_b, out + 44);
    put(pre.H_a, out + 53);
  }
  put(pre.P_meas, cov);
  return 1;
}

int ref_has_bounding_imu(int64_t n_imu, const double *imu_t, int64_t n, const double *t_query, int32_t *out) {
  Propagator prop(0, 0, 0, 0);
  Eigen::Vector3d z(0, 0, 0);
  for (int64_t i = 0; i < n_imu; i++)
    prop.feed_imu(imu_t[i], z, z);
  for (int64_t i = 0; i < n; i++)
    out[i] = prop.has_bounding_imu(t_query[i]) ? 1 : 0;
  return 0;
}

// ---- src/meas/Interpolator.cpp
void *ref_interp_create(int64_t n, const double *t, const double *q, const double *p, const double *Rq, const double *Rp,
                        int per_pose_cov) {
  InterpHandle *h = new InterpHandle();
  for (int64_t i = 0; i < n; i++) {
    const double *rq = per_pose_cov ? Rq + 9 * i : Rq;
    const double *rp = per_pose_cov ? Rp + 9 * i : Rp;
    h->interp->feed_pose(t[i], v4(q + 4 * i), v3(p + 3 * i), m3(rq), m3(rp));
  }
  for (const POSEDATA &d : h->interp->get_raw_poses())
    h->stamps.push_back(d.timestamp);
  return h;
}
void ref_interp_destroy(void *hv) { delete (InterpHandle *)hv; }
int64_t ref_interp_size(void *hv) { return (int64_t)((InterpHandle *)hv)->stamps.size(); }

// with_jacobian != 0: get_pose_with_jacobian (0.1 s gap threshold), else get_pose (5 s).
// idx0 / idx1 come from get_bounds() (the same equal_range + decrement) mapped to positions in the
// sorted unique store; -1 where get_bounds() returns false.  Outputs of a call that returned false
// without touching them are NaN.
int ref_eval_interp(void *hv, int64_t n, const double *t_query, int with_jacobian, int32_t *idx0, int32_t *idx1, int32_t *ok, double *q,
                    double *p, double *R, double *H_toff) {
  InterpHandle *h = (InterpHandle *)hv;
  const double nan = std::numeric_limits<double>::quiet_NaN();
  for (int64_t i = 0; i < n; i++) {
    Eigen::Vector4d qi(nan, nan, nan, nan);
    Eigen::Vector3d pi(nan, nan, nan);
    Eigen::Matrix<double, 6, 6> Ri = Eigen::Matrix<double, 6, 6>::Constant(nan);
    Eigen::Matrix<double, 6, 1> Hi = Eigen::Matrix<double, 6, 1>::Constant(nan);
    bool good = with_jacobian ? h->interp->get_pose_with_jacobian(t_query[i], qi, pi, Ri, Hi) : h->interp->get_pose(t_query[i], qi, pi, Ri);
    if (ok)
      ok[i] = good ? 1 : 0;
    put(qi, q ? q + 4 * i : nullptr);
    put(pi, p ? p + 3 * i : nullptr);
    put(Ri, R ? R + 36 * i : nullptr);
    put(Hi, H_toff ? H_toff + 6 * i : nullptr);
    if (idx0 || idx1) {
      double t0, t1;
      Eigen::Vector4d q0, q1;
      Eigen::Vector3d p0, p1;
      Eigen::Matrix<double, 6, 6> R0, R1;
      int32_t i0 = -1, i1 = -1;
      if (h->interp->get_bounds(t_query[i], t0, q0, p0, R0, t1, q1, p1, R1)) {
        i0 = (int32_t)(std::lower_bound(h->stamps.begin(), h->stamps.end(), t0) - h->stamps.begin());
        i1 = (int32_t)(std::lower_bound(h->stamps.begin(), h->stamps.end(), t1) - h->stamps.begin());
      }
      if (idx0)
        idx0[i] = i0;
      if (idx1)
        idx1[i] = i1;
    }
  }
  return 0;
}

// ---- src/gtsam/ImuFactorCPIv1.cpp  (unwhitened error and Jacobians, exactly what evaluateError returns)
int ref_eval_imu_factor(const double *preint288, double grav_mag, const double *xi, const double *xj, const double *thetaxy, double *e15,
                        double *H1, double *H2, double *H3) {
  const double *o = preint288;
  gtsam::ImuFactorCPIv1 f(1, 2, 3, Eigen::Matrix<double, 15, 15>::Identity(), o[0], grav_mag, v3(o + 1), v3(o + 4), v4(o + 7), v3(o + 14),
                          v3(o + 11), m3(o + 17), m3(o + 26), m3(o + 35), m3(o + 44), m3(o + 53));
  gtsam::Matrix h1, h2, h3;
  gtsam::Vector e = f.evaluateError(nav(0.0, xi), nav(o[0], xj), gtsam::RotationXY(thetaxy[0], thetaxy[1]), h1, h2, h3);
  put(e, e15);
  put(h1, H1);
  put(h2, H2);
  put(h3, H3);
  return 0;
}

// ---- src/gtsam/MeasBased_ViconPoseTimeoffsetFactor.cpp  (error / Jacobians as evaluateError returns
// them, i.e. already multiplied by the factor's own sqrt_inv_interp; the Huber weight is GTSAM's)
int ref_eval_vicon_factor(void *hv, const int32_t *estimate_flags3, double t, const double *x16, const double *globals10, double *e6,
                          double *H1, double *H2, double *H3, double *H4) {
  InterpHandle *h = (InterpHandle *)hv;
  std::shared_ptr<GtsamConfig> cfg = std::make_shared<GtsamConfig>();
  cfg->estimate_vicon_imu_toff = estimate_flags3[0] != 0;
  cfg->estimate_vicon_imu_ori = estimate_flags3[1] != 0;
  cfg->estimate_vicon_imu_pos = estimate_flags3[2] != 0;
  gtsam::MeasBased_ViconPoseTimeoffsetFactor f(1, 2, 3, 4, h->interp, cfg);
  gtsam::Vector1 toff;
  toff(0) = globals10[7];
  gtsam::Matrix h1, h2, h3, h4;
  gtsam::Vector e = f.evaluateError(nav(t, x16), gtsam::JPLQuaternion(v4(globals10)), v3(globals10 + 4), toff, h1, h2, h3, h4);
  put(e, e6);
  put(h1, H1);
  put(h2, H2);
  put(h3, H3);
  put(h4, H4);
  return 0;
}

// ---- src/gtsam/JPLNavState.cpp, JPLQuaternion.cpp, RotationXY.cpp
int ref_retract_state(const double *x16, const double *xi15, double *out16) {
  gtsam::Vector15 xi;
  for (int d = 0; d < 15; d++)
    xi(d) = xi15[d];
  put_nav(nav(0.0, x16).retract(xi), out16);
  return 0;
}
int ref_local_state(const double *x16, const double *y16, double *out15) {
  put(nav(0.0, x16).localCoordinates(nav(0.0, y16)), out15);
  return 0;
}
int ref_retract_quat(const double *q4, const double *d3, double *out4) {
  put(gtsam::JPLQuaternion(v4(q4)).retract(v3(d3)).q(), out4);
  return 0;
}
int ref_retract_rotxy(const double *th, const double *d, double *out) {
  gtsam::RotationXY r = gtsam::RotationXY(th[0], th[1]).retract(gtsam::Vector2(d[0], d[1]));
  out[0] = r.thetax();
  out[1] = r.thetay();
  return 0;
}
int ref_rotxy_rot(const double *th, double *R) {
  put(gtsam::RotationXY(th[0], th[1]).rot(), R);
  return 0;
}

// ---- src/solver/ViconGraphSolver.cpp: the reference's solver class end to end.  Timestamp cleaning, state erasure,
// initial guess, graph construction, the relinearisation loop and write_to_file are the reference's; the optimiser
// behind optimizer.optimize() is the GTSAM restatement of shim/gtsam/mini_gtsam_lm.h (dense, generic).
namespace {
struct SolverAccess : ViconGraphSolver {
  using ViconGraphSolver::ViconGraphSolver;
  const std::vector<double> &times() const { return timestamp_cameras; }
  const gtsam::Values &result() const { return values_result; }
  size_t id_of(double t) { return map_states[t]; }
};
struct SolverHandle {
  std::shared_ptr<Propagator> prop;
  std::shared_ptr<Interpolator> interp;
  std::unique_ptr<SolverAccess> solver;
};
} // namespace

// params13: gravity_magnitude, toff_imu_to_vicon, estimate_toff, estimate_ori, estimate_pos, num_loop_relin,
//           sigma_w, sigma_wb, sigma_a, sigma_ab, max_iterations_override (-1 = the reference's 30), 0, 0
void *ref_solver_create(const double *R_GtoV9, const double *R_BtoI9, const double *p_BinI3, const double *params13, int64_t n_imu,
                        const double *imu_t, const double *imu_w, const double *imu_a, int64_t n_vic, const double *vic_t,
                        const double *vic_q, const double *vic_p, const double *Rq, const double *Rp, int per_pose_cov, int64_t n_times,
                        const double *state_t) {
  ros::ParamTable &pt = ros::ParamTable::instance();
  pt.scalars.clear();
  pt.vectors.clear();
  pt.vectors["R_GtoV"] = std::vector<double>(R_GtoV9, R_GtoV9 + 9);
  pt.vectors["R_BtoI"] = std::vector<double>(R_BtoI9, R_BtoI9 + 9);
  pt.vectors["p_BinI"] = std::vector<double>(p_BinI3, p_BinI3 + 3);
  pt.scalars["gravity_magnitude"] = params13[0];
  pt.scalars["toff_imu_to_vicon"] = params13[1];
  pt.scalars["estimate_toff_vicon_to_imu"] = params13[2];
  pt.scalars["estimate_ori_vicon_to_imu"] = params13[3];
  pt.scalars["estimate_pos_vicon_to_imu"] = params13[4];
  pt.scalars["num_loop_relin"] = params13[5];
  gtsam::MiniLmLog::instance() = gtsam::MiniLmLog();
  gtsam::MiniLmLog::instance().max_iterations_override = (int)params13[10];
  SolverHandle *h = new SolverHandle();
  h->prop = std::make_shared<Propagator>(params13[6], params13[7], params13[8], params13[9]);
  for (int64_t i = 0; i < n_imu; i++)
    h->prop->feed_imu(imu_t[i], v3(imu_w + 3 * i), v3(imu_a + 3 * i));
  h->interp = std::make_shared<Interpolator>();
  for (int64_t i = 0; i < n_vic; i++)
    h->interp->feed_pose(vic_t[i], v4(vic_q + 4 * i), v3(vic_p + 3 * i), m3(per_pose_cov ? Rq + 9 * i : Rq), m3(per_pose_cov ? Rp + 9 * i : Rp));
  ros::NodeHandle nh;
  std::vector<double> times(state_t, state_t + n_times);
  std::streambuf *old = std::cout.rdbuf(nullptr); // the constructor and build_and_solve print their parameters / result
  h->solver.reset(new SolverAccess(nh, h->prop, h->interp, times));
  std::cout.rdbuf(old);
  return h;
}
void ref_solver_destroy(void *hv) { delete (SolverHandle *)hv; }
int ref_solver_build_and_solve(void *hv) {
  SolverHandle *h = (SolverHandle *)hv;
  std::streambuf *old = std::cout.rdbuf(nullptr);
  h->solver->build_and_solve();
  std::cout.rdbuf(old);
  return 0;
}
int64_t ref_solver_n_states(void *hv) { return (int64_t)((SolverHandle *)hv)->solver->times().size(); }
// states: [n][16] = q bg v ba p;  calib10 = q_BtoI4 p_BinI3 t_off theta_x theta_y;  lm6 = iterations, damped solves,
// linearisations, initial error, final error, lambda (summed over the relinearisation loops)
int ref_solver_get(void *hv, double *times, double *states, double *calib10, double *lm6) {
  using namespace gtsam::symbol_shorthand;
  SolverHandle *h = (SolverHandle *)hv;
  const std::vector<double> &t = h->solver->times();
  const gtsam::Values &v = h->solver->result();
  for (size_t i = 0; i < t.size(); i++) {
    if (times)
      times[i] = t[i];
    if (states)
      put_nav(v.at<gtsam::JPLNavState>(X(h->solver->id_of(t[i]))), states + 16 * i);
  }
  if (calib10) {
    put(v.at<gtsam::JPLQuaternion>(C(0)).q(), calib10);
    put(v.at<gtsam::Vector3>(C(1)), calib10 + 4);
    calib10[7] = v.at<gtsam::Vector1>(T(0))(0);
    calib10[8] = v.at<gtsam::RotationXY>(G(0)).thetax();
    calib10[9] = v.at<gtsam::RotationXY>(G(0)).thetay();
  }
  if (lm6) {
    const gtsam::MiniLmLog &l = gtsam::MiniLmLog::instance();
    lm6[0] = l.iterations, lm6[1] = l.tries, lm6[2] = l.linearizations;
    lm6[3] = l.initial_error, lm6[4] = l.final_error, lm6[5] = l.lambda;
  }
  return 0;
}
// ---- src/utils/stats.h: Stats::calculate -> rmse, mean, median, min, max, std, ninetynine
int ref_stats_calculate(int64_t n, const double *values, double *out7) {
  Stats st;
  st.values.assign(values, values + n);
  st.calculate();
  out7[0] = st.rmse, out7[1] = st.mean, out7[2] = st.median, out7[3] = st.min, out7[4] = st.max, out7[5] = st.std, out7[6] = st.ninetynine;
  return 0;
}

// ---- src/sim/Simulator.cpp + BsplineSE3.cpp driven as src/run_simulation.cpp:104-137 drives them
namespace {
struct SimHandle {
  std::shared_ptr<Simulator> sim;
  std::vector<double> imu, vic; // rows t w3 a3 / t q4 p3
};
} // namespace
// par13: seed, freq_imu, freq_cam, freq_vicon, sigma_w, sigma_wb, sigma_a, sigma_ab, gravity, 0, 0, 0, 0
void *ref_sim_run(const char *traj_path, const double *par13, const double *vicon_sigmas6) {
  SimulatorParams params;
  params.sim_traj_path = traj_path;
  params.seed = (int)par13[0];
  params.sim_freq_imu = par13[1];
  params.sim_freq_cam = par13[2];
  params.sim_freq_vicon = par13[3];
  params.sigma_w = par13[4];
  params.sigma_wb = par13[5];
  params.sigma_a = par13[6];
  params.sigma_ab = par13[7];
  params.gravity_magnitude = par13[8];
  params.sigma_vicon_pose << vicon_sigmas6[0], vicon_sigmas6[1], vicon_sigmas6[2], vicon_sigmas6[3], vicon_sigmas6[4], vicon_sigmas6[5];
  SimHandle *h = new SimHandle();
  FILE *old = stdout;
  h->sim = std::make_shared<Simulator>(params);
  (void)old;
  while (h->sim->ok()) {
    double time_imu;
    Eigen::Vector3d wm, am;
    if (h->sim->get_next_imu(time_imu, wm, am)) {
      h->imu.push_back(time_imu);
      for (int d = 0; d < 3; d++)
        h->imu.push_back(wm(d));
      for (int d = 0; d < 3; d++)
        h->imu.push_back(am(d));
    }
    double time_cam;
    h->sim->get_next_cam(time_cam);
    double time_vicon;
    Eigen::Vector4d q_VtoB;
    Eigen::Vector3d p_BinV;
    if (h->sim->get_next_vicon(time_vicon, q_VtoB, p_BinV)) {
      h->vic.push_back(time_vicon);
      for (int d = 0; d < 4; d++)
        h->vic.push_back(q_VtoB(d));
      for (int d = 0; d < 3; d++)
        h->vic.push_back(p_BinV(d));
    }
  }
  return h;
}
void ref_sim_destroy(void *hv) { delete (SimHandle *)hv; }
int64_t ref_sim_n_imu(void *hv) { return (int64_t)((SimHandle *)hv)->imu.size() / 7; }
int64_t ref_sim_n_vicon(void *hv) { return (int64_t)((SimHandle *)hv)->vic.size() / 8; }
// truth22 = R_BtoI9 | p_BinI3 | viconimu_dt | R_GtoV9 (the perturbed values the simulator drew)
int ref_sim_get(void *hv, double *imu7, double *vic8, double *truth22) {
  SimHandle *h = (SimHandle *)hv;
  std::memcpy(imu7, h->imu.data(), h->imu.size() * sizeof(double));
  std::memcpy(vic8, h->vic.data(), h->vic.size() * sizeof(double));
  SimulatorParams p = h->sim->get_params();
  put(p.R_BtoI, truth22);
  put(p.p_BinI, truth22 + 9);
  truth22[12] = p.viconimu_dt;
  put(p.R_GtoV, truth22 + 13);
  return 0;
}
int ref_sim_state_in_vicon(void *hv, int64_t n, const double *t, double *state17, int32_t *ok) {
  SimHandle *h = (SimHandle *)hv;
  for (int64_t i = 0; i < n; i++) {
    Eigen::Matrix<double, 17, 1> s;
    ok[i] = h->sim->get_state_in_vicon(t[i], s) ? 1 : 0;
    put(s, state17 + 17 * i);
  }
  return 0;
}

int64_t ref_solver_get_trace(int64_t capacity, double *rows5) {
  const std::vector<double> &t = gtsam::MiniLmLog::instance().trace;
  const int64_t n = (int64_t)t.size() / 5;
  for (int64_t i = 0; i < n && i < capacity; i++)
    for (int d = 0; d < 5; d++)
      rows5[5 * i + d] = t[5 * i + d];
  return n;
}
int ref_solver_write_to_file(void *hv, const char *csv, const char *info) {
  ((SolverHandle *)hv)->solver->write_to_file(csv, info);
  return 0;
}

} // extern "C"
