// ORACLE (test infrastructure, NOT product code) -- factor level pinned (oracle/_ref), optimiser level unpinned; see DESIGN.md section 2.
//
// C entry points over the CPU restatement, loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs ONLY.
// The product library (vicon2gt_b200/csrc) never links or calls any of this.
#include "solver.hpp"
#include <chrono>
#include <cstring>

using namespace orc;

namespace {
struct Handle {
  ViconGraphSolver *s = nullptr;
  LinearSystem ls;
  bool have_ls = false;
};
inline Vec3 v3(const double *p) { return Vec3{p[0], p[1], p[2]}; }
inline Mat3 m3(const double *p) {
  Mat3 m;
  for (int i = 0; i < 9; i++)
    m.a[i] = p[i];
  return m;
}
double now_s() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // namespace

extern "C" {

void orc_params_default(v2gt_params *p) {
  std::memset(p, 0, sizeof(*p));
  const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  std::memcpy(p->R_GtoV, I, sizeof(I));
  std::memcpy(p->R_BtoI, I, sizeof(I));
  p->gravity_magnitude = 9.81;
  p->estimate_toff_vicon_to_imu = p->estimate_ori_vicon_to_imu = p->estimate_pos_vicon_to_imu = 1;
  p->sigma_w = 1.6968e-04;
  p->sigma_wb = 1.9393e-05;
  p->sigma_a = 2.0000e-3;
  p->sigma_ab = 3.0000e-03;
  p->lm_max_iterations = 30;
  p->lm_lambda_initial = 1e-5;
  p->lm_lambda_factor = 10.0;
  p->lm_lambda_upper_bound = 1e20;
  p->lm_min_model_fidelity = 1e-3;
  p->lm_relative_error_tol = 1e-30;
  p->lm_absolute_error_tol = 1e-30;
  p->huber_k = 1.345;
  p->time_offset_window = 0.25;
  p->interp_gap_init = 5.0;
  p->interp_gap_factor = 0.1;
}

void *orc_create(const v2gt_params *params, int64_t n_imu, const double *imu_t, const double *imu_w, const double *imu_a,
                 int64_t n_vicon, const double *vic_t, const double *vic_q, const double *vic_p, const double *vic_Rq,
                 const double *vic_Rp, const double *vic_R_const, int64_t n_times, const double *state_t) {
  Propagator prop(params->sigma_w, params->sigma_wb, params->sigma_a, params->sigma_ab);
  prop.imu_data.reserve((size_t)n_imu);
  for (int64_t i = 0; i < n_imu; i++)
    prop.feed_imu(imu_t[i], v3(imu_w + 3 * i), v3(imu_a + 3 * i));
  Interpolator interp;
  interp.pose_data.reserve((size_t)n_vicon);
  for (int64_t i = 0; i < n_vicon; i++) {
    Vec4 q{vic_q[4 * i], vic_q[4 * i + 1], vic_q[4 * i + 2], vic_q[4 * i + 3]};
    Mat3 Rq = vic_Rq ? m3(vic_Rq + 9 * i) : m3(vic_R_const);
    Mat3 Rp = vic_Rp ? m3(vic_Rp + 9 * i) : m3(vic_R_const + 9);
    interp.feed_pose(vic_t[i], q, v3(vic_p + 3 * i), Rq, Rp);
  }
  std::vector<double> times(state_t, state_t + n_times);
  Handle *h = new Handle();
  h->s = new ViconGraphSolver(*params, prop, interp, times);
  return h;
}

void orc_destroy(void *hv) {
  Handle *h = (Handle *)hv;
  if (!h)
    return;
  delete h->s;
  delete h;
}

int orc_set_option(void *hv, const char *name, double v) {
  Handle *h = (Handle *)hv;
  std::string n(name);
  if (n == "faithful_scan")
    h->s->faithful_scan = v != 0;
  else if (n == "faithful_linerr")
    h->s->faithful_linerr = v != 0;
  else
    return V2GT_ERR_INVALID_ARG;
  return V2GT_OK;
}

int orc_build_and_solve(void *hv) {
  Handle *h = (Handle *)hv;
  ViconGraphSolver &s = *h->s;
  // same flow as ViconGraphSolver::build_and_solve but with phase timers
  s.info = v2gt_trial_info{};
  s.time_build = s.time_opt = 0;
  if (s.timestamp_cameras.empty())
    return s.info.status = V2GT_ERR_NO_TIMESTAMPS;
  if (s.interpolator.size() < 2)
    return s.info.status = V2GT_ERR_NO_VICON;
  double t0 = now_s();
  s.clean_timestamps();
  if (s.timestamp_cameras.empty())
    return s.info.status = V2GT_ERR_ALL_OUT_OF_RANGE;
  s.states.clear();
  for (int i = 0; i <= s.params.num_loop_relin; i++) {
    int st = s.build_problem(i == 0);
    double t1 = now_s();
    s.time_build += t1 - t0;
    if (st != V2GT_OK)
      return s.info.status = st;
    s.optimize_problem();
    t0 = now_s();
    s.time_opt += t0 - t1;
  }
  return s.info.status = V2GT_OK;
}

int orc_clean(void *hv) {
  Handle *h = (Handle *)hv;
  ViconGraphSolver &s = *h->s;
  s.info = v2gt_trial_info{};
  if (s.timestamp_cameras.empty())
    return s.info.status = V2GT_ERR_NO_TIMESTAMPS;
  if (s.interpolator.size() < 2)
    return s.info.status = V2GT_ERR_NO_VICON;
  s.clean_timestamps();
  if (s.timestamp_cameras.empty())
    return s.info.status = V2GT_ERR_ALL_OUT_OF_RANGE;
  s.states.clear();
  return V2GT_OK;
}
int orc_build(void *hv, int init_states) {
  Handle *h = (Handle *)hv;
  double t0 = now_s();
  int st = h->s->build_problem(init_states != 0);
  h->s->time_build += now_s() - t0;
  h->s->info.status = st;
  return st;
}
int orc_optimize(void *hv) {
  Handle *h = (Handle *)hv;
  double t0 = now_s();
  h->s->optimize_problem();
  h->s->time_opt += now_s() - t0;
  return V2GT_OK;
}
int orc_get_times(void *hv, double *build_s, double *opt_s) {
  Handle *h = (Handle *)hv;
  *build_s = h->s->time_build;
  *opt_s = h->s->time_opt;
  return V2GT_OK;
}
int orc_get_info(void *hv, v2gt_trial_info *info) {
  Handle *h = (Handle *)hv;
  *info = h->s->info;
  info->n_states = (int)h->s->states.size();
  return V2GT_OK;
}
int64_t orc_n_states(void *hv) { return (int64_t)((Handle *)hv)->s->timestamp_cameras.size(); }
int64_t orc_n_vicon(void *hv) { return (int64_t)((Handle *)hv)->s->interpolator.size(); }

int orc_get_states(void *hv, double *times, double *states) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  for (size_t i = 0; i < s.states.size(); i++) {
    if (times)
      times[i] = s.timestamp_cameras[i];
    if (states) {
      double *o = states + 16 * i;
      const NavState &x = s.states[i];
      for (int d = 0; d < 4; d++)
        o[d] = x.q(d);
      for (int d = 0; d < 3; d++) {
        o[4 + d] = x.bg(d);
        o[7 + d] = x.v(d);
        o[10 + d] = x.ba(d);
        o[13 + d] = x.p(d);
      }
    }
  }
  return V2GT_OK;
}
int orc_get_times_only(void *hv, double *times) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  for (size_t i = 0; i < s.timestamp_cameras.size(); i++)
    times[i] = s.timestamp_cameras[i];
  return V2GT_OK;
}
int orc_get_calibration(void *hv, double *q_BtoI, double *p_BinI, double *t_off, double *theta_xy) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  for (int d = 0; d < 4; d++)
    q_BtoI[d] = s.gl.q_BtoI(d);
  for (int d = 0; d < 3; d++)
    p_BinI[d] = s.gl.p_BinI(d);
  *t_off = s.gl.t_off;
  theta_xy[0] = s.gl.g.theta_x;
  theta_xy[1] = s.gl.g.theta_y;
  return V2GT_OK;
}
int orc_set_estimate(void *hv, int64_t n, const double *states, const double *globals) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  if ((size_t)n != s.states.size())
    return V2GT_ERR_INVALID_ARG;
  if (states)
    for (size_t i = 0; i < s.states.size(); i++) {
      const double *o = states + 16 * i;
      NavState &x = s.states[i];
      for (int d = 0; d < 4; d++)
        x.q(d) = o[d];
      for (int d = 0; d < 3; d++) {
        x.bg(d) = o[4 + d];
        x.v(d) = o[7 + d];
        x.ba(d) = o[10 + d];
        x.p(d) = o[13 + d];
      }
    }
  if (globals) {
    for (int d = 0; d < 4; d++)
      s.gl.q_BtoI(d) = globals[d];
    for (int d = 0; d < 3; d++)
      s.gl.p_BinI(d) = globals[4 + d];
    s.gl.t_off = globals[7];
    s.gl.g.theta_x = globals[8];
    s.gl.g.theta_y = globals[9];
  }
  return V2GT_OK;
}

int orc_eval_interp(void *hv, int64_t n, const double *t_query, double gap_thresh, int32_t *idx0, int32_t *idx1, int32_t *ok,
                    double *q, double *p, double *R, double *H_toff) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  for (int64_t i = 0; i < n; i++) {
    Interpolator::Result r = s.interpolator.interpolate(t_query[i], gap_thresh, true);
    if (idx0)
      idx0[i] = (int32_t)r.idx0;
    if (idx1)
      idx1[i] = (int32_t)r.idx1;
    if (ok)
      ok[i] = r.ok ? 1 : 0;
    if (q)
      for (int d = 0; d < 4; d++)
        q[4 * i + d] = r.q(d);
    if (p)
      for (int d = 0; d < 3; d++)
        p[3 * i + d] = r.p(d);
    if (R)
      for (int d = 0; d < 36; d++)
        R[36 * i + d] = r.R.a[d];
    if (H_toff)
      for (int d = 0; d < 6; d++)
        H_toff[6 * i + d] = r.H_toff(d);
  }
  return V2GT_OK;
}

static void pack_preint(const ImuFactorData &f, int n_steps, double *o) {
  std::memset(o, 0, sizeof(double) * V2GT_PREINT_DIM);
  o[0] = f.dt;
  for (int d = 0; d < 3; d++) {
    o[1 + d] = f.alpha(d);
    o[4 + d] = f.beta(d);
    o[11 + d] = f.bg_lin(d);
    o[14 + d] = f.ba_lin(d);
  }
  for (int d = 0; d < 4; d++)
    o[7 + d] = f.q_KtoK1(d);
  for (int d = 0; d < 9; d++) {
    o[17 + d] = f.J_q.a[d];
    o[26 + d] = f.J_beta.a[d];
    o[35 + d] = f.J_alpha.a[d];
    o[44 + d] = f.H_beta.a[d];
    o[53 + d] = f.H_alpha.a[d];
  }
  o[62] = n_steps;
  for (int d = 0; d < 225; d++)
    o[63 + d] = f.info.a[d];
}

int orc_get_preint(void *hv, double *out, double *cov) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  for (size_t i = 0; i < s.imu.size(); i++) {
    if (out)
      pack_preint(s.imu[i], 0, out + V2GT_PREINT_DIM * i);
    if (cov)
      for (int d = 0; d < 225; d++)
        cov[225 * i + d] = s.imu[i].P.a[d];
  }
  return V2GT_OK;
}

// Stand-alone preintegration of one interval: Propagator::propagate over the given store.
int orc_propagate(const double *sigmas, int64_t n_imu, const double *imu_t, const double *imu_w, const double *imu_a, double time0,
                  double time1, const double *bg_lin, const double *ba_lin, int faithful, double *out, double *cov, int32_t *n_steps) {
  Propagator prop(sigmas[0], sigmas[1], sigmas[2], sigmas[3]);
  for (int64_t i = 0; i < n_imu; i++)
    prop.feed_imu(imu_t[i], v3(imu_w + 3 * i), v3(imu_a + 3 * i));
  std::vector<ImuData> sel;
  if (!prop.select(time0, time1, sel, faithful != 0))
    return V2GT_ERR_INVALID_PREINT;
  CpiV1 pre(0, 0, 0, 0, true);
  if (!prop.propagate(time0, time1, v3(bg_lin), v3(ba_lin), pre, faithful != 0))
    return V2GT_ERR_INVALID_PREINT;
  ImuFactorData f;
  f.P = pre.P_meas;
  f.dt = pre.DT;
  f.alpha = pre.alpha_tau;
  f.beta = pre.beta_tau;
  f.q_KtoK1 = pre.q_k2tau;
  f.ba_lin = pre.b_a_lin;
  f.bg_lin = pre.b_w_lin;
  f.J_q = pre.J_q;
  f.J_beta = pre.J_b;
  f.J_alpha = pre.J_a;
  f.H_beta = pre.H_b;
  f.H_alpha = pre.H_a;
  make_imu_noise(f);
  if (out)
    pack_preint(f, (int)sel.size() - 1, out);
  if (cov)
    for (int d = 0; d < 225; d++)
      cov[d] = f.P.a[d];
  if (n_steps)
    *n_steps = (int)sel.size() - 1;
  return V2GT_OK;
}

int orc_eval_linearize(void *hv, double *D, double *O, double *B, double *g, double *Gamma, double *g_gamma, double *cost,
                       double *lin_error0) {
  Handle *h = (Handle *)hv;
  ViconGraphSolver &s = *h->s;
  s.linearize(s.states, s.gl, h->ls);
  h->have_ls = true;
  const size_t n = s.states.size();
  for (size_t i = 0; i < n; i++) {
    if (D)
      std::memcpy(D + 225 * i, h->ls.D[i].a, 225 * sizeof(double));
    if (O)
      std::memcpy(O + 225 * i, h->ls.O[i].a, 225 * sizeof(double));
    if (B)
      std::memcpy(B + 135 * i, h->ls.B[i].a, 135 * sizeof(double));
    if (g)
      std::memcpy(g + 15 * i, h->ls.g[i].a, 15 * sizeof(double));
  }
  if (Gamma)
    std::memcpy(Gamma, h->ls.Gamma.a, 81 * sizeof(double));
  if (g_gamma)
    std::memcpy(g_gamma, h->ls.gG.a, 9 * sizeof(double));
  if (cost)
    *cost = s.total_error(s.states, s.gl);
  if (lin_error0)
    *lin_error0 = h->ls.lin_error0;
  return V2GT_OK;
}

int orc_eval_solve(void *hv, double lambda, double *delta, int32_t *ok) {
  Handle *h = (Handle *)hv;
  if (!h->have_ls)
    return V2GT_ERR_NOT_BUILT;
  std::vector<double> d;
  bool good = h->s->solve(h->ls, lambda, d);
  if (ok)
    *ok = good ? 1 : 0;
  if (good && delta)
    std::memcpy(delta, d.data(), d.size() * sizeof(double));
  return V2GT_OK;
}

int orc_eval_cost(void *hv, double *cost) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  *cost = s.total_error(s.states, s.gl);
  return V2GT_OK;
}

// cost after retracting the current estimate by delta (the estimate itself is not changed)
int orc_eval_cost_at(void *hv, const double *delta, double *cost) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  std::vector<double> d(delta, delta + 15 * s.states.size() + 9);
  std::vector<NavState> xn;
  Globals gn;
  s.apply_delta(s.states, s.gl, d, xn, gn);
  *cost = s.total_error(xn, gn);
  return V2GT_OK;
}
// retract the current estimate in place
int orc_apply_delta(void *hv, const double *delta) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  std::vector<double> d(delta, delta + 15 * s.states.size() + 9);
  std::vector<NavState> xn;
  Globals gn;
  s.apply_delta(s.states, s.gl, d, xn, gn);
  s.states.swap(xn);
  s.gl = gn;
  return V2GT_OK;
}

int orc_get_lm_trace(void *hv, int64_t capacity, double *rows, int64_t *n_out) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  int64_t n = (int64_t)s.trace.size();
  if (n_out)
    *n_out = n;
  for (int64_t i = 0; i < n && i < capacity; i++) {
    rows[5 * i + 0] = s.trace[i].lambda;
    rows[5 * i + 1] = s.trace[i].cost_before;
    rows[5 * i + 2] = s.trace[i].cost_after;
    rows[5 * i + 3] = s.trace[i].lin_cost_change;
    rows[5 * i + 4] = s.trace[i].accepted;
  }
  return V2GT_OK;
}

int orc_write_to_file(void *hv, const char *csv, const char *info) {
  return ((Handle *)hv)->s->write_to_file(csv, info) ? V2GT_OK : V2GT_ERR_INVALID_ARG;
}

// ---- factor-level entry points (finite-difference and known-answer tests)
// x: [16] state rows, t: state time, globals: [10] (q_BtoI4 p_BinI3 t_off thetax thetay)
int orc_eval_imu_factor(const double *preint288, double grav_mag, const double *xi, const double *xj, const double *thetaxy, double *e15,
                        double *H1, double *H2, double *H3) {
  ImuFactorData f;
  f.dt = preint288[0];
  f.grav_mag = grav_mag;
  for (int d = 0; d < 3; d++) {
    f.alpha(d) = preint288[1 + d];
    f.beta(d) = preint288[4 + d];
    f.bg_lin(d) = preint288[11 + d];
    f.ba_lin(d) = preint288[14 + d];
  }
  for (int d = 0; d < 4; d++)
    f.q_KtoK1(d) = preint288[7 + d];
  for (int d = 0; d < 9; d++) {
    f.J_q.a[d] = preint288[17 + d];
    f.J_beta.a[d] = preint288[26 + d];
    f.J_alpha.a[d] = preint288[35 + d];
    f.H_beta.a[d] = preint288[44 + d];
    f.H_alpha.a[d] = preint288[53 + d];
  }
  auto st = [](const double *o) {
    NavState x;
    for (int d = 0; d < 4; d++)
      x.q(d) = o[d];
    for (int d = 0; d < 3; d++) {
      x.bg(d) = o[4 + d];
      x.v(d) = o[7 + d];
      x.ba(d) = o[10 + d];
      x.p(d) = o[13 + d];
    }
    return x;
  };
  RotXY g;
  g.theta_x = thetaxy[0];
  g.theta_y = thetaxy[1];
  ImuEval ev = eval_imu_factor(f, st(xi), st(xj), g, true);
  std::memcpy(e15, ev.e.a, 15 * sizeof(double));
  if (H1)
    std::memcpy(H1, ev.H1.a, 225 * sizeof(double));
  if (H2)
    std::memcpy(H2, ev.H2.a, 225 * sizeof(double));
  if (H3)
    std::memcpy(H3, ev.H3.a, 30 * sizeof(double));
  return V2GT_OK;
}

int orc_eval_vicon_factor(void *hv, double t, const double *x16, const double *globals10, double *e6, double *H1, double *H2,
                          double *H3, double *H4, int32_t *has_vicon) {
  ViconGraphSolver &s = *((Handle *)hv)->s;
  NavState x;
  x.t = t;
  for (int d = 0; d < 4; d++)
    x.q(d) = x16[d];
  for (int d = 0; d < 3; d++) {
    x.bg(d) = x16[4 + d];
    x.v(d) = x16[7 + d];
    x.ba(d) = x16[10 + d];
    x.p(d) = x16[13 + d];
  }
  Globals g;
  for (int d = 0; d < 4; d++)
    g.q_BtoI(d) = globals10[d];
  for (int d = 0; d < 3; d++)
    g.p_BinI(d) = globals10[4 + d];
  g.t_off = globals10[7];
  g.g.theta_x = globals10[8];
  g.g.theta_y = globals10[9];
  ViconEval ev = eval_vicon_factor(s.interpolator, s.cfg, x, g, true);
  std::memcpy(e6, ev.e.a, 6 * sizeof(double));
  if (H1)
    std::memcpy(H1, ev.H1.a, 90 * sizeof(double));
  if (H2)
    std::memcpy(H2, ev.H2.a, 18 * sizeof(double));
  if (H3)
    std::memcpy(H3, ev.H3.a, 18 * sizeof(double));
  if (H4)
    std::memcpy(H4, ev.H4.a, 6 * sizeof(double));
  if (has_vicon)
    *has_vicon = ev.has_vicon ? 1 : 0;
  return V2GT_OK;
}

// retract of a JPLNavState row / JPLQuaternion / RotationXY (known-answer tests)
int orc_retract_state(const double *x16, const double *xi15, double *out16) {
  NavState x;
  for (int d = 0; d < 4; d++)
    x.q(d) = x16[d];
  for (int d = 0; d < 3; d++) {
    x.bg(d) = x16[4 + d];
    x.v(d) = x16[7 + d];
    x.ba(d) = x16[10 + d];
    x.p(d) = x16[13 + d];
  }
  Vec15 xi;
  for (int d = 0; d < 15; d++)
    xi(d) = xi15[d];
  NavState y = retract(x, xi);
  for (int d = 0; d < 4; d++)
    out16[d] = y.q(d);
  for (int d = 0; d < 3; d++) {
    out16[4 + d] = y.bg(d);
    out16[7 + d] = y.v(d);
    out16[10 + d] = y.ba(d);
    out16[13 + d] = y.p(d);
  }
  return V2GT_OK;
}
int orc_retract_rotxy(const double *th, const double *d, double *out) {
  RotXY g;
  g.theta_x = th[0];
  g.theta_y = th[1];
  RotXY o = retract_rotxy(g, d[0], d[1]);
  out[0] = o.theta_x;
  out[1] = o.theta_y;
  return V2GT_OK;
}

// quaternion / SO(3) helpers exposed for cross-checks against the numpy restatement
void orc_rot_2_quat(const double *R, double *q) {
  Vec4 r = rot_2_quat(m3(R));
  for (int d = 0; d < 4; d++)
    q[d] = r(d);
}
void orc_quat_2_Rot(const double *q, double *R) {
  Mat3 m = quat_2_Rot(Vec4{q[0], q[1], q[2], q[3]});
  std::memcpy(R, m.a, 9 * sizeof(double));
}
void orc_quat_multiply(const double *q, const double *p, double *o) {
  Vec4 r = quat_multiply(Vec4{q[0], q[1], q[2], q[3]}, Vec4{p[0], p[1], p[2], p[3]});
  for (int d = 0; d < 4; d++)
    o[d] = r(d);
}
void orc_exp_so3(const double *w, double *R) {
  Mat3 m = exp_so3(v3(w));
  std::memcpy(R, m.a, 9 * sizeof(double));
}
void orc_log_so3(const double *R, double *w) {
  Vec3 r = log_so3(m3(R));
  for (int d = 0; d < 3; d++)
    w[d] = r(d);
}
void orc_Jl_so3(const double *w, double *J) {
  Mat3 m = Jl_so3(v3(w));
  std::memcpy(J, m.a, 9 * sizeof(double));
}
void orc_rot2rpy(const double *R, double *rpy) {
  Vec3 r = rot2rpy(m3(R));
  for (int d = 0; d < 3; d++)
    rpy[d] = r(d);
}

} // extern "C"
