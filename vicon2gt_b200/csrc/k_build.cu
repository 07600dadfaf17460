// vicon2gt_b200 kernels of the build stage (ViconGraphSolver::build_problem, .cpp:390-536):
//   k_init_states   initial guess + validity of every state (3 vicon lookups per state)
//   k_compact       erase invalid states (stream compaction per trial, order preserving)
//   k_preintegrate  kernel (1): IMU bracket search + continuous preintegration (CpiV1) of every
//                   interval: means and bias Jacobians
//   k_preint_cov    kernel (1'): the interval's covariance (RK4) -> information matrix
#include "v2gt_interp.cuh"
#include <stdio.h>

namespace v2 {

// ------------------------------------------------------------------------------------------------
// 6x6 general inverse by Gauss-Jordan with partial pivoting; only used for the reference's
// "isnan(R.norm()) || isnan(R.inverse().norm())" state filter (ViconGraphSolver.cpp:453).
__device__ bool cov_is_bad(const double *R36) {
  double M[36], I[36];
  double n2 = 0;
  for (int i = 0; i < 36; i++) {
    M[i] = R36[i];
    I[i] = (i % 7 == 0) ? 1.0 : 0.0;
    n2 += R36[i] * R36[i];
  }
  if (isnan(sqrt(n2)))
    return true;
  for (int c = 0; c < 6; c++) {
    int piv = c;
    double best = fabs(M[6 * c + c]);
    for (int r = c + 1; r < 6; r++)
      if (fabs(M[6 * r + c]) > best) {
        best = fabs(M[6 * r + c]);
        piv = r;
      }
    if (piv != c)
      for (int k = 0; k < 6; k++) {
        double t = M[6 * c + k];
        M[6 * c + k] = M[6 * piv + k];
        M[6 * piv + k] = t;
        t = I[6 * c + k];
        I[6 * c + k] = I[6 * piv + k];
        I[6 * piv + k] = t;
      }
    const double d = 1.0 / M[6 * c + c];
    for (int k = 0; k < 6; k++) {
      M[6 * c + k] *= d;
      I[6 * c + k] *= d;
    }
    for (int r = 0; r < 6; r++) {
      if (r == c)
        continue;
      const double f = M[6 * r + c];
      if (f == 0.0)
        continue;
      for (int k = 0; k < 6; k++) {
        M[6 * r + k] -= f * M[6 * c + k];
        I[6 * r + k] -= f * I[6 * c + k];
      }
    }
  }
  double m2 = 0;
  for (int i = 0; i < 36; i++)
    m2 += I[i] * I[i];
  return isnan(sqrt(m2));
}

// ------------------------------------------------------------------------------------------------
// grid: (chunks, T); one thread per live state slot.  ViconGraphSolver.cpp:424-483.
__global__ void __launch_bounds__(V2GT_CHUNK) k_init_states(DevPtrs d, int init_states) {
  const int t = blockIdx.y;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int i = blockIdx.x * V2GT_CHUNK + threadIdx.x;
  if (i >= n || d.lm[t].status != V2GT_OK || d.lm[t].frozen)
    return;
  const long long s = tr.st_off + i;
  const int cur = d.lm[t].cur;
  const double *gl = d.glob + ((long long)cur * d.T + t) * GLOB_STRIDE;
  const double t_inI = d.st_t[s];
  const double t_inV = t_inI - gl[GL_TOFF];
  const double TO = tr.prm.time_offset_window;
  const double *vt = d.vic_t + tr.vic_off;
  const double *vs = d.vic_s + 8 * tr.vic_off;
  const bool cc = tr.vic_cov_const != 0;
  const double *vc = cc ? tr.vic_Rconst : d.vic_c + d.vic_c_off[t];
  InterpResult r0, r2, r1;
  vicon_interp(vt, vs, vc, cc, tr.n_vic, t_inV - TO, tr.prm.interp_gap_init, false, r0);
  vicon_interp(vt, vs, vc, cc, tr.n_vic, t_inV + TO, tr.prm.interp_gap_init, false, r2);
  vicon_interp(vt, vs, vc, cc, tr.n_vic, t_inV, tr.prm.interp_gap_init, false, r1);
  bool ok = r0.ok && r2.ok && r1.ok;
  if (ok) {
    double R36[36];
    interp_dense_cov(r1, R36);
    if (cov_is_bad(R36))
      ok = false;
  }
  d.valid[s] = ok ? 1 : 0;
  if (ok && init_states) {
    // q_VtoI = q_BtoI0 (x) q_VtoB ; p_IinV = p_BinV - R(q_VtoB)^T R_BtoI0^T p_BinI0 ; v by 3-point midpoint
    double qB[4], RB[9];
    for (int k = 0; k < 9; k++)
      RB[k] = tr.prm.R_BtoI[k];
    rot_2_quat(RB, qB);
    double *x = d.x + ((long long)cur * d.S + s) * X_STRIDE;
    double q[4];
    quat_mul(qB, r1.q, q);
    double u[3]; // R_BtoI0^T p_BinI0
    mv3_t(RB, tr.prm.p_BinI, u);
    auto p_imu = [&](const InterpResult &r, double *pI) {
      double qi[4], Ri[9], w[3];
      quat_inv(r.q, qi);
      quat_2_Rot(qi, Ri);
      mv3(Ri, u, w);
      pI[0] = r.p[0] - w[0];
      pI[1] = r.p[1] - w[1];
      pI[2] = r.p[2] - w[2];
    };
    double pI[3], pI0[3], pI2[3];
    p_imu(r1, pI);
    p_imu(r0, pI0);
    p_imu(r2, pI2);
    x[0] = q[0]; x[1] = q[1]; x[2] = q[2]; x[3] = q[3];
    for (int k = 0; k < 3; k++) {
      x[4 + k] = 0.0;
      x[7 + k] = (pI2[k] - pI0[k]) / (2 * TO);
      x[10 + k] = 0.0;
      x[13 + k] = pI[k];
    }
  }
}

// ------------------------------------------------------------------------------------------------
// grid: T CTAs of 256 threads; order-preserving compaction of (time, state row) from the current
// half of x into the other half; flips lm.cur and updates n_states.
__global__ void __launch_bounds__(256) k_compact(DevPtrs d) {
  const int t = blockIdx.x;
  if (d.lm[t].frozen)
    return;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int cur = d.lm[t].cur;
  const double *src = d.x + ((long long)cur * d.S + tr.st_off) * X_STRIDE;
  double *dst = d.x + ((long long)(1 - cur) * d.S + tr.st_off) * X_STRIDE;
  __shared__ int s_warp[8];
  __shared__ int s_base;
  if (threadIdx.x == 0)
    s_base = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  for (int base = 0; base < n; base += 256) {
    const int i = base + threadIdx.x;
    const int keep = (i < n) ? d.valid[tr.st_off + i] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (lane == 0)
      s_warp[wid] = __popc(m);
    __syncthreads();
    int off = s_base;
    for (int w = 0; w < wid; w++)
      off += s_warp[w];
    off += __popc(m & ((1u << lane) - 1));
    if (keep) {
      d.st_t_tmp[tr.st_off + off] = d.st_t[tr.st_off + i];
      for (int k = 0; k < X_STRIDE; k++)
        dst[(long long)off * X_STRIDE + k] = src[(long long)i * X_STRIDE + k];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int tot = 0;
      for (int w = 0; w < 8; w++)
        tot += s_warp[w];
      s_base += tot;
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    d.removed_no_vicon[t] += n - s_base;
    d.n_states[t] = s_base;
    d.lm[t].cur = 1 - cur;
    // globals follow the state buffer
    const double *g0 = d.glob + ((long long)cur * d.T + t) * GLOB_STRIDE;
    double *g1 = d.glob + ((long long)(1 - cur) * d.T + t) * GLOB_STRIDE;
    for (int k = 0; k < GLOB_STRIDE; k++)
      g1[k] = g0[k];
    if (s_base == 0 && d.lm[t].status == V2GT_OK)
      d.lm[t].status = V2GT_ERR_ALL_OUT_OF_RANGE;
  }
}

// ------------------------------------------------------------------------------------------------
// kernel (1): continuous preintegration (CpiV1) of every interval, in two launches over the same sample stream:
//   k_preintegrate   one THREAD per interval: means and bias Jacobians (registers only)
//   k_preint_cov     sixteen LANES per interval (two intervals per warp): the 15x15 covariance by RK4, one column
//                    per lane in registers, and the information matrix P^-1 by Gauss-Jordan sweeps
// (one thread per interval for both needed four 15x15 copies per thread: 4.3 kB of shared / local memory per interval,
//  three warps per SM, 13 kB of DRAM traffic per interval for 1.7 kB of results)

// ---- sample selection of Propagator::propagate (Propagator.cpp:33-132) as a stream, the zero-dt filter of :118-126
// applied on the fly.  step(s0, s1) is called for every consecutive pair of the selected samples, at ONE call site, so
// that the lanes of a warp working on different intervals stay at the same instruction.  Returns the number of
// selected samples.
struct Smp {
  double t, w[3], a[3];
};
__device__ __forceinline__ void smp_load(const double *is, long long idx, Smp &o) {
  const double4 u0 = ldg4(is + 8 * idx);
  const double4 u1 = ldg4(is + 8 * idx + 4);
  o.t = u0.x; o.w[0] = u0.y; o.w[1] = u0.z; o.w[2] = u0.w;
  o.a[0] = u1.x; o.a[1] = u1.y; o.a[2] = u1.z;
}
// first raw sample of the scan for an interval starting at time0.  Time-ordered store: the sample in front of time0
// (every earlier iteration of the reference's scan from sample 0 matches none of its three cases); unordered store:
// the reference's scan itself (Propagator.cpp:46)
__device__ __forceinline__ long long scan_start(const double *__restrict__ it, long long M, int unsorted, double time0) {
  if (unsorted || M < 2)
    return 0;
  const long long lb = lower_bound_d(it, M, time0);
  return (lb >= 1) ? lb - 1 : 0;
}
// the same by the sixteen lanes of an interval together: seventeen-way sections instead of halvings (five rounds of one
// load for 10^5 samples instead of seventeen dependent ones)
__device__ __forceinline__ long long scan_start16(const double *__restrict__ it, long long M, int unsorted, double time0, unsigned hm,
                                                  int half, int j) {
  if (unsorted || M < 2)
    return 0;
  long long lo = 0, hi = M; // the lower bound lies in [lo, hi]
  while (lo < hi) {
    const long long span = hi - lo;
    const long long pj = lo + (span * (j + 1)) / 17; // < hi; non-decreasing in j
    const bool below = __ldg(it + pj) < time0;
    const int c = __popc((__ballot_sync(hm, below) >> (16 * half)) & 0xFFFFu); // lanes 0..c-1 are below (the store is sorted)
    const long long p_lo = lo + (span * c) / 17, p_hi = lo + (span * (c + 1)) / 17;
    if (c > 0)
      lo = p_lo + 1;
    if (c < 16)
      hi = p_hi;
  }
  return (lo >= 1) ? lo - 1 : 0;
}

// Propagator.h:63-73
__device__ __forceinline__ void smp_lerp(const double *is, long long i1, long long i2, double ts, Smp &o) {
  Smp s1, s2;
  smp_load(is, i1, s1);
  smp_load(is, i2, s2);
  const double lambda = (ts - s1.t) / (s2.t - s1.t);
  o.t = ts;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    o.a[k] = (1 - lambda) * s1.a[k] + lambda * s2.a[k];
    o.w[k] = (1 - lambda) * s1.w[k] + lambda * s2.w[k];
  }
}

template <class Step>
__device__ __forceinline__ int select_samples(const double *__restrict__ it, const double *__restrict__ is, long long M, long long k0,
                                              double time0, double time1, Step step) {
  Smp pend = {}, fin = {}; // pend: last pushed sample (may still be dropped by the zero-dt filter); fin: last finalised one
  int n_pushed = 0, n_final = 0;
  bool more = (M >= 2);
  long long k = k0; // scan_start()
  for (;;) {
    // what this iteration of the scan pushes: kind 0 nothing, 1 the sample interpolated at time0 (CASE 1), 2 sample k
    // (CASE 2 / 3), 3 the sample interpolated at time1 from k-1, k (CASE 3 with d0 beyond time1); `last`: CASE 3 was
    // reached, a second sample interpolated at time1 follows unless the first one already sits there
    int kind = 0;
    bool last = false;
    const long long kk = k;
    if (more && k + 1 < M) {
      const double t0 = __ldg(it + k), t1 = __ldg(it + k + 1);
      if (t1 > time0 && t0 < time0) {
        kind = 1;
      } else if (t0 >= time0 && t1 <= time1) {
        kind = 2;
      } else if (t1 > time1) {
        more = false;
        if (!(t0 > time1 && k == 0)) {
          kind = (t0 > time1) ? 3 : 2;
          last = true;
        }
      }
      k++;
    } else {
      more = false;
    }
#pragma unroll 1
    for (int u = 0; u < 3; u++) {
      // u = 0, 1: the pushes of this iteration; u = 2: the flush after the scan has ended
      bool do_push = false, do_fin = false;
      Smp s = {};
      if (u == 0) {
        do_push = (kind != 0);
        if (kind == 1)
          smp_lerp(is, kk, kk + 1, time0, s);
        else if (kind == 2)
          smp_load(is, kk, s);
        else if (kind == 3)
          smp_lerp(is, kk - 1, kk, time1, s);
      } else if (u == 1) {
        do_push = last && (pend.t != time1);
        if (do_push)
          smp_lerp(is, kk, kk + 1, time1, s);
      } else {
        do_fin = !more && n_pushed > 0;
      }
      if (do_push)
        do_fin = (n_pushed > 0) && !(fabs(s.t - pend.t) < 1e-12); // the earlier entry of a zero-dt pair is erased
      if (do_fin) {
        if (n_final > 0)
          step(fin, pend);
        fin = pend;
        n_final++;
      }
      if (do_push) {
        pend = s;
        n_pushed++;
      }
    }
    if (!more)
      break;
  }
  return n_final;
}

// averaged, bias-corrected measurement of one step (CpiV1.h:70-76, imu_avg = true)
__device__ __forceinline__ void step_rates(const Smp &s0, const Smp &s1, const double *bw, const double *ba, double *w, double *a) {
#pragma unroll
  for (int k = 0; k < 3; k++) {
    w[k] = 0.5 * ((s0.w[k] - bw[k]) + (s1.w[k] - bw[k]));
    a[k] = .5 * ((s0.a[k] - ba[k]) + (s1.a[k] - ba[k]));
  }
}
// exp(-[w]x tau) as the reference forms it (CpiV1.h:108-118 with tau = dt, :252-258 with tau = dt / 2)
__device__ __forceinline__ void step_rotation(const double *w_x, const double *w_x_2, double mag_w, bool small_w, double tau, double *Rd) {
  double s, c;
  sincos(mag_w * tau, &s, &c);
  const double c1 = small_w ? tau : (s / mag_w);
  const double c2 = small_w ? ((tau * tau) / 2) : ((1.0 - c) / (mag_w * mag_w));
#pragma unroll
  for (int e = 0; e < 9; e++)
    Rd[e] = ((e % 4 == 0) ? 1.0 : 0.0) - c1 * w_x[e] + c2 * w_x_2[e];
}

// state of the running preintegration of one interval (registers)
struct CpiRegs {
  double DT;
  double alpha[3], beta[3];
  double R[9];
  double Jq[9], Ja[9], Jb[9], Ha[9], Hb[9];
};

// One CpiV1::feed_IMU step (CpiV1.h:58-244, :337-341) with imu_avg = true: means and bias Jacobians.
__device__ __forceinline__ void cpi_step(CpiRegs &c, const Smp &s0, const Smp &s1, const double *bw, const double *ba) {
  const double dt = s1.t - s0.t;
  c.DT += dt;
  if (dt == 0)
    return;
  double w[3], a[3];
  step_rates(s0, s1, bw, ba, w, a);
  const double mag_w = norm3(w);
  const double w_dt = mag_w * dt;
  const bool small_w = (mag_w < 0.008726646);
  const double dt_2 = dt * dt;
  double sin_wt, cos_wt;
  sincos(w_dt, &sin_wt, &cos_wt);
  double w_x[9], w_x_2[9];
  skew(w, w_x);
  mm3(w_x, w_x, w_x_2);

  // ---- means (CpiV1.h:108-148)
  double Rd[9]; // R_tau2tau1
  {
    const double c1 = small_w ? dt : (sin_wt / mag_w);
    const double c2 = small_w ? (dt_2 / 2) : ((1.0 - cos_wt) / (mag_w * mag_w));
#pragma unroll
    for (int e = 0; e < 9; e++)
      Rd[e] = ((e % 4 == 0) ? 1.0 : 0.0) - c1 * w_x[e] + c2 * w_x_2[e];
  }
  double R1[9]; // R_k2tau1
  mm3(Rd, c.R, R1);
  double f_1, f_2, f_3, f_4;
  if (small_w) {
    f_1 = -(dt * dt * dt / 3);
    f_2 = (dt_2 * dt_2 / 8);
    f_3 = -(dt_2 / 2);
    f_4 = (dt * dt * dt / 6);
  } else {
    const double m2 = mag_w * mag_w, m3 = m2 * mag_w, m4 = m2 * m2;
    f_1 = (w_dt * cos_wt - sin_wt) / m3;
    f_2 = (w_dt * w_dt - 2 * cos_wt - 2 * w_dt * sin_wt + 2) / (2 * m4);
    f_3 = -(1 - cos_wt) / m2;
    f_4 = (w_dt - sin_wt) / m3;
  }
  double alpha_arg[9], Beta_arg[9];
#pragma unroll
  for (int e = 0; e < 9; e++) {
    const double I = (e % 4 == 0) ? 1.0 : 0.0;
    alpha_arg[e] = (dt_2 / 2.0) * I + f_1 * w_x[e] + f_2 * w_x_2[e];
    Beta_arg[e] = dt * I + f_3 * w_x[e] + f_4 * w_x_2[e];
  }
  double H_al[9], H_be[9]; // R_tau12k * arg  (R_tau12k = R1^T)
  mm3_tn(R1, alpha_arg, H_al);
  mm3_tn(R1, Beta_arg, H_be);
  {
    double va[3], vb[3];
    mv3(H_al, a, va);
    mv3(H_be, a, vb);
#pragma unroll
    for (int k = 0; k < 3; k++) {
      c.alpha[k] += c.beta[k] * dt + va[k];
      c.beta[k] += vb[k];
    }
  }
  // ---- bias Jacobians (CpiV1.h:150-244)
  {
    double Jr[9];
    if (small_w) {
#pragma unroll
      for (int e = 0; e < 9; e++) // w_tx = dt w_x ; w_tx^2 = dt^2 w_x^2
        Jr[e] = ((e % 4 == 0) ? 1.0 : 0.0) - .5 * (dt * w_x[e]) + (1.0 / 6.0) * (dt_2 * w_x_2[e]);
    } else {
      const double k1 = (1 - cos_wt) / (w_dt * w_dt), k2 = (w_dt - sin_wt) / (w_dt * w_dt * w_dt);
#pragma unroll
      for (int e = 0; e < 9; e++)
        Jr[e] = ((e % 4 == 0) ? 1.0 : 0.0) - k1 * (dt * w_x[e]) + k2 * (dt_2 * w_x_2[e]);
    }
    double T[9];
    mm3(Rd, c.Jq, T);
#pragma unroll
    for (int e = 0; e < 9; e++)
      c.Jq[e] = T[e] + Jr[e] * dt;
  }
#pragma unroll
  for (int e = 0; e < 9; e++) {
    c.Ha[e] = (c.Ha[e] - H_al[e]) + dt * c.Hb[e]; // uses the old H_b
    c.Hb[e] -= H_be[e];
  }
  {
    double df1, df2, df3, df4;
    if (small_w) {
      const double dt4 = dt_2 * dt_2;
      df1 = -(dt4 * dt / 15);
      df2 = (dt4 * dt_2 / 72);
      df3 = -(dt4 / 12);
      df4 = (dt4 * dt / 60);
    } else {
      const double m2 = mag_w * mag_w, m4 = m2 * m2, m5 = m4 * mag_w, m6 = m4 * m2;
      const double wd2 = w_dt * w_dt;
      df1 = (wd2 * sin_wt - 3 * sin_wt + 3 * w_dt * cos_wt) / m5;
      df2 = (wd2 - 4 * cos_wt - 4 * w_dt * sin_wt + wd2 * cos_wt + 4) / m6;
      df3 = (2 * (cos_wt - 1) + w_dt * sin_wt) / m4;
      df4 = (2 * w_dt + w_dt * cos_wt - 3 * sin_wt) / m5;
    }
#pragma unroll
    for (int e = 0; e < 9; e++)
      c.Ja[e] += c.Jb[e] * dt; // old J_b
#pragma unroll
    for (int j = 0; j < 3; j++) {
      // d_R_bw_j = -R_tau12k [J_q e_j]x with the UPDATED J_q
      double jq[3] = {c.Jq[j], c.Jq[3 + j], c.Jq[6 + j]};
      double S[9], dR[9];
      skew(jq, S);
      mm3_tn(R1, S, dR);
#pragma unroll
      for (int e = 0; e < 9; e++)
        dR[e] = -dR[e];
      double ej[3] = {j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
      double ex[9], sym[9], t1m[9], t2m[9];
      skew(ej, ex);
      mm3(ex, w_x, t1m);
      mm3(w_x, ex, t2m);
#pragma unroll
      for (int e = 0; e < 9; e++)
        sym[e] = t1m[e] + t2m[e];
      const double d1 = w[j] * df1, d2 = w[j] * df2, d3 = w[j] * df3, d4 = w[j] * df4;
      double Ma[9], Mb[9], Ta[9], Tb[9];
#pragma unroll
      for (int e = 0; e < 9; e++) {
        Ma[e] = d1 * w_x[e] - f_1 * ex[e] + d2 * w_x_2[e] - f_2 * sym[e];
        Mb[e] = d3 * w_x[e] - f_3 * ex[e] + d4 * w_x_2[e] - f_4 * sym[e];
      }
      mm3(dR, alpha_arg, Ta);
      mm3_tn(R1, Ma, t1m);
      mm3(dR, Beta_arg, Tb);
      mm3_tn(R1, Mb, t2m);
#pragma unroll
      for (int e = 0; e < 9; e++) {
        Ta[e] += t1m[e];
        Tb[e] += t2m[e];
      }
      double ca[3], cb[3];
      mv3(Ta, a, ca);
      mv3(Tb, a, cb);
#pragma unroll
      for (int r = 0; r < 3; r++) {
        c.Ja[3 * r + j] += ca[r];
        c.Jb[3 * r + j] += cb[r];
      }
    }
  }
#pragma unroll
  for (int e = 0; e < 9; e++)
    c.R[e] = R1[e];
}

__device__ inline void set_status(LmState *lm, int code) { atomicCAS(&lm->status, V2GT_OK, code); }

constexpr int PRE_NT = 128; // intervals (threads) per CTA of k_preintegrate

// grid: (ceil(n_max / PRE_NT), T), PRE_NT threads.
__global__ void __launch_bounds__(PRE_NT) k_preintegrate(DevPtrs d) {
  const int t = blockIdx.y;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int i = blockIdx.x * PRE_NT + threadIdx.x;
  if (d.lm[t].status != V2GT_OK || d.lm[t].frozen || i >= n - 1)
    return;
  const long long s = tr.st_off + i;
  const int cur = d.lm[t].cur;
  const double *x0 = d.x + ((long long)cur * d.S + s) * X_STRIDE;
  const double bw[3] = {x0[4], x0[5], x0[6]};
  const double ba[3] = {x0[10], x0[11], x0[12]};
  const double time0 = d.st_t[s], time1 = d.st_t[s + 1];

  CpiRegs c;
  c.DT = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    c.alpha[k] = 0;
    c.beta[k] = 0;
  }
  eye3(c.R);
#pragma unroll
  for (int e = 0; e < 9; e++) {
    c.Jq[e] = 0; c.Ja[e] = 0; c.Jb[e] = 0; c.Ha[e] = 0; c.Hb[e] = 0;
  }
  int n_steps = 0;
  const long long k0 = scan_start(d.imu_t + tr.imu_off, tr.n_imu, tr.imu_unsorted, time0);
  const int n_final = select_samples(d.imu_t + tr.imu_off, d.imu_s + 8 * tr.imu_off, tr.n_imu, k0, time0, time1,
                                     [&](const Smp &s0, const Smp &s1) {
                                       cpi_step(c, s0, s1, bw, ba);
                                       n_steps++;
                                     });
  if (n_final < 2 || c.DT != (time1 - time0)) { // ViconGraphSolver.cpp:508-512
    set_status(d.lm + t, V2GT_ERR_INVALID_PREINT);
    return;
  }

  // ---- outputs
  const SoaW pr = soa_w(d.pre, s, PRE_STRIDE);
  double q[4];
  rot_2_quat(c.R, q);
  pr[PRE_DT] = c.DT;
  for (int k = 0; k < 3; k++) {
    pr[PRE_ALPHA + k] = c.alpha[k];
    pr[PRE_BETA + k] = c.beta[k];
    pr[PRE_BGLIN + k] = bw[k];
    pr[PRE_BALIN + k] = ba[k];
  }
  for (int k = 0; k < 4; k++)
    pr[PRE_Q + k] = q[k];
  for (int e = 0; e < 9; e++) {
    pr[PRE_JQ + e] = c.Jq[e];
    pr[PRE_JB + e] = c.Jb[e];
    pr[PRE_JA + e] = c.Ja[e];
    pr[PRE_HB + e] = c.Hb[e];
    pr[PRE_HA + e] = c.Ha[e];
  }
  pr[PRE_NSTEPS] = (double)n_steps;
  pr[63] = 0.0;
}

// ---- covariance: P' = F P + P F^T + G Qc G^T (CpiV1.h:246-335) over the state order [th bg v ba p], with
//   F = [A -I 0 0 0; 0; C 0 0 E 0; 0; 0 0 I 0 0],  A = -[w]x, C = -R^T [a]x, E = -R^T  (R the rotation of the RK4 stage),
//   G Qc G^T = diag(qw, qwb, qa, qab, 0).
// P is symmetric, so P F^T = (F P)^T: lane j of an interval's sixteen holds COLUMN j of P, forms column j of M = F P
// from its own registers (rows th, v, p; the bias rows of F are zero), and gets row j of M -- one element from each
// other lane -- through a small transposition buffer in shared memory.  Both triangles are carried and stay bitwise
// symmetric (M_ij + M_ji is formed from the same two numbers on either side), which makes the reference's
// P = (P + P^T) / 2 after every step the identity.
constexpr int COV_WARPS = 4;            // warps per CTA, two intervals each
// doubles per column of the transposition buffer: slots 0..8 = the th, v, p rows of M in this column; slots 9..15 are
// constants, one per lane whose row of M is zero (bg, ba, the idle lane): slot 9 + n of column e holds the process
// noise q of that lane where e is its own column, zero elsewhere -- so the diagonal of Q arrives with the read of
// M[j][e] instead of costing a compare per element; one pad (odd stride: conflict-free)
constexpr int COV_TS = 17;
constexpr int COV_TBUF = 16 * COV_TS;
#ifndef COV_MIN_CTAS_N
#define COV_MIN_CTAS_N 3
#endif
constexpr int COV_MIN_CTAS = COV_MIN_CTAS_N;

// shared memory of one interval: two transposition buffers (alternating between the RK4 stages: one barrier per stage)
// and the rotations of the step: slots 0 and 2 alternate as start / end of a step, slot 1 is the half step
struct CovSmem {
  double tb[2][COV_TBUF];
  double R[3][9];
};

// sin and cos of a step's angle |w| tau: a few hundredths of a radian at IMU rates, where the Taylor polynomials are
// exact to the last bit or two; the library routine for anything larger
__device__ __forceinline__ void sincos_step(double x, double *s, double *c) {
  if (fabs(x) < 0.25) {
    const double x2 = x * x;
    double ps = 1.0 / 1307674368000.0; // 1 / 15!
    ps = fma(ps, x2, -1.0 / 6227020800.0);
    ps = fma(ps, x2, 1.0 / 39916800.0);
    ps = fma(ps, x2, -1.0 / 362880.0);
    ps = fma(ps, x2, 1.0 / 5040.0);
    ps = fma(ps, x2, -1.0 / 120.0);
    ps = fma(ps, x2, 1.0 / 6.0);
    *s = fma(-x * x2, ps, x);
    double pc = -1.0 / 87178291200.0; // 1 / 14!
    pc = fma(pc, x2, 1.0 / 479001600.0);
    pc = fma(pc, x2, -1.0 / 3628800.0);
    pc = fma(pc, x2, 1.0 / 40320.0);
    pc = fma(pc, x2, -1.0 / 720.0);
    pc = fma(pc, x2, 1.0 / 24.0);
    pc = fma(pc, x2, -0.5);
    *c = fma(x2, pc, 1.0);
  } else {
    sincos(x, s, c);
  }
}

// grid: (ceil(n_max / (2 COV_WARPS)), T), 32 COV_WARPS threads.  Runs after k_preintegrate, which has flagged the trials
// with an invalid interval.
__global__ void __launch_bounds__(32 * COV_WARPS, COV_MIN_CTAS) k_preint_cov(DevPtrs d, int keep_cov) {
  __shared__ CovSmem sm_all[COV_WARPS * 2];
  const int t = blockIdx.y;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int lane = threadIdx.x & 31, half = lane >> 4, j = lane & 15, wrp = threadIdx.x >> 5;
  const int i = (blockIdx.x * COV_WARPS + wrp) * 2 + half;
  if (d.lm[t].status != V2GT_OK || d.lm[t].frozen || i >= n - 1)
    return;
  const unsigned hm = 0xFFFFu << (16 * half); // the lanes of this interval: every barrier below is among them
  CovSmem &sm = sm_all[wrp * 2 + half];
  const long long s = tr.st_off + i;
  const int cur = d.lm[t].cur;
  const double *x0 = d.x + ((long long)cur * d.S + s) * X_STRIDE;
  const double time0 = d.st_t[s], time1 = d.st_t[s + 1];
  const int jb = j / 3; // block of this lane's column: 0 th, 1 bg, 2 v, 3 ba, 4 p, 5 = the idle sixteenth lane (a zero column)
  const int jc = j - 3 * jb;
  const double qd = (jb == 0)   ? tr.prm.sigma_w * tr.prm.sigma_w
                    : (jb == 1) ? tr.prm.sigma_wb * tr.prm.sigma_wb
                    : (jb == 2) ? tr.prm.sigma_a * tr.prm.sigma_a
                    : (jb == 3) ? tr.prm.sigma_ab * tr.prm.sigma_ab
                                : 0.0;
  // where row j of M sits in a column of the buffer: th 0..2, v 3..5, p 6..8; bg 9..11, ba 12..14, idle lane 15
  const int rs = (jb == 0) ? j : (jb == 2) ? j - 3 : (jb == 4) ? j - 6 : (jb == 1) ? j + 6 : (jb == 3) ? j + 3 : 15;
  const double qm = (rs < 9) ? qd : 0.0; // added to the stored M[j][j]; the zero-row lanes have q in their constant slot
#pragma unroll
  for (int r = 9; r < 16; r++) {
    sm.tb[0][j * COV_TS + r] = (r == rs) ? qd : 0.0;
    sm.tb[1][j * COV_TS + r] = (r == rs) ? qd : 0.0;
  }
  if (j < 9)
    sm.R[0][j] = (j % 4 == 0) ? 1.0 : 0.0;
  __syncwarp(hm);

  double P[15], DT = 0;
#pragma unroll
  for (int e = 0; e < 15; e++)
    P[e] = 0.0;
  int r_start = 0; // slot of the rotation at the start of the step

  const long long k0 = scan_start16(d.imu_t + tr.imu_off, tr.n_imu, tr.imu_unsorted, time0, hm, half, j);
  const int n_final = select_samples(
      d.imu_t + tr.imu_off, d.imu_s + 8 * tr.imu_off, tr.n_imu, k0, time0, time1, [&](const Smp &s0, const Smp &s1) {
        const double dt = s1.t - s0.t;
        DT += dt;
        if (dt == 0)
          return;
        double w[3], a[3];
        {
          const double bw[3] = {__ldg(x0 + 4), __ldg(x0 + 5), __ldg(x0 + 6)};
          const double ba[3] = {__ldg(x0 + 10), __ldg(x0 + 11), __ldg(x0 + 12)};
          step_rates(s0, s1, bw, ba, w, a);
        }
        const double *R0 = sm.R[r_start], *Rh = sm.R[1], *R1 = sm.R[2 - r_start];
        {
          // the rotations of the step, R(tau) = exp(-[w]x tau) R_k = R_k - c1 [w]x R_k + c2 [w]x^2 R_k at tau = dt / 2 and
          // dt (CpiV1.h:108-118, :252-258), one COLUMN per lane: [w]x r = w x r.  Lanes 0..2 store.
          const double mag_w = norm3(w);
          const bool small_w = (mag_w < 0.008726646);
          const double hdt = .5 * dt;
          double sf, cf, sh, ch;
          sincos_step(mag_w * dt, &sf, &cf);
          sincos_step(mag_w * hdt, &sh, &ch);
          const double im = 1.0 / mag_w, im2 = 1.0 / (mag_w * mag_w);
          const double c1f = small_w ? dt : sf * im, c2f = small_w ? (dt * dt) / 2 : (1.0 - cf) * im2;
          const double c1h = small_w ? hdt : sh * im, c2h = small_w ? (hdt * hdt) / 2 : (1.0 - ch) * im2;
          const double r[3] = {R0[jc], R0[3 + jc], R0[6 + jc]};
          const double x1[3] = {w[1] * r[2] - w[2] * r[1], w[2] * r[0] - w[0] * r[2], w[0] * r[1] - w[1] * r[0]};
          const double x2[3] = {w[1] * x1[2] - w[2] * x1[1], w[2] * x1[0] - w[0] * x1[2], w[0] * x1[1] - w[1] * x1[0]};
          if (j < 3) {
#pragma unroll
            for (int k = 0; k < 3; k++) {
              sm.R[1][3 * k + jc] = r[k] - c1h * x1[k] + c2h * x2[k];
              sm.R[2 - r_start][3 * k + jc] = r[k] - c1f * x1[k] + c2f * x2[k];
            }
          }
        }
        double Pacc[15], Ps[12]; // Ps: stage input; its p rows are never read (no column of F meets them)
        // one RK4 stage: K = F Pin + (F Pin)^T + Q, column j; emit(e, K_e) per element
        auto stage = [&](const double *Rk, const double *Pin, double *buf, auto emit) {
          double m[9];
          // th rows: A pth - pg
          m[0] = (w[2] * Pin[1] - w[1] * Pin[2]) - Pin[3];
          m[1] = (w[0] * Pin[2] - w[2] * Pin[0]) - Pin[4];
          m[2] = (w[1] * Pin[0] - w[0] * Pin[1]) - Pin[5];
          // v rows: C pth + E pa = -R^T (a x pth + pa)
          const double u0 = (a[1] * Pin[2] - a[2] * Pin[1]) + Pin[9];
          const double u1 = (a[2] * Pin[0] - a[0] * Pin[2]) + Pin[10];
          const double u2 = (a[0] * Pin[1] - a[1] * Pin[0]) + Pin[11];
#pragma unroll
          for (int r = 0; r < 3; r++)
            m[3 + r] = -(Rk[r] * u0 + Rk[3 + r] * u1 + Rk[6 + r] * u2);
          // p rows: pv
#pragma unroll
          for (int r = 0; r < 3; r++)
            m[6 + r] = Pin[6 + r];
#pragma unroll
          for (int r = 0; r < 9; r++)
            buf[j * COV_TS + r] = m[r];
          buf[j * COV_TS + rs] += qm; // M[j][j] + q, read back by this lane alone
          __syncwarp(hm);
#pragma unroll
          for (int e = 0; e < 15; e++) {
            const double mt = buf[e * COV_TS + rs]; // M[j][e] (+ q on the diagonal)
            const int blk = e / 3;
            const double K = (blk == 0) ? m[e] + mt : (blk == 2) ? m[e - 3] + mt : (blk == 4) ? m[e - 6] + mt : mt;
            emit(e, K);
          }
        };
        // k1 at R_k (CpiV1.h:274-281), k2 and k3 at the half-step rotation (:283-311), k4 at R_tau1 (:313-327); the
        // barrier of stage 1 also orders the stores of the two new rotations before their first use
        const double h2 = dt / 2.0, h6 = dt / 6.0;
        stage(R0, P, sm.tb[0], [&](int e, double K) {
          Pacc[e] = K;
          if (e < 12)
            Ps[e] = P[e] + K * h2;
        });
        stage(Rh, Ps, sm.tb[1], [&](int e, double K) {
          Pacc[e] += 2.0 * K;
          if (e < 12)
            Ps[e] = P[e] + K * h2;
        });
        stage(Rh, Ps, sm.tb[0], [&](int e, double K) {
          Pacc[e] += 2.0 * K;
          if (e < 12)
            Ps[e] = P[e] + K * dt;
        });
        stage(R1, Ps, sm.tb[1], [&](int e, double K) { P[e] += h6 * (Pacc[e] + K); });
        r_start = 2 - r_start;
      });
  if (n_final < 2 || DT != (time1 - time0)) // flagged by k_preintegrate
    return;

  if (keep_cov && d.cov && j < 15) {
    double *cv = d.cov + s * 225;
#pragma unroll
    for (int r = 0; r < 15; r++)
      cv[15 * r + j] = P[r];
  }
  // ---- information matrix P^-1 (Gaussian::Covariance, ImuFactorCPIv1.h:74) by Gauss-Jordan sweeps in place, without
  // pivoting (P is symmetric positive definite; the pivots are those of its Cholesky factorisation, squared): sweep k
  // broadcasts column k through shared memory, every lane updates its own column
  bool bad = false;
#pragma unroll
  for (int k = 0; k < 15; k++) {
    double *buf = sm.tb[k & 1];
    if (j == k) {
#pragma unroll
      for (int r = 0; r < 15; r++)
        buf[r] = P[r];
    }
    __syncwarp(hm);
    const double pk = buf[k];
    if (!(pk > 0.0))
      bad = true; // same value on the sixteen lanes
    const double dinv = 1.0 / pk;
    const double f = (j == k) ? dinv : P[k] * dinv; // row k of the swept matrix
    // column k of the swept matrix is -c / p_kk: lane k restarts from zero (one predicated move per element; the
    // compiler's own lowering of "if (j == k) P[r] = 0" is a zero and two predicated copies)
#pragma unroll
    for (int r = 0; r < 15; r++)
      if (r != k)
        asm("{\n\t.reg .pred p;\n\tsetp.eq.s32 p, %1, %2;\n\t@p mov.f64 %0, 0d0000000000000000;\n\t}" : "+d"(P[r]) : "r"(j), "r"(k));
#pragma unroll
    for (int r = 0; r < 15; r++)
      if (r != k)
        P[r] = fma(-buf[r], f, P[r]);
    P[k] = f;
  }
  if (bad) {
    if (j == 0)
      set_status(d.lm + t, V2GT_ERR_NAN_COVARIANCE);
    return;
  }
  const SoaW inf = soa_w(d.info, s, INFO_STRIDE);
  bool nan_found = false;
#pragma unroll
  for (int r = 0; r < 15; r++)
    if (r >= j) {
      inf[tri(r, j)] = P[r];
      if (P[r] != P[r])
        nan_found = true;
    }
  if (nan_found)
    set_status(d.lm + t, V2GT_ERR_NAN_COVARIANCE);
}

// ------------------------------------------------------------------------------------------------
// host launchers
int launch_init_states(const DevPtrs &d, int n_max, int init_states, cudaStream_t st) {
  dim3 grid((n_max + V2GT_CHUNK - 1) / V2GT_CHUNK, d.T);
  if (grid.x == 0)
    return V2GT_OK;
  k_init_states<<<grid, V2GT_CHUNK, 0, st>>>(d, init_states);
  k_compact<<<d.T, 256, 0, st>>>(d);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

int launch_preintegrate(const DevPtrs &d, int n_max, int keep_cov, cudaStream_t st) {
  if (n_max <= 0)
    return V2GT_OK;
  k_preintegrate<<<dim3((n_max + PRE_NT - 1) / PRE_NT, d.T), PRE_NT, 0, st>>>(d);
  k_preint_cov<<<dim3((n_max + 2 * COV_WARPS - 1) / (2 * COV_WARPS), d.T), 32 * COV_WARPS, 0, st>>>(d, keep_cov);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

} // namespace v2
