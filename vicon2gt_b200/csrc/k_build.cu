// vicon2gt_b200 kernels of the build stage (ViconGraphSolver::build_problem, .cpp:390-536):
//   k_init_states   initial guess + validity of every state (3 vicon lookups per state)
//   k_compact       erase invalid states (stream compaction per trial, order preserving)
//   k_preintegrate  kernel (1): IMU bracket search + continuous preintegration (CpiV1) of every
//                   interval + covariance -> information matrix
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
// kernel (1): continuous preintegration, one thread per interval; the 15x15 covariance lives in
// shared memory as fifteen 3x3 blocks per thread, element-major ([entry][thread]) so that a warp
// touches 32 consecutive words for every access.
constexpr int PRE_NT = 32;           // threads (intervals) per CTA
constexpr int PB = 15 * 9;           // doubles per covariance copy (15 blocks)
// block ids of the symmetric 5x5 block matrix over [th bg v ba p]
enum { B_TT = 0, B_TG, B_TV, B_TA, B_TP, B_GG, B_GV, B_GA, B_GP, B_VV, B_VA, B_VP, B_AA, B_AP, B_PP };

struct SmemCov {
  double *base; // points at [0][tid]
  int st;       // element stride: PRE_NT for the shared-memory copies ([entry][thread]), 1 for thread-local ones
  __device__ __forceinline__ double &at(int blk, int e) { return base[(blk * 9 + e) * st]; }
  __device__ __forceinline__ void ld(int blk, double *o) const {
#pragma unroll
    for (int e = 0; e < 9; e++)
      o[e] = base[(blk * 9 + e) * st];
  }
  __device__ __forceinline__ void ldT(int blk, double *o) const {
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        o[3 * r + c] = base[(blk * 9 + 3 * c + r) * st];
  }
};

// K = F P + P F^T + G Qc G^T in block form; R is the rotation of this RK4 stage.
//   A = -[w]x, C = -R^T [a]x, E = -R^T          (CpiV1.h:259-272)
// emit(blk, K9) is called once per block.
template <typename Emit>
__device__ __forceinline__ void cov_derivative(const SmemCov &P, const double *A, const double *C, const double *E, double qw,
                                               double qwb, double qa, double qab, Emit emit) {
  double K[9], X[9], Y[9], T[9], U[9];
  // ---- theta row
  {
    double Ptt[9], PtgT[9];
    P.ld(B_TT, Ptt);
    P.ldT(B_TG, PtgT);
    mm3(A, Ptt, X);
#pragma unroll
    for (int e = 0; e < 9; e++)
      X[e] -= PtgT[e];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        K[3 * r + c] = X[3 * r + c] + X[3 * c + r] + ((r == c) ? qw : 0.0);
    emit(B_TT, K);
    // K_tv = A Ptv - Pgv + Ptt C^T + Pta E^T
    double Ptv[9], Pgv[9], Pta[9];
    P.ld(B_TV, Ptv);
    P.ld(B_GV, Pgv);
    P.ld(B_TA, Pta);
    mm3(A, Ptv, X);
    mm3_nt(Ptt, C, Y);
    mm3_nt(Pta, E, T);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] - Pgv[e] + Y[e] + T[e];
    emit(B_TV, K);
    // K_ta = A Pta - Pga
    double Pga[9];
    P.ld(B_GA, Pga);
    mm3(A, Pta, X);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] - Pga[e];
    emit(B_TA, K);
    // K_tp = A Ptp - Pgp + Ptv
    double Ptp[9], Pgp[9];
    P.ld(B_TP, Ptp);
    P.ld(B_GP, Pgp);
    mm3(A, Ptp, X);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] - Pgp[e] + Ptv[e];
    emit(B_TP, K);
    // K_vp = C Ptp + E Pap + Pvv
    double Pap[9], Pvv[9];
    P.ld(B_AP, Pap);
    P.ld(B_VV, Pvv);
    mm3(C, Ptp, X);
    mm3(E, Pap, Y);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] + Y[e] + Pvv[e];
    emit(B_VP, K);
    // K_vv = Yv + Yv^T + qa I, Yv = C Ptv + E Pva^T
    double PvaT[9];
    P.ldT(B_VA, PvaT);
    mm3(C, Ptv, X);
    mm3(E, PvaT, Y);
#pragma unroll
    for (int e = 0; e < 9; e++)
      U[e] = X[e] + Y[e];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        K[3 * r + c] = U[3 * r + c] + U[3 * c + r] + ((r == c) ? qa : 0.0);
    emit(B_VV, K);
    // K_va = C Pta + E Paa
    double Paa[9];
    P.ld(B_AA, Paa);
    mm3(C, Pta, X);
    mm3(E, Paa, Y);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] + Y[e];
    emit(B_VA, K);
  }
  {
    // K_tg = A Ptg - Pgg ; K_gv = Pgt C^T + Pga E^T
    double Ptg[9], Pgg[9], Pga[9];
    P.ld(B_TG, Ptg);
    P.ld(B_GG, Pgg);
    P.ld(B_GA, Pga);
    mm3(A, Ptg, X);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] - Pgg[e];
    emit(B_TG, K);
    double Pgt[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        Pgt[3 * r + c] = Ptg[3 * c + r];
    mm3_nt(Pgt, C, X);
    mm3_nt(Pga, E, Y);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = X[e] + Y[e];
    emit(B_GV, K);
  }
  {
    // K_gg = qwb I ; K_ga = 0 ; K_gp = Pgv ; K_aa = qab I ; K_ap = Pva^T ; K_pp = Pvp + Pvp^T
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = (e % 4 == 0) ? qwb : 0.0;
    emit(B_GG, K);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = 0.0;
    emit(B_GA, K);
    P.ld(B_GV, K);
    emit(B_GP, K);
#pragma unroll
    for (int e = 0; e < 9; e++)
      K[e] = (e % 4 == 0) ? qab : 0.0;
    emit(B_AA, K);
    P.ldT(B_VA, K);
    emit(B_AP, K);
    double Pvp[9];
    P.ld(B_VP, Pvp);
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++)
        K[3 * r + c] = Pvp[3 * r + c] + Pvp[3 * c + r];
    emit(B_PP, K);
  }
}

// state of the running preintegration of one interval (registers)
struct CpiRegs {
  double DT;
  double alpha[3], beta[3];
  double R[9];
  double Jq[9], Ja[9], Jb[9], Ha[9], Hb[9];
};

// One CpiV1::feed_IMU step (CpiV1.h:58-341) with imu_avg = true.
__device__ inline void cpi_step(CpiRegs &c, double t0, double t1, const double *w0, const double *a0, const double *w1,
                                const double *a1, const double *bw, const double *ba, double qw, double qwb, double qa,
                                double qab, double *smem_thread) {
  const double dt = t1 - t0;
  c.DT += dt;
  if (dt == 0)
    return;
  double w[3], a[3];
#pragma unroll
  for (int k = 0; k < 3; k++) {
    w[k] = 0.5 * ((w0[k] - bw[k]) + (w1[k] - bw[k]));
    a[k] = .5 * ((a0[k] - ba[k]) + (a1[k] - ba[k]));
  }
  const double mag_w = norm3(w);
  const double w_dt = mag_w * dt;
  const bool small_w = (mag_w < 0.008726646);
  const double dt_2 = dt * dt;
  double sin_wt, cos_wt;
  sincos(w_dt, &sin_wt, &cos_wt);
  double w_x[9], a_x[9], w_x_2[9];
  skew(w, w_x);
  skew(a, a_x);
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
  // ---- covariance, RK4 (CpiV1.h:246-335) on the block form
  {
    double Rm[9];
    {
      double s_h, c_h;
      sincos(mag_w * .5 * dt, &s_h, &c_h);
      const double hdt = .5 * dt;
      const double c1 = small_w ? hdt : (s_h / mag_w);
      const double c2 = small_w ? ((hdt * hdt) / 2) : ((1.0 - c_h) / (mag_w * mag_w));
      double Rh[9];
#pragma unroll
      for (int e = 0; e < 9; e++)
        Rh[e] = ((e % 4 == 0) ? 1.0 : 0.0) - c1 * w_x[e] + c2 * w_x_2[e];
      mm3(Rh, c.R, Rm);
    }
    // P and the RK4 accumulator live in shared memory; the two stage inputs are thread-local (L1-resident local
    // memory), which keeps the shared-memory footprint at 2.1 kB per interval (3 CTAs per SM instead of 1)
    double psa[PB], psb[PB];
    SmemCov P0{smem_thread, PRE_NT}, Pacc{smem_thread + PB * PRE_NT, PRE_NT}, PsA{psa, 1}, PsB{psb, 1};
    double A[9], C[9], E[9];
#pragma unroll
    for (int e = 0; e < 9; e++)
      A[e] = -w_x[e];
    auto stage_mats = [&](const double *Rk) {
      double T[9];
      mm3_tn(Rk, a_x, T);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = 0; cc < 3; cc++) {
          C[3 * r + cc] = -T[3 * r + cc];
          E[3 * r + cc] = -Rk[3 * cc + r];
        }
    };
    const double h2 = dt / 2.0;
    // k1 : input P0 -> Pacc = k1 ; PsA = P0 + k1*dt/2
    stage_mats(c.R);
    cov_derivative(P0, A, C, E, qw, qwb, qa, qab, [&](int blk, const double *K) {
#pragma unroll
      for (int e = 0; e < 9; e++) {
        Pacc.at(blk, e) = K[e];
        PsA.at(blk, e) = P0.base[(blk * 9 + e) * PRE_NT] + K[e] * dt / 2.0;
      }
    });
    // k2 : input PsA -> Pacc += 2 k2 ; PsB = P0 + k2*dt/2
    stage_mats(Rm);
    cov_derivative(PsA, A, C, E, qw, qwb, qa, qab, [&](int blk, const double *K) {
#pragma unroll
      for (int e = 0; e < 9; e++) {
        Pacc.at(blk, e) += 2.0 * K[e];
        PsB.at(blk, e) = P0.base[(blk * 9 + e) * PRE_NT] + K[e] * dt / 2.0;
      }
    });
    // k3 : input PsB -> Pacc += 2 k3 ; PsA = P0 + k3*dt
    cov_derivative(PsB, A, C, E, qw, qwb, qa, qab, [&](int blk, const double *K) {
#pragma unroll
      for (int e = 0; e < 9; e++) {
        Pacc.at(blk, e) += 2.0 * K[e];
        PsA.at(blk, e) = P0.base[(blk * 9 + e) * PRE_NT] + K[e] * dt;
      }
    });
    // k4 : input PsA -> P0 += dt/6 (Pacc + k4)
    stage_mats(R1);
    (void)h2;
    cov_derivative(PsA, A, C, E, qw, qwb, qa, qab, [&](int blk, const double *K) {
#pragma unroll
      for (int e = 0; e < 9; e++)
        Pacc.at(blk, e) += K[e];
    });
#pragma unroll 1
    for (int i = 0; i < PB; i++)
      P0.base[i * PRE_NT] += (dt / 6.0) * Pacc.base[i * PRE_NT];
    // symmetrise the diagonal blocks (P = 1/2 (P + P^T), CpiV1.h:335)
    const int diag[5] = {B_TT, B_GG, B_VV, B_AA, B_PP};
#pragma unroll
    for (int b = 0; b < 5; b++) {
      const int blk = diag[b];
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int cc = r + 1; cc < 3; cc++) {
          const double m = 0.5 * (P0.at(blk, 3 * r + cc) + P0.at(blk, 3 * cc + r));
          P0.at(blk, 3 * r + cc) = m;
          P0.at(blk, 3 * cc + r) = m;
        }
    }
  }
#pragma unroll
  for (int e = 0; e < 9; e++)
    c.R[e] = R1[e];
}

// (block row, block col) -> block id and transposition for the upper-stored 5x5 block matrix
__device__ __forceinline__ double cov_entry(const SmemCov &P, int r, int c) {
  const int br = r / 3, bc = c / 3, er = r % 3, ec = c % 3;
  const int id[5][5] = {{B_TT, B_TG, B_TV, B_TA, B_TP},
                        {B_TG, B_GG, B_GV, B_GA, B_GP},
                        {B_TV, B_GV, B_VV, B_VA, B_VP},
                        {B_TA, B_GA, B_VA, B_AA, B_AP},
                        {B_TP, B_GP, B_VP, B_AP, B_PP}};
  if (br <= bc)
    return P.base[(id[br][bc] * 9 + 3 * er + ec) * PRE_NT];
  return P.base[(id[br][bc] * 9 + 3 * ec + er) * PRE_NT];
}

__device__ inline void set_status(LmState *lm, int code) { atomicCAS(&lm->status, V2GT_OK, code); }

// grid: (ceil(n_max / PRE_NT), T), PRE_NT threads, dynamic smem 2*PB*PRE_NT doubles.
__global__ void __launch_bounds__(PRE_NT) k_preintegrate(DevPtrs d, int keep_cov) {
  extern __shared__ double smem[];
  const int t = blockIdx.y;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int i = blockIdx.x * PRE_NT + threadIdx.x;
  if (d.lm[t].status != V2GT_OK || d.lm[t].frozen)
    return;
  const bool active = (i < n - 1);
  double *sm = smem + threadIdx.x;
  if (!active)
    return;
  const long long s = tr.st_off + i;
  const int cur = d.lm[t].cur;
  const double *x0 = d.x + ((long long)cur * d.S + s) * X_STRIDE;
  const double bw[3] = {x0[4], x0[5], x0[6]};
  const double ba[3] = {x0[10], x0[11], x0[12]};
  const double time0 = d.st_t[s], time1 = d.st_t[s + 1];
  const double qw = tr.prm.sigma_w * tr.prm.sigma_w, qwb = tr.prm.sigma_wb * tr.prm.sigma_wb;
  const double qa = tr.prm.sigma_a * tr.prm.sigma_a, qab = tr.prm.sigma_ab * tr.prm.sigma_ab;
  const double *it = d.imu_t + tr.imu_off;
  const double *is = d.imu_s + 8 * tr.imu_off;
  const long long M = tr.n_imu;

  for (int k = 0; k < 2 * PB; k++)
    sm[k * PRE_NT] = 0.0;
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

  // ---- sample selection (Propagator.cpp:33-132) as a stream; zero-dt filter of :118-126 applied on the fly
  struct Smp {
    double t, w[3], a[3];
  };
  auto load = [&](long long idx, Smp &o) {
    const double4 u0 = ldg4(is + 8 * idx);
    const double4 u1 = ldg4(is + 8 * idx + 4);
    o.t = u0.x; o.w[0] = u0.y; o.w[1] = u0.z; o.w[2] = u0.w;
    o.a[0] = u1.x; o.a[1] = u1.y; o.a[2] = u1.z;
  };
  auto lerp = [&](const Smp &s1, const Smp &s2, double ts, Smp &o) { // Propagator.h:63-73
    const double lambda = (ts - s1.t) / (s2.t - s1.t);
    o.t = ts;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      o.a[k] = (1 - lambda) * s1.a[k] + lambda * s2.a[k];
      o.w[k] = (1 - lambda) * s1.w[k] + lambda * s2.w[k];
    }
  };
  Smp pend = {}, fin = {}; // pend: last pushed sample (may still be dropped by the zero-dt filter); fin: last finalised one
  int n_pushed = 0, n_final = 0, n_steps = 0;
  auto finalize = [&](const Smp &smp) {
    if (n_final > 0) {
      cpi_step(c, fin.t, smp.t, fin.w, fin.a, smp.w, smp.a, bw, ba, qw, qwb, qa, qab, sm);
      n_steps++;
    }
    fin = smp;
    n_final++;
  };
  auto push = [&](const Smp &smp) {
    if (n_pushed > 0) {
      if (fabs(smp.t - pend.t) < 1e-12) {
        // the earlier entry of a zero-dt pair is erased
      } else {
        finalize(pend);
      }
    }
    pend = smp;
    n_pushed++;
  };
  if (M >= 2) {
    // time-ordered store: start at the sample in front of time0 (every earlier iteration of the reference's scan from
    // sample 0 matches none of its three cases); unordered store: the reference's scan itself (Propagator.cpp:46)
    long long lb = tr.imu_unsorted ? 0 : lower_bound_d(it, M, time0);
    long long ib = (lb >= 1) ? lb - 1 : 0;
    Smp d0, d1, tmp;
    load(ib, d1);
    for (long long k = ib; k + 1 < M; k++) {
      d0 = d1;
      load(k + 1, d1);
      if (d1.t > time0 && d0.t < time0) { // CASE 1
        lerp(d0, d1, time0, tmp);
        push(tmp);
        continue;
      }
      if (d0.t >= time0 && d1.t <= time1) { // CASE 2
        push(d0);
        continue;
      }
      if (d1.t > time1) { // CASE 3
        if (d0.t > time1 && k == 0) {
          break;
        } else if (d0.t > time1) {
          Smp dm;
          load(k - 1, dm);
          lerp(dm, d0, time1, tmp);
          push(tmp);
        } else {
          push(d0);
        }
        if (pend.t != time1) {
          lerp(d0, d1, time1, tmp);
          push(tmp);
        }
        break;
      }
    }
  }
  if (n_pushed > 0)
    finalize(pend);
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

  // ---- information matrix P^-1 = L^-T L^-1 (Gaussian::Covariance, ImuFactorCPIv1.h:74)
  SmemCov P0{sm, PRE_NT};
  double *Lm = sm + PB * PRE_NT;      // packed lower L  (120), in the accumulator's shared-memory region
  double Li[120];                     // packed lower L^-1, thread-local
  if (keep_cov && d.cov) {
    double *cv = d.cov + s * 225;
    for (int r = 0; r < 15; r++)
      for (int cc = 0; cc < 15; cc++)
        cv[15 * r + cc] = cov_entry(P0, r, cc);
  }
  bool bad = false;
  for (int j = 0; j < 15; j++) {
    double dd = cov_entry(P0, j, j);
    for (int k = 0; k < j; k++) {
      const double l = Lm[tri(j, k) * PRE_NT];
      dd -= l * l;
    }
    if (!(dd > 0.0)) {
      bad = true;
      break;
    }
    const double ljj = sqrt(dd);
    Lm[tri(j, j) * PRE_NT] = ljj;
    for (int r = j + 1; r < 15; r++) {
      double sacc = cov_entry(P0, r, j);
      for (int k = 0; k < j; k++)
        sacc -= Lm[tri(r, k) * PRE_NT] * Lm[tri(j, k) * PRE_NT];
      Lm[tri(r, j) * PRE_NT] = sacc / ljj;
    }
  }
  if (bad) {
    set_status(d.lm + t, V2GT_ERR_NAN_COVARIANCE);
    return;
  }
  for (int cc = 0; cc < 15; cc++) {
    for (int r = cc; r < 15; r++) {
      double sacc = (r == cc) ? 1.0 : 0.0;
      for (int k = cc; k < r; k++)
        sacc -= Lm[tri(r, k) * PRE_NT] * Li[tri(k, cc)];
      Li[tri(r, cc)] = sacc / Lm[tri(r, r) * PRE_NT];
    }
  }
  const SoaW inf = soa_w(d.info, s, INFO_STRIDE);
  bool nan_found = false;
  for (int r = 0; r < 15; r++)
    for (int cc = 0; cc <= r; cc++) {
      double sacc = 0;
      for (int k = r; k < 15; k++) // (Li^T Li)[r][cc] = sum_k Li[k][r] Li[k][cc], k >= max(r,cc) = r
        sacc += Li[tri(k, r)] * Li[tri(k, cc)];
      inf[tri(r, cc)] = sacc;
      if (sacc != sacc)
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
  const size_t smem = sizeof(double) * 2 * PB * PRE_NT;
  // function attributes are per device (context): set on every launch, a handle may live on any GPU of the process
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_preintegrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((n_max + PRE_NT - 1) / PRE_NT, d.T);
  if (grid.x == 0)
    return V2GT_OK;
  k_preintegrate<<<grid, PRE_NT, smem, st>>>(d, keep_cov);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

} // namespace v2
