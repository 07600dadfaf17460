// vicon2gt_b200 kernel (1): continuous preintegration (CpiV1) of every state interval of every trial.
// Replaces Propagator::propagate (src/meas/Propagator.cpp:33-163) + CpiV1::feed_IMU (src/cpi/CpiV1.h:58-341) +
// noiseModel::Gaussian::Covariance(P_meas) (src/gtsam/ImuFactorCPIv1.h:74).
//
// Sixteen lanes (half a warp) own one interval:
//   selection   lane 0 walks the interval's IMU samples (two binary searches, end-point interpolation, zero-dt filter:
//               Propagator.cpp:46-126) and lists them in shared memory, at most 16 steps at a time
//   means +     lane l turns step l (samples l, l+1) into one preintegrated increment (dR, dT, beta, alpha and the five
//   Jacobians   bias Jacobians of that step alone); an inclusive scan over the lanes with the ASSOCIATIVE composition
//               "A then B" of such increments gives every prefix -- the last one is the interval's
//               preintegration, the prefix rotations are what the covariance needs.  Longer intervals take several
//               rounds, the running composite being carried along
//   covariance  RK4 on Pdot = F P + P F^T + G Qc G^T is not a scan: it runs step by step, but the fifteen 3x3 blocks of
//               the symmetric P are advanced by fifteen lanes at once (stage inputs exchanged through shared memory)
//   information P = L L^T column by column across the lanes (rows), L^-1 one column per lane, P^-1 = L^-T L^-1
// The CTA's eight intervals are consecutive states of one trial: their outputs are staged in shared memory and written
// to the struct-of-arrays `pre` / `info` as 64-byte runs.
#include "v2gt_interp.cuh"
#include <stdio.h>

namespace v2 {

constexpr int PG = 16;           // lanes per interval
constexpr int PI_CTA = 8;        // intervals per CTA
constexpr int PNT = PG * PI_CTA; // threads per CTA
// per-interval shared memory (doubles)
constexpr int SM_SMP = 0;    // [17][8]  samples of the round: t, w[3], a[3], pad
constexpr int SM_STEP = 136; // [16][8]  per step: w_hat[3], a_hat[3], dt, pad
constexpr int SM_RPRE = 264; // [17][9]  rotation at the start of step k
constexpr int SM_PIN = 424;  // [2][136] RK4 stage inputs (15 blocks x 9), double buffered
constexpr int SM_CARRY = 696; // [64] composite of the rounds before this one (flat Elem)
constexpr int SM_PER = 760;
// epilogue aliases: dense P / L (225) and then the output staging (184) at SM_PIN; L^-1 (225) at SM_SMP
constexpr int OUT_N = PRE_STRIDE + INFO_STRIDE; // 64 + 120

// One preintegrated increment ("element"): what CpiV1::feed_IMU produces from the zero state for one or more steps.
struct Elem {
  double R[9]; // R_k2tau
  double T;
  double be[3], al[3];
  double Hb[9], Ha[9], Jq[9], Jb[9], Ja[9];
};
__device__ __forceinline__ void elem_identity(Elem &e) {
  eye3(e.R);
  e.T = 0.0;
#pragma unroll
  for (int k = 0; k < 3; k++)
    e.be[k] = e.al[k] = 0.0;
#pragma unroll
  for (int k = 0; k < 9; k++)
    e.Hb[k] = e.Ha[k] = e.Jq[k] = e.Jb[k] = e.Ja[k] = 0.0;
}

// The element of ONE step between two samples (CpiV1.h:58-244 applied to the zero state, imu_avg = true); also returns
// the bias-corrected averages the covariance needs.
__device__ __forceinline__ void elem_from_step(Elem &c, double *wh, double *ah, const double dt, const double *w0, const double *a0,
                                               const double *w1, const double *a1, const double *bw, const double *ba) {
  elem_identity(c);
#pragma unroll
  for (int k = 0; k < 3; k++) {
    wh[k] = 0.5 * ((w0[k] - bw[k]) + (w1[k] - bw[k]));
    ah[k] = .5 * ((a0[k] - ba[k]) + (a1[k] - ba[k]));
  }
  if (dt == 0)
    return;
  c.T = dt;
  const double *w = wh, *a = ah;
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
  {
    const double c1 = small_w ? dt : (sin_wt / mag_w);
    const double c2 = small_w ? (dt_2 / 2) : ((1.0 - cos_wt) / (mag_w * mag_w));
#pragma unroll
    for (int e = 0; e < 9; e++)
      c.R[e] = ((e % 4 == 0) ? 1.0 : 0.0) - c1 * w_x[e] + c2 * w_x_2[e];
  }
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
  double H_al[9], H_be[9]; // R_tau12k * arg
  mm3_tn(c.R, alpha_arg, H_al);
  mm3_tn(c.R, Beta_arg, H_be);
  mv3(H_al, a, c.al);
  mv3(H_be, a, c.be);
  // ---- bias Jacobians (CpiV1.h:150-244)
  {
    if (small_w) {
#pragma unroll
      for (int e = 0; e < 9; e++)
        c.Jq[e] = (((e % 4 == 0) ? 1.0 : 0.0) - .5 * (dt * w_x[e]) + (1.0 / 6.0) * (dt_2 * w_x_2[e])) * dt;
    } else {
      const double k1 = (1 - cos_wt) / (w_dt * w_dt), k2 = (w_dt - sin_wt) / (w_dt * w_dt * w_dt);
#pragma unroll
      for (int e = 0; e < 9; e++)
        c.Jq[e] = (((e % 4 == 0) ? 1.0 : 0.0) - k1 * (dt * w_x[e]) + k2 * (dt_2 * w_x_2[e])) * dt;
    }
  }
#pragma unroll
  for (int e = 0; e < 9; e++) {
    c.Ha[e] = -H_al[e];
    c.Hb[e] = -H_be[e];
  }
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
  for (int j = 0; j < 3; j++) {
    // d_R_bw_j = -R_tau12k [J_q e_j]x with the UPDATED J_q
    const double jq[3] = {c.Jq[j], c.Jq[3 + j], c.Jq[6 + j]};
    double S[9], dR[9];
    skew(jq, S);
    mm3_tn(c.R, S, dR);
#pragma unroll
    for (int e = 0; e < 9; e++)
      dR[e] = -dR[e];
    const double ej[3] = {j == 0 ? 1.0 : 0.0, j == 1 ? 1.0 : 0.0, j == 2 ? 1.0 : 0.0};
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
    mm3_tn(c.R, Ma, t1m);
    mm3(dR, Beta_arg, Tb);
    mm3_tn(c.R, Mb, t2m);
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
      c.Ja[3 * r + j] = ca[r];
      c.Jb[3 * r + j] = cb[r];
    }
  }
}

// flat layout of an Elem (shared-memory copy of the running composite)
constexpr int EO_R = 0, EO_T = 9, EO_BE = 10, EO_AL = 13, EO_HB = 16, EO_HA = 25, EO_JQ = 34, EO_JB = 43, EO_JA = 52;
__device__ __forceinline__ void elem_store(const Elem &e, double *f) {
#pragma unroll
  for (int k = 0; k < 9; k++) {
    f[EO_R + k] = e.R[k];
    f[EO_HB + k] = e.Hb[k];
    f[EO_HA + k] = e.Ha[k];
    f[EO_JQ + k] = e.Jq[k];
    f[EO_JB + k] = e.Jb[k];
    f[EO_JA + k] = e.Ja[k];
  }
  f[EO_T] = e.T;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    f[EO_BE + k] = e.be[k];
    f[EO_AL + k] = e.al[k];
  }
}

// B <- "A then B" (A covers the earlier samples), in place:
//   R = R_B R_A, T = T_A + T_B, beta = beta_A + R_A^T beta_B, alpha = alpha_A + beta_A T_B + R_A^T alpha_B,
//   H_b = H_bA + R_A^T H_bB, H_a = H_aA + H_bA T_B + R_A^T H_aB, J_q = R_B J_qA + J_qB,
//   J_b = J_bA + R_A^T ([beta_B]x J_qA + J_bB), J_a = J_aA + J_bA T_B + R_A^T ([alpha_B]x J_qA + J_aB)
// get(k, own) returns element k of A's flat layout; `own` is this lane's (still unmodified) value of the same element,
// which is what a shuffle sends.  A is fetched piece by piece so that only one Elem is ever live in registers; lanes
// with apply == false take part in the fetches and keep their B.
template <typename Get> __device__ __forceinline__ void elem_compose_inplace(Get get, Elem &B, const bool apply) {
  double AR[9], AJq[9], X[9], t9[9], u9[9];
#pragma unroll
  for (int e = 0; e < 9; e++) {
    AR[e] = get(EO_R + e, B.R[e]);
    AJq[e] = get(EO_JQ + e, B.Jq[e]);
  }
  const double BT = B.T;
  // J_b (needs the old beta_B, J_bB), J_a (old alpha_B, J_aB)
  skew(B.be, u9);
  mm3(u9, AJq, t9);
#pragma unroll
  for (int e = 0; e < 9; e++)
    t9[e] += B.Jb[e];
  mm3_tn(AR, t9, u9);
#pragma unroll
  for (int e = 0; e < 9; e++) {
    X[e] = get(EO_JB + e, B.Jb[e]); // J_bA
    const double v = X[e] + u9[e];
    B.Jb[e] = apply ? v : B.Jb[e];
  }
  skew(B.al, u9);
  mm3(u9, AJq, t9);
#pragma unroll
  for (int e = 0; e < 9; e++)
    t9[e] += B.Ja[e];
  mm3_tn(AR, t9, u9);
#pragma unroll
  for (int e = 0; e < 9; e++) {
    const double a = get(EO_JA + e, B.Ja[e]);
    const double v = (a + X[e] * BT) + u9[e];
    B.Ja[e] = apply ? v : B.Ja[e];
  }
  // J_q (old R_B)
  mm3(B.R, AJq, t9);
#pragma unroll
  for (int e = 0; e < 9; e++) {
    const double v = t9[e] + B.Jq[e];
    B.Jq[e] = apply ? v : B.Jq[e];
  }
  // H_b, H_a
  mm3_tn(AR, B.Hb, t9);
  mm3_tn(AR, B.Ha, u9);
#pragma unroll
  for (int e = 0; e < 9; e++) {
    X[e] = get(EO_HB + e, B.Hb[e]); // H_bA
    const double a = get(EO_HA + e, B.Ha[e]);
    const double vb = X[e] + t9[e];
    const double va = (a + X[e] * BT) + u9[e];
    B.Hb[e] = apply ? vb : B.Hb[e];
    B.Ha[e] = apply ? va : B.Ha[e];
  }
  // beta, alpha
  {
    double rb[3], ra[3];
#pragma unroll
    for (int r = 0; r < 3; r++) { // R_A^T v
      rb[r] = AR[r] * B.be[0] + AR[3 + r] * B.be[1] + AR[6 + r] * B.be[2];
      ra[r] = AR[r] * B.al[0] + AR[3 + r] * B.al[1] + AR[6 + r] * B.al[2];
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
      const double abe = get(EO_BE + r, B.be[r]);
      const double aal = get(EO_AL + r, B.al[r]);
      const double vb = abe + rb[r];
      const double va = (aal + abe * BT) + ra[r];
      B.be[r] = apply ? vb : B.be[r];
      B.al[r] = apply ? va : B.al[r];
    }
  }
  // T, R
  {
    const double at = get(EO_T, B.T);
    B.T = apply ? (at + BT) : B.T;
    mm3(B.R, AR, t9);
#pragma unroll
    for (int e = 0; e < 9; e++)
      B.R[e] = apply ? t9[e] : B.R[e];
  }
}

// upper-stored symmetric 5x5 block matrix over [th bg v ba p]: block (r, c), r <= c, has id 5 r - r (r - 1) / 2 + (c - r)
__device__ __forceinline__ int blk_id(int r, int c) { return 5 * r - (r * (r - 1)) / 2 + (c - r); }
// load block (j, y) of the symmetric matrix held in `buf` (15 blocks x 9)
__device__ __forceinline__ void ld_blk(const double *buf, int j, int y, double *o) {
  const bool tr = j > y;
  const double *b = buf + 9 * (tr ? blk_id(y, j) : blk_id(j, y));
  const int s1 = tr ? 1 : 3, s2 = tr ? 3 : 1;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int k = 0; k < 3; k++)
      o[3 * i + k] = b[i * s1 + k * s2];
}
// (F P)(x, y) for F of CpiV1.h:259-272: rows th: [A P_th,y - P_bg,y], v: [C P_th,y + E P_ba,y], p: [P_v,y], bg / ba: 0
__device__ __forceinline__ void fp_block(const double *buf, int x, int y, const double *A, const double *C, const double *E, double *o) {
  const bool u1 = (x == 0) || (x == 2), u2 = u1 || (x == 4);
  const int j2 = (x == 0) ? 1 : ((x == 2) ? 3 : 2);
  double M1[9], M2[9], P1[9], P2[9];
#pragma unroll
  for (int e = 0; e < 9; e++) {
    const double I = (e % 4 == 0) ? 1.0 : 0.0;
    M1[e] = u1 ? ((x == 0) ? A[e] : C[e]) : 0.0;
    M2[e] = u2 ? ((x == 0) ? -I : ((x == 2) ? E[e] : I)) : 0.0;
  }
  ld_blk(buf, 0, y, P1);
  ld_blk(buf, j2, y, P2);
  double t1[9], t2[9];
  mm3(M1, P1, t1);
  mm3(M2, P2, t2);
#pragma unroll
  for (int e = 0; e < 9; e++)
    o[e] = t1[e] + t2[e];
}

__device__ inline void set_status_pre(LmState *lm, int code) { atomicCAS(&lm->status, V2GT_OK, code); }

struct Smp {
  double t, w[3], a[3];
};

// grid: (ceil((n_max - 1) / PI_CTA), T), PNT threads, dynamic smem PI_CTA * SM_PER doubles (+ flags)
__global__ void __launch_bounds__(PNT, 3) k_preintegrate(DevPtrs d, int keep_cov) {
  extern __shared__ __align__(16) double smem[];
  __shared__ int s_ok[PI_CTA];
  const int t = blockIdx.y;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int g = threadIdx.x / PG, l = threadIdx.x % PG;
  const unsigned gmask = 0xFFFFu << (16 * ((threadIdx.x >> 4) & 1));
  const int i0 = blockIdx.x * PI_CTA;
  const int i = i0 + g;
  const bool trial_ok = (d.lm[t].status == V2GT_OK);
  const bool active = trial_ok && (i < n - 1);
  double *sm = smem + g * SM_PER;
  if (l == 0)
    s_ok[g] = 0;
  if (active) {
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

    // ---- selection state of lane 0 (Propagator.cpp:33-132 as a resumable stream; zero-dt filter of :118-126 on the fly)
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
    Smp pend = {}, d1 = {};
    long long kk = 0;
    int n_pushed = 0, n_final = 0, n_list = 0;
    bool sel_done = true, flushed = false;
    double DT = 0.0;
    double *smp = sm + SM_SMP;
    auto finalize = [&](const Smp &q) {
      double *o = smp + 8 * n_list;
      o[0] = q.t; o[1] = q.w[0]; o[2] = q.w[1]; o[3] = q.w[2]; o[4] = q.a[0]; o[5] = q.a[1]; o[6] = q.a[2];
      n_list++;
      n_final++;
    };
    auto push = [&](const Smp &q) {
      if (n_pushed > 0 && !(fabs(q.t - pend.t) < 1e-12))
        finalize(pend); // (the earlier entry of a zero-dt pair is erased)
      pend = q;
      n_pushed++;
    };
    if (l == 0 && M >= 2) {
      const long long lb = lower_bound_d(it, M, time0);
      kk = (lb >= 1) ? lb - 1 : 0;
      load(kk, d1);
      sel_done = !(kk + 1 < M);
    }
    double *carry = sm + SM_CARRY; // running composite of the earlier rounds (flat Elem)
    if (l == 0) {
      Elem id;
      elem_identity(id);
      elem_store(id, carry);
    }
    double p0[9], acc[9]; // this lane's block of P and of the RK4 accumulator (lanes 0..14)
#pragma unroll
    for (int e = 0; e < 9; e++)
      p0[e] = 0.0;
    // block (br, bc) of this lane
    int br = 0, bc = 0;
    {
      int id = l < 15 ? l : 14, r = 0;
      while (id >= 5 - r) {
        id -= 5 - r;
        r++;
      }
      br = r;
      bc = r + id;
    }
    const double qdiag = (br != bc) ? 0.0 : ((br == 0) ? qw : ((br == 1) ? qwb : ((br == 2) ? qa : ((br == 3) ? qab : 0.0))));
    double *pin = sm + SM_PIN;
    if (l < 15) {
#pragma unroll
      for (int e = 0; e < 9; e++)
        pin[9 * l + e] = 0.0;
    }
    int round = 0;
    bool more = true;
    while (more) {
      // ---- lane 0: the samples of this round (the last sample of the previous round first)
      int c_steps = 0;
      if (l == 0) {
        if (round > 0 && n_list > 0) {
#pragma unroll
          for (int k = 0; k < 8; k++)
            smp[k] = smp[8 * (n_list - 1) + k];
          n_list = 1;
        }
        while (!sel_done && n_list < 14) {
          const Smp d0 = d1;
          load(kk + 1, d1);
          Smp tmp;
          bool stop = false;
          if (d1.t > time0 && d0.t < time0) { // CASE 1
            lerp(d0, d1, time0, tmp);
            push(tmp);
          } else if (d0.t >= time0 && d1.t <= time1) { // CASE 2
            push(d0);
          } else if (d1.t > time1) { // CASE 3
            if (d0.t > time1 && kk == 0) {
              stop = true;
            } else {
              if (d0.t > time1) {
                Smp dm;
                load(kk - 1, dm);
                lerp(dm, d0, time1, tmp);
                push(tmp);
              } else {
                push(d0);
              }
              if (pend.t != time1) {
                lerp(d0, d1, time1, tmp);
                push(tmp);
              }
              stop = true;
            }
          }
          kk++;
          if (stop || !(kk + 1 < M))
            sel_done = true;
        }
        if (sel_done && !flushed) {
          if (n_pushed > 0)
            finalize(pend);
          flushed = true;
        }
        c_steps = n_list > 0 ? n_list - 1 : 0;
        for (int k = 0; k < c_steps; k++)
          DT += smp[8 * (k + 1)] - smp[8 * k]; // sequential, as DT += dt in feed_IMU
      }
      c_steps = __shfl_sync(gmask, c_steps, 0, PG);
      more = __shfl_sync(gmask, (l == 0 && !flushed) ? 1 : 0, 0, PG) != 0;
      __syncwarp(gmask);
      // ---- lane l: the increment of step l
      Elem e;
      elem_identity(e);
      if (l < c_steps) {
        const double *q0 = smp + 8 * l, *q1 = smp + 8 * (l + 1);
        double wh[3], ah[3];
        const double dt = q1[0] - q0[0];
        elem_from_step(e, wh, ah, dt, q0 + 1, q0 + 4, q1 + 1, q1 + 4, bw, ba);
        double *sd = sm + SM_STEP + 8 * l;
        sd[0] = wh[0]; sd[1] = wh[1]; sd[2] = wh[2]; sd[3] = ah[0]; sd[4] = ah[1]; sd[5] = ah[2]; sd[6] = dt;
      }
      // ---- inclusive scan over the lanes with the associative composition
      for (int dlt = 1; dlt < c_steps; dlt <<= 1)
        elem_compose_inplace([&](int, double own) { return __shfl_up_sync(gmask, own, dlt, PG); }, e, l >= dlt);
      if (round > 0) // the rounds before this one come first
        elem_compose_inplace([&](int k, double) { return carry[k]; }, e, true);
      // rotation at the start of every step of the round, then the running composite
      if (l == 0) {
#pragma unroll
        for (int k = 0; k < 9; k++)
          sm[SM_RPRE + k] = carry[EO_R + k];
      }
      if (l < c_steps) {
#pragma unroll
        for (int k = 0; k < 9; k++)
          sm[SM_RPRE + 9 * (l + 1) + k] = e.R[k];
      }
      __syncwarp(gmask); // every lane has read the old composite
      if (c_steps > 0 && l == c_steps - 1)
        elem_store(e, carry);
      __syncwarp(gmask);
      // ---- covariance: RK4 step by step (CpiV1.h:246-335), one 3x3 block of the symmetric P per lane
      for (int k = 0; k < c_steps; k++) {
        const double *sd = sm + SM_STEP + 8 * k;
        const double wv[3] = {sd[0], sd[1], sd[2]}, av[3] = {sd[3], sd[4], sd[5]};
        const double dt = sd[6];
        if (dt == 0)
          continue;
        double A[9], a_x[9], w_x[9], w_x_2[9];
        skew(wv, w_x);
        skew(av, a_x);
        mm3(w_x, w_x, w_x_2);
#pragma unroll
        for (int e9 = 0; e9 < 9; e9++)
          A[e9] = -w_x[e9];
        double Rk[9], Rk1[9], Rm[9];
#pragma unroll
        for (int e9 = 0; e9 < 9; e9++) {
          Rk[e9] = sm[SM_RPRE + 9 * k + e9];
          Rk1[e9] = sm[SM_RPRE + 9 * (k + 1) + e9];
        }
        {
          const double mag_w = norm3(wv);
          const bool small_w = (mag_w < 0.008726646);
          double s_h, c_h;
          sincos(mag_w * .5 * dt, &s_h, &c_h);
          const double hdt = .5 * dt;
          const double c1 = small_w ? hdt : (s_h / mag_w);
          const double c2 = small_w ? ((hdt * hdt) / 2) : ((1.0 - c_h) / (mag_w * mag_w));
          double Rh[9];
#pragma unroll
          for (int e9 = 0; e9 < 9; e9++)
            Rh[e9] = ((e9 % 4 == 0) ? 1.0 : 0.0) - c1 * w_x[e9] + c2 * w_x_2[e9];
          mm3(Rh, Rk, Rm);
        }
#pragma unroll 1
        for (int stage = 0; stage < 4; stage++) {
          const double *Rs = (stage == 0) ? Rk : ((stage == 3) ? Rk1 : Rm);
          const double *in = pin + 136 * (stage & 1);
          double *nx = pin + 136 * ((stage + 1) & 1);
          double K[9];
          if (l < 15) {
            double C[9], E[9], T[9];
            mm3_tn(Rs, a_x, T);
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
              for (int cc = 0; cc < 3; cc++) {
                C[3 * r + cc] = -T[3 * r + cc];
                E[3 * r + cc] = -Rs[3 * cc + r];
              }
            double X[9], Y[9];
            fp_block(in, br, bc, A, C, E, X);
            fp_block(in, bc, br, A, C, E, Y);
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
              for (int cc = 0; cc < 3; cc++)
                K[3 * r + cc] = X[3 * r + cc] + Y[3 * cc + r] + ((r == cc) ? qdiag : 0.0);
            double nv[9];
            if (stage == 0) {
#pragma unroll
              for (int e9 = 0; e9 < 9; e9++) {
                acc[e9] = K[e9];
                nv[e9] = p0[e9] + K[e9] * dt / 2.0;
              }
            } else if (stage == 1) {
#pragma unroll
              for (int e9 = 0; e9 < 9; e9++) {
                acc[e9] += 2.0 * K[e9];
                nv[e9] = p0[e9] + K[e9] * dt / 2.0;
              }
            } else if (stage == 2) {
#pragma unroll
              for (int e9 = 0; e9 < 9; e9++) {
                acc[e9] += 2.0 * K[e9];
                nv[e9] = p0[e9] + K[e9] * dt;
              }
            } else {
#pragma unroll
              for (int e9 = 0; e9 < 9; e9++) {
                acc[e9] += K[e9];
                p0[e9] += (dt / 6.0) * acc[e9];
              }
              if (br == bc) { // P = 1/2 (P + P^T), CpiV1.h:335
#pragma unroll
                for (int r = 0; r < 3; r++)
#pragma unroll
                  for (int cc = r + 1; cc < 3; cc++) {
                    const double m = 0.5 * (p0[3 * r + cc] + p0[3 * cc + r]);
                    p0[3 * r + cc] = m;
                    p0[3 * cc + r] = m;
                  }
              }
#pragma unroll
              for (int e9 = 0; e9 < 9; e9++)
                nv[e9] = p0[e9];
            }
#pragma unroll
            for (int e9 = 0; e9 < 9; e9++)
              nx[9 * l + e9] = nv[e9];
          }
          __syncwarp(gmask);
        }
      }
      round++;
    }
    // ---- validity (ViconGraphSolver.cpp:508-512: at least two samples and DT == time1 - time0, exactly)
    int ok = 1;
    if (l == 0 && (n_final < 2 || DT != (time1 - time0)))
      ok = 0;
    ok = __shfl_sync(gmask, ok, 0, PG);
    const double DTb = __shfl_sync(gmask, DT, 0, PG);
    const int n_steps = __shfl_sync(gmask, n_final, 0, PG) - 1;
    if (!ok) {
      if (l == 0)
        set_status_pre(d.lm + t, V2GT_ERR_INVALID_PREINT);
    } else {
      // ---- dense P (15x15) in shared memory
      double *Pd = sm + SM_PIN;
      __syncwarp(gmask);
      if (l < 15) {
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            Pd[15 * (3 * br + r) + 3 * bc + cc] = p0[3 * r + cc];
            Pd[15 * (3 * bc + cc) + 3 * br + r] = p0[3 * r + cc];
          }
      }
      __syncwarp(gmask);
      if (keep_cov && d.cov) {
        double *cv = d.cov + s * 225;
        for (int k = l; k < 225; k += PG)
          cv[k] = Pd[k];
        __syncwarp(gmask);
      }
      // ---- P = L L^T in place (lane r owns row r), column by column
      bool bad = false;
      for (int j = 0; j < 15; j++) {
        double sacc = 0.0;
        if (l >= j && l < 15) {
          sacc = Pd[15 * l + j];
          for (int k = 0; k < j; k++)
            sacc -= Pd[15 * l + k] * Pd[15 * j + k];
        }
        const double dd = __shfl_sync(gmask, sacc, j, PG);
        if (!(dd > 0.0)) {
          bad = true;
          break;
        }
        const double ljj = sqrt(dd);
        __syncwarp(gmask);
        if (l == j)
          Pd[15 * j + j] = ljj;
        else if (l > j && l < 15)
          Pd[15 * l + j] = sacc / ljj;
        __syncwarp(gmask);
      }
      if (bad) {
        if (l == 0)
          set_status_pre(d.lm + t, V2GT_ERR_NAN_COVARIANCE);
        ok = 0;
      } else {
        // ---- L^-1: lane c owns column c (forward substitution)
        double *Li = sm + SM_SMP; // dense 15x15, row-major
        if (l < 15) {
          for (int r = l; r < 15; r++) {
            double sacc = (r == l) ? 1.0 : 0.0;
            for (int k = l; k < r; k++)
              sacc -= Pd[15 * r + k] * Li[15 * k + l];
            Li[15 * r + l] = sacc / Pd[15 * r + r];
          }
        }
        __syncwarp(gmask);
        // ---- outputs into the staging area: pre (64) then the information matrix P^-1 = L^-T L^-1, packed lower (120)
        double *out = sm + SM_PIN;
        bool nan_found = false;
        for (int idx = l; idx < INFO_STRIDE; idx += PG) {
          int r = 0;
          while ((r + 1) * (r + 2) / 2 <= idx)
            r++;
          const int cc = idx - r * (r + 1) / 2;
          double sacc = 0;
          for (int k = r; k < 15; k++) // (Li^T Li)[r][cc] = sum_k Li[k][r] Li[k][cc], k >= max(r, cc) = r
            sacc += Li[15 * k + r] * Li[15 * k + cc];
          out[PRE_STRIDE + idx] = sacc;
          if (sacc != sacc)
            nan_found = true;
        }
        if (__any_sync(gmask, nan_found)) {
          if (l == 0)
            set_status_pre(d.lm + t, V2GT_ERR_NAN_COVARIANCE);
        }
        if (l == 0) {
          double q[4];
          rot_2_quat(carry + EO_R, q);
          out[PRE_DT] = DTb;
          for (int k = 0; k < 3; k++) {
            out[PRE_ALPHA + k] = carry[EO_AL + k];
            out[PRE_BETA + k] = carry[EO_BE + k];
            out[PRE_BGLIN + k] = bw[k];
            out[PRE_BALIN + k] = ba[k];
          }
          for (int k = 0; k < 4; k++)
            out[PRE_Q + k] = q[k];
          for (int e9 = 0; e9 < 9; e9++) {
            out[PRE_JQ + e9] = carry[EO_JQ + e9];
            out[PRE_JB + e9] = carry[EO_JB + e9];
            out[PRE_JA + e9] = carry[EO_JA + e9];
            out[PRE_HB + e9] = carry[EO_HB + e9];
            out[PRE_HA + e9] = carry[EO_HA + e9];
          }
          out[PRE_NSTEPS] = (double)n_steps;
          out[63] = 0.0;
          s_ok[g] = 1;
        }
      }
    }
  }
  __syncthreads();
  // ---- the CTA's outputs: 8 consecutive slots per element -> 64-byte runs
  if (trial_ok) {
    const long long s0 = tr.st_off + i0;
    for (int idx = threadIdx.x; idx < OUT_N * PI_CTA; idx += PNT) {
      const int e = idx / PI_CTA, gg = idx % PI_CTA;
      if (!s_ok[gg])
        continue;
      const double v = smem[gg * SM_PER + SM_PIN + e];
      if (e < PRE_STRIDE)
        d.pre[(long long)e * d.S + s0 + gg] = v;
      else
        d.info[(long long)(e - PRE_STRIDE) * d.S + s0 + gg] = v;
    }
  }
}

int launch_preintegrate(const DevPtrs &d, int n_max, int keep_cov, cudaStream_t st) {
  static bool attr_set = false;
  const size_t smem = sizeof(double) * SM_PER * PI_CTA;
  if (!attr_set) {
    V2_CUDA_CHECK(cudaFuncSetAttribute(k_preintegrate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  if (n_max <= 1)
    return V2GT_OK;
  dim3 grid((n_max - 1 + PI_CTA - 1) / PI_CTA, d.T);
  k_preintegrate<<<grid, PNT, smem, st>>>(d, keep_cov);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

} // namespace v2
