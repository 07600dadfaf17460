// vicon2gt_b200 kernels (3) and (4a): factor evaluation and normal-equation assembly.
//
//   k_factor_eval<JAC>  one thread per state: vicon factor of the state + IMU factor to the next
//                       state; cost partials (always) and Jacobian pieces (JAC)
//   k_assemble          five threads per state (one per 3-column block of the state): the
//                       state's 15x15 diagonal block, its 15x15 coupling to the next state, its
//                       15x9 border and gradient, straight from the 3x3 block structure of the
//                       Jacobians (H1 has 13 non-zero 3x3 blocks, H2 five) and the information
//                       matrix; each state owns its blocks, no atomics
//   k_reduce_gamma      fixed-order reduction of the 9x9 global block / gradient / linear error
// Replaces GTSAM's NonlinearFactorGraph::linearize + error for the graph built in
// ViconGraphSolver::build_problem (SURVEY Appendix B gives the block structure).
#include "v2gt_factors.cuh"
#include <stdio.h>

namespace v2 {

// JPLNavState::retract (src/gtsam/JPLNavState.cpp:25-54) of one state row: q <- dq(dtheta) (x) q, the rest adds
__device__ __forceinline__ void retract_state(const double *x0, const double *dx, double *x1) {
  double dq[4], q[4];
  small_quat(dx, dq);
  quat_mul(dq, x0, q);
  x1[0] = q[0]; x1[1] = q[1]; x1[2] = q[2]; x1[3] = q[3];
#pragma unroll
  for (int k = 0; k < 12; k++)
    x1[4 + k] = x0[4 + k] + dx[3 + k];
}
__device__ __forceinline__ void ld16(const double *p, double *o) {
  const double4 *p4 = reinterpret_cast<const double4 *>(p);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const double4 v = p4[k];
    o[4 * k] = v.x; o[4 * k + 1] = v.y; o[4 * k + 2] = v.z; o[4 * k + 3] = v.w;
  }
}

// ------------------------------------------------------------------------------------------------
// grid (nchunk, T), V2GT_CHUNK threads.  which = 0: evaluate the current estimate; 1: the trial point
// x (+) delta, which this kernel forms itself (Values::retract fused into the error evaluation: the state
// rows of the trial point are written to the other half of the double buffer) together with the two dot
// products g.delta and |delta|^2 of the LM model-fidelity test.
template <bool JAC>
#ifndef FE_MIN_CTAS_COST
#define FE_MIN_CTAS_COST 3
#endif
#ifndef FE_MIN_CTAS_JAC
#define FE_MIN_CTAS_JAC 2
#endif
__global__ void __launch_bounds__(V2GT_CHUNK, JAC ? FE_MIN_CTAS_JAC : FE_MIN_CTAS_COST) k_factor_eval(DevPtrs d, int which, int force_all) {
  const int t = blockIdx.y;
  const LmState &lm = d.lm[t];
  __shared__ double s_red[4][V2GT_CHUNK / 32];
  // trial-level predicates are uniform across the CTA
  if (!force_all) {
    if (lm.lm_state != 0 || lm.status != V2GT_OK)
      return;
    if (JAC && !lm.need_lin)
      return;
    if (!JAC && lm.solve_failed)
      return;
  }
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  const int i = blockIdx.x * V2GT_CHUNK + threadIdx.x;
  const int buf = which ? (1 - lm.cur) : lm.cur;
  double cost = 0.0, lin0 = 0.0, gd = 0.0, dd = 0.0;
  // chain-sharded mode: halo slots carry the neighbours' boundary states; their vicon factors and the IMU factor
  // that starts at them belong to the neighbour (they are evaluated where needed, but not counted here)
  const bool own = !d.chain || (i >= d.c_lo && i < d.c_hi);
  if (i < n) {
    const long long s = tr.st_off + i;
    const double *gl = d.glob + ((long long)buf * d.T + t) * GLOB_STRIDE;
    const double *xb = d.x + (long long)buf * d.S * X_STRIDE;
    const double *xcur = d.x + (long long)lm.cur * d.S * X_STRIDE;
    double xi[16];
    if (!JAC && which) {
      double x0[16], dl[16];
      ld16(xcur + s * X_STRIDE, x0);
      ld16(d.delta + s * 16, dl);
      retract_state(x0, dl, xi);
      double4 *o4 = reinterpret_cast<double4 *>(d.x + ((long long)buf * d.S + s) * X_STRIDE);
#pragma unroll
      for (int k = 0; k < 4; k++)
        o4[k] = make_double4(xi[4 * k], xi[4 * k + 1], xi[4 * k + 2], xi[4 * k + 3]);
      if (own) {
        double gk[16];
        ld16(d.grad + s * 16, gk);
#pragma unroll
        for (int k = 0; k < 15; k++) {
          gd += gk[k] * dl[k];
          dd += dl[k] * dl[k];
        }
      }
    } else {
      ld16(xb + s * X_STRIDE, xi);
    }
    // ---- vicon factor
    {
      ViconPieces vp;
      const bool cc = tr.vic_cov_const != 0;
      const double *vc = cc ? tr.vic_Rconst : d.vic_c + d.vic_c_off[t];
      // a halo slot queries a time no pose can bracket, which makes the factor empty (has = 0, zero blocks)
      // d.valid (scratch flags of the build stage) holds every state's last bracket between evaluations
      vicon_factor<JAC>(tr, d.vic_t + tr.vic_off, d.vic_s + 8 * tr.vic_off, vc, cc, own ? d.st_t[s] : -1.0e300, xi, gl, vp, d.valid + s);
      cost += huber_loss(tr.prm.huber_k, vp.dist);
      double r2 = 0;
#pragma unroll
      for (int k = 0; k < 6; k++)
        r2 += vp.r[k] * vp.r[k];
      lin0 += 0.5 * r2;
      if (JAC) {
        const SoaW o = soa_w(d.fpv, s, FPV_STRIDE);
#pragma unroll
        for (int k = 0; k < 6; k++)
          o[FPV_R + k] = vp.r[k];
#pragma unroll
        for (int k = 0; k < 36; k++)
          o[FPV_A + k] = vp.A[k];
#pragma unroll
        for (int k = 0; k < 42; k++)
          o[FPV_G + k] = vp.G[k];
        o[FPV_HAS] = vp.has;
        o[FPV_IDX0] = vp.idx0;
        o[FPV_IDX1] = vp.idx1;
        o[FPV_W] = vp.w;
      }
    }
    // ---- IMU factor i -> i+1
    if (i < n - 1) {
      double xj[16];
      if (!JAC && which) {
        double x0[16], dl[16];
        ld16(xcur + (s + 1) * X_STRIDE, x0);
        ld16(d.delta + (s + 1) * 16, dl);
        retract_state(x0, dl, xj);
      } else {
        ld16(xb + (s + 1) * X_STRIDE, xj);
      }
      GravTerms G;
      grav_terms(gl[GL_THX], gl[GL_THY], tr.prm.gravity_magnitude, G);
      ImuPieces ip;
      imu_factor<JAC>(soa_r(d.pre, s, PRE_STRIDE), xi, xj, G, ip);
      const double q = own ? 0.5 * quad_form15(soa_r(d.info, s, INFO_STRIDE), ip.e) : 0.0;
      cost += q;
      lin0 += q;
      if (JAC) {
        const SoaW o = soa_w(d.fpi, s, FPI_STRIDE);
#pragma unroll
        for (int k = 0; k < 15; k++)
          o[FPI_E + k] = ip.e[k];
#pragma unroll
        for (int k = 0; k < 9; k++) {
          o[FPI_TTH + k] = ip.Tth[k];
          o[FPI_TB + k] = ip.Tb[k];
          o[FPI_RK + k] = ip.Rk[k];
          o[FPI_JTH2 + k] = ip.Jth2[k];
        }
#pragma unroll
        for (int k = 0; k < 3; k++) {
          o[FPI_SV + k] = ip.sv[k];
          o[FPI_SP + k] = ip.sp[k];
        }
      }
    }
  }
  // ---- fixed-order CTA reduction of (cost, lin0, g.delta, |delta|^2)
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cost += __shfl_down_sync(0xffffffffu, cost, o);
    lin0 += __shfl_down_sync(0xffffffffu, lin0, o);
    gd += __shfl_down_sync(0xffffffffu, gd, o);
    dd += __shfl_down_sync(0xffffffffu, dd, o);
  }
  if (lane == 0) {
    s_red[0][wid] = cost;
    s_red[1][wid] = lin0;
    s_red[2][wid] = gd;
    s_red[3][wid] = dd;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double c = 0;
    for (int w = 0; w < V2GT_CHUNK / 32; w++)
      c += s_red[threadIdx.x][w];
    d.part_cost[((long long)t * d.nchunk + blockIdx.x) * 4 + threadIdx.x] = c;
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int ASM_STATES = 32;            // states per CTA (one warp per 3-column block of these states)
constexpr int ASM_THREADS = ASM_STATES * 5;
constexpr int ASM_STRIP = 45;             // doubles per (state, column block) strip staged in shared memory (odd: conflict-free)
constexpr int ASM_IN_ROWS = INFO_STRIDE + FPI_VALID; // staged inputs per state: information matrix (120) + IMU factor pieces (57)
constexpr int ASM_IN_COLS = ASM_STATES + 1;          // the CTA's states + the state in front of them (odd: conflict-free)
constexpr int ASM_SMEM_DOUBLES = 5 * ASM_STATES * ASM_STRIP + ASM_STATES * 7 + ASM_IN_ROWS * ASM_IN_COLS;

// Store one strip per state of the warp: buf[state][ASM_STRIP] (shared) -> dst + state * NE_STRIDE, `len` contiguous
// doubles each.  The lanes write consecutive doubles, so every 32-byte sector leaves the SM complete (the direct
// one-thread-per-state stores touched 32 sectors per instruction and doubled the DRAM write traffic).
__device__ __forceinline__ int lane_of_warp() { return threadIdx.x & 31; }
__device__ __forceinline__ void cp_async8(double *smem, const double *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void flush_strip(const double *buf, double *dst, const int nvalid, const int len, const int lane) {
  __syncwarp();
  for (int k = 0; k < nvalid; k++) {
    const double *b = buf + k * ASM_STRIP;
    double *o = dst + (long long)k * NE_STRIDE;
    if (lane < len)
      o[lane] = b[lane];
    if (lane + 32 < len)
      o[lane + 32] = b[lane + 32];
  }
  __syncwarp();
}

// 3x3 helpers on registers:  C += A B, C += A^T B, C -= ...
__device__ __forceinline__ void mm3_acc(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] += A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
}
__device__ __forceinline__ void mm3_tn_acc(const double *A, const double *B, double *C) {
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++)
      C[3 * r + c] += A[r] * B[c] + A[3 + r] * B[3 + c] + A[6 + r] * B[6 + c];
}
__device__ __forceinline__ void zero9(double *A) {
#pragma unroll
  for (int e = 0; e < 9; e++)
    A[e] = 0.0;
}
template <class View> __device__ __forceinline__ void ld9(const View p, double *o) {
#pragma unroll
  for (int e = 0; e < 9; e++)
    o[e] = p[e];
}

// H1 block (row-block a, column-block dcol) of the IMU factor; returns false when the block is zero.
// kind: 0 dense in M, 1 = -I, (others dense).   fpi: pieces row, pre: preintegration row.
template <class VF, class VP>
__device__ __forceinline__ bool h1_block(const VF fpi, const VP pre, int a, int dcol, double *M) {
  // rows a: 0 th, 1 bg, 2 v, 3 ba, 4 p
  if (dcol == 0) {
    if (a == 0) { ld9(fpi + FPI_TTH, M); return true; }
    if (a == 2) { double s[3] = {fpi[FPI_SV], fpi[FPI_SV + 1], fpi[FPI_SV + 2]}; skew(s, M); return true; }
    if (a == 4) { double s[3] = {fpi[FPI_SP], fpi[FPI_SP + 1], fpi[FPI_SP + 2]}; skew(s, M); return true; }
    return false;
  }
  if (dcol == 1) {
    if (a == 0) { ld9(fpi + FPI_TB, M); return true; }
    if (a == 1) { eye3(M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    if (a == 2) { ld9(pre + PRE_JB, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    if (a == 4) { ld9(pre + PRE_JA, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    return false;
  }
  if (dcol == 2) {
    if (a == 2) { ld9(fpi + FPI_RK, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    if (a == 4) { const double dt = pre[PRE_DT]; ld9(fpi + FPI_RK, M); for (int e = 0; e < 9; e++) M[e] = -dt * M[e]; return true; }
    return false;
  }
  if (dcol == 3) {
    if (a == 2) { ld9(pre + PRE_HB, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    if (a == 3) { eye3(M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    if (a == 4) { ld9(pre + PRE_HA, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
    return false;
  }
  // dcol == 4
  if (a == 4) { ld9(fpi + FPI_RK, M); for (int e = 0; e < 9; e++) M[e] = -M[e]; return true; }
  return false;
}
// H2 diagonal block c of the IMU factor
template <class VF> __device__ __forceinline__ void h2_block(const VF fpi, int c, double *M) {
  if (c == 0)
    ld9(fpi + FPI_JTH2, M);
  else if (c == 2 || c == 4)
    ld9(fpi + FPI_RK, M);
  else
    eye3(M);
}
// H3 row blocks (v and p rows), 3x2 each:  R dt g Hxy  and  1/2 R dt^2 g Hxy   (ImuFactorCPIv1.cpp:199-201)
template <class VF, class VP>
__device__ __forceinline__ void h3_blocks(const VF fpi, const VP pre, const GravTerms &G, double *H3v, double *H3p) {
  double Rk[9];
  ld9(fpi + FPI_RK, Rk);
  const double dt = pre[PRE_DT];
  const double dt2 = dt * dt;
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 2; c++) {
      const double a = ((Rk[3 * r] * dt) * G.gmag) * G.Hxy[c] + ((Rk[3 * r + 1] * dt) * G.gmag) * G.Hxy[2 + c] +
                       ((Rk[3 * r + 2] * dt) * G.gmag) * G.Hxy[4 + c];
      const double b = (((0.5 * Rk[3 * r]) * dt2) * G.gmag) * G.Hxy[c] + (((0.5 * Rk[3 * r + 1]) * dt2) * G.gmag) * G.Hxy[2 + c] +
                       (((0.5 * Rk[3 * r + 2]) * dt2) * G.gmag) * G.Hxy[4 + c];
      H3v[2 * r + c] = a;
      H3p[2 * r + c] = b;
    }
}

// grid (ceil(n_max / ASM_STATES), T), ASM_THREADS threads
#ifndef ASM_MIN_CTAS
#define ASM_MIN_CTAS 2
#endif
__global__ void __launch_bounds__(ASM_THREADS, ASM_MIN_CTAS) k_assemble(DevPtrs d, int force_all) {
  const int t = blockIdx.y;
  const LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK || !lm.need_lin))
    return;
  const TrialDev &tr = d.trials[t];
  const int n = d.n_states[t];
  // thread = (column block dcol, state li) with dcol uniform per warp: no divergence on the block structure, and
  // lane <-> state makes every struct-of-arrays read coalesced
  const int li = threadIdx.x % ASM_STATES, dcol = threadIdx.x / ASM_STATES;
  const int i = blockIdx.x * ASM_STATES + li;
  extern __shared__ __align__(16) double s_strip[]; // [5][ASM_STATES * ASM_STRIP]
  double *wbuf = s_strip + dcol * (ASM_STATES * ASM_STRIP); // this warp's staging buffer
  double *mybuf = wbuf + li * ASM_STRIP; // this state's strip
  const int i0 = blockIdx.x * ASM_STATES;
  const int nvalid = min(ASM_STATES, n - i0);
  if (nvalid <= 0)
    return;
  double *ne0 = d.ne + (tr.st_off + i0) * NE_STRIDE; // record of the warp's first state
  double Y[5][9];    // Y[a] = sum_b Lambda_ab H1[b][dcol] of the IMU factor i -> i+1 (zero for the last state)
  double Bg[6];      // B[dcol][7:9]  3x2
  double gv[3];      // g[dcol]
  // warp dcol 3: gravity 2x2 block + gradient 2 of the IMU factor, parked in shared memory until the final reduction
  double *part6 = s_strip + 5 * ASM_STATES * ASM_STRIP + li * 7;
  if (dcol == 3) {
#pragma unroll
    for (int k = 0; k < 6; k++)
      part6[k] = 0.0;
  }
  const long long s = tr.st_off + i;
  const double *gl = d.glob + ((long long)lm.cur * d.T + t) * GLOB_STRIDE;
  const bool act = (i < n), has_next = (i < n - 1), has_prev = act && (i > 0);
  // ---- stage the information matrices and factor pieces of the CTA's states (+ the state in front) in shared memory:
  // every element is read ~5 x 3 times by the block products below; from global memory those reads are what the
  // kernel waited on (one CTA-wide coalesced copy instead)
  double *s_in = s_strip + 5 * ASM_STATES * ASM_STRIP + ASM_STATES * 7; // [ASM_IN_ROWS][ASM_IN_COLS], column j <-> state i0 - 1 + j
  {
    const int lane = lane_of_warp();
    const long long sl0 = tr.st_off + i0 - 1; // slot of column 0
    const bool v0 = (i0 + lane - 1 >= 0) && (i0 + lane - 1 < n);
    const bool v1 = (lane == 0) && (i0 + 31 < n);
    // asynchronous 8-byte copies: all of a thread's ~36 loads are in flight at once, no register staging
    const double *i0p = d.info + soa_base(v0 ? sl0 + lane : 0, INFO_STRIDE), *i1p = d.info + soa_base(v1 ? sl0 + 32 : 0, INFO_STRIDE);
    const double *f0p = d.fpi + soa_base(v0 ? sl0 + lane : 0, FPI_STRIDE), *f1p = d.fpi + soa_base(v1 ? sl0 + 32 : 0, FPI_STRIDE);
    for (int row = dcol; row < ASM_IN_ROWS; row += 5) {
      const bool is_info = row < INFO_STRIDE;
      const int e = is_info ? row : row - INFO_STRIDE;
      if (v0)
        cp_async8(s_in + row * ASM_IN_COLS + lane, (is_info ? i0p : f0p) + e * SOA_W);
      if (v1)
        cp_async8(s_in + row * ASM_IN_COLS + 32, (is_info ? i1p : f1p) + e * SOA_W);
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    asm volatile("cp.async.wait_group 0;" ::: "memory");
  }
  __syncthreads();
  const SoaS inf{s_in + li + 1, ASM_IN_COLS};
  const SoaS fpi{s_in + INFO_STRIDE * ASM_IN_COLS + li + 1, ASM_IN_COLS};
  const SoaR pre = soa_r(d.pre, act ? s : tr.st_off, PRE_STRIDE);
#pragma unroll
  for (int a = 0; a < 5; a++)
    zero9(Y[a]);
#pragma unroll
  for (int k = 0; k < 6; k++)
    Bg[k] = 0.0;
  gv[0] = gv[1] = gv[2] = 0.0;

  // ================= (1) IMU factor i -> i+1: Y, the coupling strip O, its border / gradient terms
  if (has_next) {
#pragma unroll
    for (int b = 0; b < 5; b++) {
      double Hb[9];
      if (!h1_block(fpi, pre, b, dcol, Hb))
        continue;
#pragma unroll
      for (int a = 0; a < 5; a++) {
        double L[9];
        ld_sym_blk(inf, a, b, L);
        mm3_acc(L, Hb, Y[a]);
      }
    }
    // O[dcol][e] = Y[e]^T H2[e][e]  -> rows 3 dcol .. 3 dcol + 2 of O (row stride 15 inside the strip)
#pragma unroll
    for (int e = 0; e < 5; e++) {
      double H2e[9], Ob[9];
      h2_block(fpi, e, H2e);
      zero9(Ob);
      mm3_tn_acc(Y[e], H2e, Ob);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 3; c++)
          mybuf[r * 15 + 3 * e + c] = Ob[3 * r + c];
    }
    // B1[dcol] = Y[v]^T H3v + Y[p]^T H3p ; g1[dcol] = -sum_a Y[a]^T e[a]
    double H3v[6], H3p[6];
    {
      GravTerms G;
      grav_terms(gl[GL_THX], gl[GL_THY], tr.prm.gravity_magnitude, G);
      h3_blocks(fpi, pre, G, H3v, H3p);
    }
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 2; c++) {
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
          sacc += Y[2][3 * k + r] * H3v[2 * k + c] + Y[4][3 * k + r] * H3p[2 * k + c];
        Bg[2 * r + c] += sacc;
      }
#pragma unroll
    for (int a = 0; a < 5; a++)
#pragma unroll
      for (int r = 0; r < 3; r++) {
        double sacc = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
          sacc += Y[a][3 * k + r] * fpi[FPI_E + 3 * a + k];
        gv[r] -= sacc;
      }
    // gravity 2x2 block and gradient of this factor (thread dcol == 3 owns it)
    if (dcol == 3 && (!d.chain || (i >= d.c_lo && i < d.c_hi))) {
      // Lambda rows v and p applied to H3 and e
      double LH[2][6]; // (Lambda H3)[a] for a = v, p  (3x2)
      double Le[2][3];
#pragma unroll
      for (int ai = 0; ai < 2; ai++) {
        const int a = (ai == 0) ? 2 : 4;
        double Lv[9], Lp[9];
        ld_sym_blk(inf, a, 2, Lv);
        ld_sym_blk(inf, a, 4, Lp);
#pragma unroll
        for (int r = 0; r < 3; r++) {
#pragma unroll
          for (int c = 0; c < 2; c++) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 3; k++)
              sacc += Lv[3 * r + k] * H3v[2 * k + c] + Lp[3 * r + k] * H3p[2 * k + c];
            LH[ai][2 * r + c] = sacc;
          }
        }
#pragma unroll
        for (int r = 0; r < 3; r++) {
          double sacc = 0;
#pragma unroll
          for (int bb = 0; bb < 5; bb++) {
            double Lb[9];
            ld_sym_blk(inf, a, bb, Lb);
#pragma unroll
            for (int k = 0; k < 3; k++)
              sacc += Lb[3 * r + k] * fpi[FPI_E + 3 * bb + k];
          }
          Le[ai][r] = sacc;
        }
      }
#pragma unroll
      for (int r = 0; r < 2; r++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double sacc = 0;
#pragma unroll
          for (int k = 0; k < 3; k++)
            sacc += H3v[2 * k + r] * LH[0][2 * k + c] + H3p[2 * k + r] * LH[1][2 * k + c];
          part6[2 * r + c] = sacc;
        }
        double sg = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
          sg += H3v[2 * k + r] * Le[0][k] + H3p[2 * k + r] * Le[1][k];
        part6[4 + r] = -sg;
      }
    }
  } else if (act) {
    // last state: no coupling block
    for (int k = 0; k < 45; k++)
      mybuf[k] = 0.0;
  }
  flush_strip(wbuf, ne0 + NE_O + 45 * dcol, nvalid, 45, lane_of_warp());

  // ================= (2) the diagonal strip, one 3x3 block at a time:
  //   D[c][dcol] = sum_a H1[a][c]^T Y[a]  +  H2[c]^T Lambda'_{c,dcol} H2[dcol] (factor i-1 -> i)  +  A_c^T A_dcol (vicon)
  // D is symmetric and stored as its packed lower triangle: the column strip this thread computes is written as the
  // ROW strip 3 dcol .. 3 dcol + 2, whose lower part is the blocks c <= dcol (rows of 3 dcol + 1, + 2, + 3 entries,
  // contiguous in the record: 9 dcol + 6 doubles starting at tri(3 dcol, 0)); the blocks c > dcol are not formed
  const SoaS infP{s_in + li, ASM_IN_COLS};
  const SoaS fpiP{s_in + INFO_STRIDE * ASM_IN_COLS + li, ASM_IN_COLS};
  const SoaR preP = soa_r(d.pre, has_prev ? s - 1 : tr.st_off, PRE_STRIDE);
  const SoaR fpv = soa_r(d.fpv, act ? s : tr.st_off, FPV_STRIDE);
  const bool vic_col = (dcol == 0 || dcol == 4);
  const int co = (dcol == 0) ? 0 : 3; // my columns inside the vicon A (theta | p)
  if (act) {
    double H2d[9];
    if (has_prev)
      h2_block(fpiP, dcol, H2d);
#pragma unroll
    for (int c = 0; c < 5; c++) {
      if (c > dcol) // warp-uniform
        continue;
      double Dc[9];
      zero9(Dc);
      if (has_next) {
#pragma unroll
        for (int a = 0; a < 5; a++) {
          double Hc[9];
          if (!h1_block(fpi, pre, a, c, Hc))
            continue;
          mm3_tn_acc(Hc, Y[a], Dc);
        }
      }
      if (has_prev) {
        double L[9], T[9], H2c[9];
        ld_sym_blk(infP, c, dcol, L);
        zero9(T);
        mm3_acc(L, H2d, T);
        h2_block(fpiP, c, H2c);
        mm3_tn_acc(H2c, T, Dc);
      }
      if (vic_col && (c == 0 || c == 4)) {
        const int cco = (c == 4) ? 3 : 0;
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int cc = 0; cc < 3; cc++) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
              sacc += fpv[FPV_A + 6 * k + cco + r] * fpv[FPV_A + 6 * k + co + cc];
            Dc[3 * r + cc] += sacc;
          }
      }
      // element (row 3 dcol + cc, column 3 c + r) of D; row cc of the strip starts at cc (3 dcol) + cc (cc + 1) / 2
#pragma unroll
      for (int cc = 0; cc < 3; cc++)
#pragma unroll
        for (int r = 0; r < 3; r++)
          if (c < dcol || r <= cc)
            mybuf[cc * 3 * dcol + (cc * (cc + 1)) / 2 + 3 * c + r] = Dc[3 * r + cc];
    }
  }
  flush_strip(wbuf, ne0 + NE_D + (3 * dcol * (3 * dcol + 1)) / 2, nvalid, 9 * dcol + 6, lane_of_warp());

  // ================= (3) border / gradient strip
  if (act) {
    if (has_prev) {
      // B2[dcol] = H2[dcol]^T (Lambda_{d,v} H3v + Lambda_{d,p} H3p) ; g2[dcol] = -H2[dcol]^T sum_a Lambda_{d,a} e_a
      double H2d[9];
      h2_block(fpiP, dcol, H2d);
      double H3v[6], H3p[6];
      {
        GravTerms G;
        grav_terms(gl[GL_THX], gl[GL_THY], tr.prm.gravity_magnitude, G);
        h3_blocks(fpiP, preP, G, H3v, H3p);
      }
      double Lv[9], Lp[9], U[6], le[3] = {0, 0, 0};
      ld_sym_blk(infP, dcol, 2, Lv);
      ld_sym_blk(infP, dcol, 4, Lp);
#pragma unroll
      for (int r = 0; r < 3; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double sacc = 0;
#pragma unroll
          for (int k = 0; k < 3; k++)
            sacc += Lv[3 * r + k] * H3v[2 * k + c] + Lp[3 * r + k] * H3p[2 * k + c];
          U[2 * r + c] = sacc;
        }
#pragma unroll
      for (int a = 0; a < 5; a++) {
        double L[9];
        ld_sym_blk(infP, dcol, a, L);
#pragma unroll
        for (int r = 0; r < 3; r++)
#pragma unroll
          for (int k = 0; k < 3; k++)
            le[r] += L[3 * r + k] * fpiP[FPI_E + 3 * a + k];
      }
#pragma unroll
      for (int r = 0; r < 3; r++) {
#pragma unroll
        for (int c = 0; c < 2; c++) {
          double sacc = 0;
#pragma unroll
          for (int k = 0; k < 3; k++)
            sacc += H2d[3 * k + r] * U[2 * k + c];
          Bg[2 * r + c] += sacc;
        }
        double sg = 0;
#pragma unroll
        for (int k = 0; k < 3; k++)
          sg += H2d[3 * k + r] * le[k];
        gv[r] -= sg;
      }
    }
    // vicon factor of state i: B[dcol][0:7] = A_d^T G, g[dcol] -= A_d^T r
#pragma unroll
    for (int r = 0; r < 3; r++) {
#pragma unroll
      for (int c = 0; c < 7; c++) {
        double sacc = 0;
        if (vic_col) {
#pragma unroll
          for (int k = 0; k < 6; k++)
            sacc += fpv[FPV_A + 6 * k + co + r] * fpv[FPV_G + 7 * k + c];
        }
        mybuf[NE_BGW * r + c] = sacc;
      }
      if (vic_col) {
        double sg = 0;
#pragma unroll
        for (int k = 0; k < 6; k++)
          sg += fpv[FPV_A + 6 * k + co + r] * fpv[FPV_R + k];
        gv[r] -= sg;
      }
      mybuf[NE_BGW * r + 7] = Bg[2 * r];
      mybuf[NE_BGW * r + 8] = Bg[2 * r + 1];
      mybuf[NE_BGW * r + 9] = gv[r];
      d.grad[s * 16 + 3 * dcol + r] = gv[r];
    }
  }
  flush_strip(wbuf, ne0 + NE_BG + 3 * NE_BGW * dcol, nvalid, 3 * NE_BGW, lane_of_warp());
  // ================= per-CTA partial sums of the global blocks (fixed order over the CTA's states)
  //   warp dcol 1: vicon 7x7 block G^T G (28 unique entries) and gradient -G^T r (7);  warp dcol 3: gravity 2x2 + gradient 2
  if (dcol == 1 || dcol == 3) {
    if (dcol == 1) {
      if (i < n) {
        const SoaR fpv = soa_r(d.fpv, s, FPV_STRIDE);
        int q = 0;
#pragma unroll
        for (int r = 0; r < 7; r++) {
#pragma unroll
          for (int c = r; c < 7; c++) {
            double sacc = 0;
#pragma unroll
            for (int k = 0; k < 6; k++)
              sacc += fpv[FPV_G + 7 * k + r] * fpv[FPV_G + 7 * k + c];
            mybuf[q++] = sacc;
          }
          double sg = 0;
#pragma unroll
          for (int k = 0; k < 6; k++)
            sg += fpv[FPV_G + 7 * k + r] * fpv[FPV_R + k];
          mybuf[28 + r] = -sg;
        }
      } else {
        for (int k = 0; k < 35; k++)
          mybuf[k] = 0.0;
      }
    } else {
#pragma unroll
      for (int k = 0; k < 6; k++)
        mybuf[k] = part6[k];
    }
    __syncwarp();
    const int lane = lane_of_warp();
    const int ncol = (dcol == 1) ? 35 : 6;
    double sacc = 0.0, sacc2 = 0.0;
    for (int k = 0; k < ASM_STATES; k++) {
      if (lane < ncol)
        sacc += wbuf[k * ASM_STRIP + lane];
      if (lane + 32 < ncol)
        sacc2 += wbuf[k * ASM_STRIP + lane + 32];
    }
    double *po = d.part_asm + ((long long)t * gridDim.x + blockIdx.x) * PART_STRIDE;
    if (dcol == 1) {
      // unpack the upper triangle (row-major, q = index of (r, c >= r)) into the full 7x7 + gradient
      auto put = [&](int q, double v) {
        if (q < 28) {
          int r = 0, base = 0;
          while (q >= base + (7 - r)) {
            base += 7 - r;
            r++;
          }
          const int c = r + (q - base);
          po[7 * r + c] = v;
          po[7 * c + r] = v;
        } else {
          po[49 + (q - 28)] = v;
        }
      };
      put(lane, sacc);
      if (lane + 32 < ncol)
        put(lane + 32, sacc2);
    } else if (lane < 6) {
      po[56 + lane] = sacc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// grid T, 64 threads: gamma[t] = [Gamma 9x9 | g 9], lm.lin_error0, from the per-CTA partials (fixed order).
__global__ void __launch_bounds__(64) k_reduce_gamma(DevPtrs d, int n_asm_chunks, int force_all) {
  const int t = blockIdx.x;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK || !lm.need_lin))
    return;
  const int n = d.n_states[t];
  const int used = (n + ASM_STATES - 1) / ASM_STATES;
  __shared__ double s_v[64];
  double sacc = 0;
  if (threadIdx.x < 62)
    for (int k = 0; k < used && k < n_asm_chunks; k++)
      sacc += d.part_asm[((long long)t * n_asm_chunks + k) * PART_STRIDE + threadIdx.x];
  s_v[threadIdx.x] = sacc;
  __syncthreads();
  // chain-sharded mode: this is only this rank's part; it is all-reduced into gamma by the host side
  double *gm = d.chain ? d.c_gamma_loc : d.gamma + (long long)t * 96;
  for (int k = threadIdx.x; k < 90; k += 64) {
    double v = 0.0;
    if (k < 81) {
      const int r = k / 9, c = k % 9;
      if (r < 7 && c < 7)
        v = s_v[7 * r + c];
      else if (r >= 7 && c >= 7)
        v = s_v[56 + 2 * (r - 7) + (c - 7)];
    } else {
      const int r = k - 81;
      v = (r < 7) ? s_v[49 + r] : s_v[60 + (r - 7)];
    }
    gm[k] = v;
  }
  if (threadIdx.x == 0) {
    const int usedc = (n + V2GT_CHUNK - 1) / V2GT_CHUNK;
    double l0 = 0;
    for (int k = 0; k < usedc; k++)
      l0 += d.part_cost[((long long)t * d.nchunk + k) * 4 + 1];
    lm.lin_error0 = l0;
    if (d.chain)
      gm[90] = l0;
  }
}

// ------------------------------------------------------------------------------------------------
int launch_linearize(const DevPtrs &d, int n_max, int force_all, cudaStream_t st) {
  if (n_max <= 0)
    return V2GT_OK;
  dim3 g1((n_max + V2GT_CHUNK - 1) / V2GT_CHUNK, d.T);
  k_factor_eval<true><<<g1, V2GT_CHUNK, 0, st>>>(d, 0, force_all);
  const int nasm = (n_max + ASM_STATES - 1) / ASM_STATES;
  dim3 g2(nasm, d.T);
  constexpr size_t asm_smem = sizeof(double) * ASM_SMEM_DOUBLES;
  // per device (context), hence on every launch: a handle may live on any GPU of the process
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_assemble, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)asm_smem));
  k_assemble<<<g2, ASM_THREADS, asm_smem, st>>>(d, force_all);
  k_reduce_gamma<<<d.T, 64, 0, st>>>(d, nasm, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

int launch_cost(const DevPtrs &d, int n_max, int which, int force_all, cudaStream_t st) {
  if (n_max <= 0)
    return V2GT_OK;
  dim3 g1((n_max + V2GT_CHUNK - 1) / V2GT_CHUNK, d.T);
  k_factor_eval<false><<<g1, V2GT_CHUNK, 0, st>>>(d, which, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

int asm_chunks(int n_max) { return (n_max + ASM_STATES - 1) / ASM_STATES; }

} // namespace v2
