// vicon2gt_b200 kernel (4b): damped normal-equation solve on the chain's block-tridiagonal-plus-arrow
// structure.  Replaces GTSAM's buildDampedSystem + multifrontal Cholesky (called from
// LevenbergMarquardtOptimizer::tryLambda for the graph of ViconGraphSolver.cpp:390-536).
//
// The chain of n 15x15 blocks of one trial is cut into segments separated by single "separator"
// states.  One warp owns one segment:
//   k_seg_eliminate  block Cholesky through the segment's interior, left to right, carrying the
//                    fill-in towards the segment's left separator (F columns), its border to the 9
//                    globals (B columns) and the right-hand side; Schur contributions to the two
//                    separators and to the global 9x9 block are left in registers/shared memory and
//                    written once per segment
//   k_reduced        one warp per trial: block Cholesky over the separator chain, Schur complement
//                    onto the 9 globals, 9x9 solve, back-substitution of the separators, retract of
//                    the globals
//   k_seg_backsub    back-substitution through the segment interiors + retract of the states
// Every trial and every segment takes the same arithmetic path in the same order: results are
// deterministic run to run.
#include "v2gt_common.cuh"
#include <stdio.h>

namespace v2 {

struct SegGeom {
  int stride, nseg, nsep;
};
__host__ __device__ inline SegGeom seg_geom(int n, int P) {
  SegGeom g;
  int st = (n + 1 + P - 1) / P;
  g.stride = st < 2 ? 2 : st;
  g.nseg = (n + g.stride - 1) / g.stride;
  g.nsep = n / g.stride;
  return g;
}

// per-warp shared memory working set of the elimination (doubles)
constexpr int W_D = 0;              // 15x15 current diagonal block -> L (lower)
constexpr int W_M = 225;            // 15x40 [O | F | B | g] -> W = L^-1 [.]
constexpr int W_CD = W_M + 600;     // carry to the next node: D 225, F 225, B 135, g 15
constexpr int W_CF = W_CD + 225;
constexpr int W_CB = W_CF + 225;
constexpr int W_CG = W_CB + 135;
constexpr int W_SL = W_CG + 15;     // accumulators for the left separator: 225, 135, 15
constexpr int W_BL = W_SL + 225;
constexpr int W_GL = W_BL + 135;
constexpr int W_GS = W_GL + 15;     // accumulators for the globals: 81, 9
constexpr int W_GG = W_GS + 81;
constexpr int W_TOTAL = W_GG + 9 + 1; // 1891 -> even
constexpr int ELIM_WARPS = 4;

// In-place Cholesky of the 15x15 block at w[W_D] (lower triangle used) followed by the forward
// substitution of the 40 columns of w[W_M].  Returns false on a non-positive pivot.
__device__ __forceinline__ bool chol_trsm(double *w, int lane, int ncols_active_from) {
  double *D = w + W_D;
  double *M = w + W_M;
  bool ok = true;
  for (int j = 0; j < 15; j++) {
    const double djj = D[16 * j];
    if (!(djj > 0.0))
      ok = false;
    const double ljj = sqrt(djj);
    const double inv = 1.0 / ljj;
    __syncwarp();
    if (lane == 0)
      D[16 * j] = ljj;
    const int r = j + 1 + lane;
    if (r < 15)
      D[15 * r + j] *= inv;
    __syncwarp();
    // trailing update of the lower triangle: rows r > j, columns j < c <= r
    if (r < 15) {
      const double lrj = D[15 * r + j];
      for (int c = j + 1; c <= r; c++)
        D[15 * r + c] -= lrj * D[15 * c + j];
    }
    __syncwarp();
  }
  // forward substitution, one column per lane (two passes over the 40 columns)
  for (int c = lane + ncols_active_from; c < 40; c += 32) {
    double x[15];
#pragma unroll
    for (int r = 0; r < 15; r++)
      x[r] = M[40 * r + c];
#pragma unroll
    for (int r = 0; r < 15; r++) {
      double s = x[r];
#pragma unroll
      for (int k = 0; k < r; k++)
        s -= D[15 * r + k] * x[k];
      x[r] = s / D[16 * r];
    }
#pragma unroll
    for (int r = 0; r < 15; r++)
      M[40 * r + c] = x[r];
  }
  __syncwarp();
  return ok;
}

// Schur updates from W (15x40 at w[W_M]): carry (next node) = -W_O^T [W_O W_F W_B w_g];
// left-separator accumulators -= W_F^T [W_F W_B w_g]; global accumulators -= W_B^T [W_B w_g].
template <bool HAS_F>
__device__ __forceinline__ void schur_update(double *w, int lane) {
  const double *M = w + W_M;
  const int total = HAS_F ? 1065 : 1065; // index space is shared; F pairs are skipped when !HAS_F
  for (int idx = lane; idx < total; idx += 32) {
    int a, b;
    if (idx < 600) {
      a = idx / 40;
      b = idx - 40 * a;
    } else if (idx < 975) {
      const int i2 = idx - 600;
      a = 15 + i2 / 25;
      b = 15 + (i2 - 25 * (i2 / 25));
    } else {
      const int i3 = idx - 975;
      a = 30 + i3 / 10;
      b = 30 + (i3 - 10 * (i3 / 10));
    }
    if (!HAS_F && ((a >= 15 && a < 30) || (b >= 15 && b < 30)))
      continue;
    double s = 0;
#pragma unroll
    for (int k = 0; k < 15; k++)
      s += M[40 * k + a] * M[40 * k + b];
    if (a < 15) {
      if (b < 15)
        w[W_CD + 15 * a + b] = -s;
      else if (b < 30)
        w[W_CF + 15 * a + (b - 15)] = -s;
      else if (b < 39)
        w[W_CB + 9 * a + (b - 30)] = -s;
      else
        w[W_CG + a] = -s;
    } else if (a < 30) {
      if (b < 30)
        w[W_SL + 15 * (a - 15) + (b - 15)] -= s;
      else if (b < 39)
        w[W_BL + 9 * (a - 15) + (b - 30)] -= s;
      else
        w[W_GL + (a - 15)] -= s;
    } else {
      if (b < 39)
        w[W_GS + 9 * (a - 30) + (b - 30)] -= s;
      else
        w[W_GG + (a - 30)] -= s;
    }
  }
  __syncwarp();
}

// store L (packed lower) and W to the factor row of a node
__device__ __forceinline__ void store_factor(const double *w, double *__restrict__ fac, int lane, bool has_f) {
  for (int idx = lane; idx < 120; idx += 32) {
    // idx -> (r, c) of the packed lower triangle
    int r = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
    while (tri(r + 1, 0) <= idx)
      r++;
    while (tri(r, 0) > idx)
      r--;
    const int c = idx - tri(r, 0);
    fac[FAC_L + idx] = w[W_D + 15 * r + c];
  }
  for (int idx = lane; idx < 225; idx += 32) {
    const int r = idx / 15, c = idx - 15 * r;
    fac[FAC_WO + idx] = w[W_M + 40 * r + c];
    fac[FAC_WF + idx] = has_f ? w[W_M + 40 * r + 15 + c] : 0.0;
  }
  for (int idx = lane; idx < 135; idx += 32) {
    const int r = idx / 9, c = idx - 9 * r;
    fac[FAC_WB + idx] = w[W_M + 40 * r + 30 + c];
  }
  if (lane < 15)
    fac[FAC_WG + lane] = w[W_M + 40 * lane + 39];
}

// ------------------------------------------------------------------------------------------------
// grid: (ceil(P / ELIM_WARPS), T); block: ELIM_WARPS warps; warp -> segment j of trial t
__global__ void __launch_bounds__(ELIM_WARPS * 32) k_seg_eliminate(DevPtrs d, int force_all) {
  extern __shared__ double smem[];
  const int t = blockIdx.y;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * ELIM_WARPS + wid;
  const int n = d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  if (j >= sg.nseg)
    return;
  double *w = smem + (size_t)wid * W_TOTAL;
  const TrialDev &tr = d.trials[t];
  const double lambda = lm.lambda;
  const int a0 = j * sg.stride;
  const int b1 = min(n, (j + 1) * sg.stride - 1); // exclusive end of the interior
  const bool has_left = (j >= 1);
  const bool has_right = ((j + 1) * sg.stride - 1 < n);
  for (int k = lane; k < W_TOTAL; k += 32)
    w[k] = 0.0;
  __syncwarp();
  bool ok = true;
  for (int k = a0; k < b1; k++) {
    const long long s = tr.st_off + k;
    const double *HD = d.HD + s * HD_STRIDE;
    const double *HO = d.HO + s * HO_STRIDE;
    const double *HB = d.HB + s * HB_STRIDE;
    const double *Hg = d.Hg + s * HG_STRIDE;
    // current blocks = assembled blocks + carry (+ lambda I)
    for (int idx = lane; idx < 225; idx += 32) {
      const int r = idx / 15, c = idx - 15 * r;
      w[W_D + idx] = HD[idx] + w[W_CD + idx] + ((r == c) ? lambda : 0.0);
      w[W_M + 40 * r + c] = HO[idx];
      double f = w[W_CF + idx];
      if (k == a0 && has_left)
        f = d.HO[(s - 1) * HO_STRIDE + 15 * c + r]; // H[a0][L] = O_L^T
      w[W_M + 40 * r + 15 + c] = f;
    }
    for (int idx = lane; idx < 135; idx += 32) {
      const int r = idx / 9, c = idx - 9 * r;
      w[W_M + 40 * r + 30 + c] = HB[idx] + w[W_CB + idx];
    }
    if (lane < 15)
      w[W_M + 40 * lane + 39] = Hg[lane] + w[W_CG + lane];
    __syncwarp();
    ok = chol_trsm(w, lane, 0) && ok;
    store_factor(w, d.fac + s * FAC_STRIDE, lane, has_left);
    if (has_left)
      schur_update<true>(w, lane);
    else
      schur_update<false>(w, lane);
  }
  // segment outputs
  double *so = d.seg + ((long long)t * d.P + j) * SEG_STRIDE;
  for (int idx = lane; idx < 225; idx += 32) {
    so[SEG_SL + idx] = has_left ? w[W_SL + idx] : 0.0;
    so[SEG_DR + idx] = has_right ? w[W_CD + idx] : 0.0;
    so[SEG_C + idx] = (has_right && has_left) ? w[W_CF + idx] : 0.0;
  }
  for (int idx = lane; idx < 135; idx += 32) {
    so[SEG_BL + idx] = has_left ? w[W_BL + idx] : 0.0;
    so[SEG_BR + idx] = has_right ? w[W_CB + idx] : 0.0;
  }
  if (lane < 15) {
    so[SEG_GL + lane] = has_left ? w[W_GL + lane] : 0.0;
    so[SEG_GR + lane] = has_right ? w[W_CG + lane] : 0.0;
  }
  for (int idx = lane; idx < 81; idx += 32)
    so[SEG_GS + idx] = w[W_GS + idx];
  if (lane < 9)
    so[SEG_GG + lane] = w[W_GG + lane];
  if (!ok && lane == 0)
    atomicExch(&lm.solve_failed, 1);
}

// ------------------------------------------------------------------------------------------------
// retract of the four global nodes: JPLQuaternion C(0), Vector3 C(1), Vector1 T(0), RotationXY G(0)
__device__ inline void retract_globals(const double *g0, const double *dx, double *g1) {
  double dq[4], q[4];
  small_quat(dx, dq);
  quat_mul(dq, g0 + GL_Q, q);
  g1[0] = q[0]; g1[1] = q[1]; g1[2] = q[2]; g1[3] = q[3];
  g1[GL_P] = g0[GL_P] + dx[3];
  g1[GL_P + 1] = g0[GL_P + 1] + dx[4];
  g1[GL_P + 2] = g0[GL_P + 2] + dx[5];
  g1[GL_TOFF] = g0[GL_TOFF] + dx[6];
  g1[GL_THX] = wrap2pi(g0[GL_THX] + wrap2pi(dx[7]));
  g1[GL_THY] = wrap2pi(g0[GL_THY] + wrap2pi(dx[8]));
  for (int k = 10; k < GLOB_STRIDE; k++)
    g1[k] = 0.0;
}

// back-substitution of one node held in shared memory:  x = L^-T (w_g - W_O x_next - W_F x_left - W_B x_glob)
// f: factor row (720 doubles, shared or global), lanes 0..14 return x[lane].
__device__ __forceinline__ double backsub_node(const double *f, int lane, const double *x_next, const double *x_left,
                                               const double *x_glob, bool has_next, bool has_left) {
  double rhs = 0.0;
  if (lane < 15) {
    rhs = f[FAC_WG + lane];
    if (has_next) {
#pragma unroll
      for (int a = 0; a < 15; a++)
        rhs -= f[FAC_WO + 15 * lane + a] * x_next[a];
    }
    if (has_left) {
#pragma unroll
      for (int a = 0; a < 15; a++)
        rhs -= f[FAC_WF + 15 * lane + a] * x_left[a];
    }
#pragma unroll
    for (int a = 0; a < 9; a++)
      rhs -= f[FAC_WB + 9 * lane + a] * x_glob[a];
  }
  // L^T x = rhs, backward; lane c finalises x_c, everyone below subtracts
  double x = 0.0;
  for (int c = 14; c >= 0; c--) {
    double xc = 0.0;
    if (lane == c) {
      xc = rhs / f[FAC_L + tri(c, c)];
      x = xc;
    }
    xc = __shfl_sync(0xffffffffu, xc, c);
    if (lane < c)
      rhs -= f[FAC_L + tri(c, lane)] * xc;
  }
  return x;
}

constexpr int RED_SMEM = W_TOTAL + 32; // + x_next (15) + x_glob (9) scratch

// grid: T CTAs of one warp.
__global__ void __launch_bounds__(32) k_reduced(DevPtrs d, int force_all) {
  extern __shared__ double smem[];
  const int t = blockIdx.x;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int lane = threadIdx.x;
  const int n = d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  const TrialDev &tr = d.trials[t];
  const double lambda = lm.lambda;
  double *w = smem;
  double *xs = smem + W_TOTAL; // [0:15) x_next, [16:25) x_glob
  for (int k = lane; k < RED_SMEM; k += 32)
    w[k] = 0.0;
  __syncwarp();
  bool ok = true;
  const double *segs = d.seg + (long long)t * d.P * SEG_STRIDE;
  // ---- forward elimination over the separators m = 1..nsep (node m*stride-1)
  for (int m = 1; m <= sg.nsep; m++) {
    const long long s = tr.st_off + (long long)m * sg.stride - 1;
    const double *HD = d.HD + s * HD_STRIDE;
    const double *HB = d.HB + s * HB_STRIDE;
    const double *Hg = d.Hg + s * HG_STRIDE;
    const double *sl = segs + (long long)(m - 1) * SEG_STRIDE; // segment on the left  (this node is its right separator)
    const bool has_rseg = (m < sg.nseg);                        // segment on the right (this node is its left separator)
    const double *sr = segs + (long long)m * SEG_STRIDE;
    const bool has_next = (m + 1 <= sg.nsep);
    for (int idx = lane; idx < 225; idx += 32) {
      const int r = idx / 15, c = idx - 15 * r;
      double v = HD[idx] + w[W_CD + idx] + sl[SEG_DR + idx] + ((r == c) ? lambda : 0.0);
      if (has_rseg)
        v += sr[SEG_SL + idx];
      w[W_D + idx] = v;
      // coupling to the next separator: H[s_m][s_m+1] = C^T of the right segment
      w[W_M + 40 * r + c] = (has_next && has_rseg) ? sr[SEG_C + 15 * c + r] : 0.0;
      w[W_M + 40 * r + 15 + c] = 0.0;
    }
    for (int idx = lane; idx < 135; idx += 32) {
      const int r = idx / 9, c = idx - 9 * r;
      double v = HB[idx] + w[W_CB + idx] + sl[SEG_BR + idx];
      if (has_rseg)
        v += sr[SEG_BL + idx];
      w[W_M + 40 * r + 30 + c] = v;
    }
    if (lane < 15) {
      double v = Hg[lane] + w[W_CG + lane] + sl[SEG_GR + lane];
      if (has_rseg)
        v += sr[SEG_GL + lane];
      w[W_M + 40 * lane + 39] = v;
    }
    __syncwarp();
    ok = chol_trsm(w, lane, 0) && ok;
    // factor row of the reduced chain
    double *rf = d.red + ((long long)t * d.P + m) * RED_STRIDE;
    for (int idx = lane; idx < 120; idx += 32) {
      int r = 0;
      while (tri(r + 1, 0) <= idx)
        r++;
      rf[RED_L + idx] = w[W_D + 15 * r + (idx - tri(r, 0))];
    }
    for (int idx = lane; idx < 225; idx += 32) {
      const int r = idx / 15, c = idx - 15 * r;
      rf[RED_WO + idx] = w[W_M + 40 * r + c];
    }
    for (int idx = lane; idx < 135; idx += 32) {
      const int r = idx / 9, c = idx - 9 * r;
      rf[RED_WB + idx] = w[W_M + 40 * r + 30 + c];
    }
    if (lane < 15)
      rf[RED_WG + lane] = w[W_M + 40 * lane + 39];
    schur_update<false>(w, lane);
  }
  // ---- global 9x9: Gamma + lambda I + segment contributions (fixed order) + separator contributions
  __syncwarp();
  const double *gm = d.gamma + (long long)t * 96;
  for (int idx = lane; idx < 90; idx += 32) {
    double v = gm[idx];
    if (idx < 81) {
      v += w[W_GS + idx];
      for (int jj = 0; jj < sg.nseg; jj++)
        v += segs[(long long)jj * SEG_STRIDE + SEG_GS + idx];
      if (idx / 9 == idx % 9)
        v += lambda;
    } else {
      v += w[W_GG + idx - 81];
      for (int jj = 0; jj < sg.nseg; jj++)
        v += segs[(long long)jj * SEG_STRIDE + SEG_GG + idx - 81];
    }
    xs[32 + idx] = v; // staged, then copied to G9 (W_GS/W_GG live inside w, W_D is free now)
  }
  __syncwarp();
  // (RED_SMEM is W_TOTAL + 32; stage area overlaps nothing we still need: use the M region instead)
  for (int idx = lane; idx < 90; idx += 32)
    w[W_M + idx] = xs[32 + idx];
  __syncwarp();
  double *A = w + W_M; // 9x9 + rhs(9)
  if (lane == 0) {
    // symmetrise, Cholesky, solve (tiny, sequential)
    for (int r = 0; r < 9; r++)
      for (int c = 0; c < r; c++) {
        const double v = 0.5 * (A[9 * r + c] + A[9 * c + r]);
        A[9 * r + c] = v;
        A[9 * c + r] = v;
      }
    for (int jj = 0; jj < 9; jj++) {
      double dd = A[10 * jj];
      for (int k = 0; k < jj; k++)
        dd -= A[9 * jj + k] * A[9 * jj + k];
      if (!(dd > 0.0))
        ok = false;
      const double l = sqrt(dd);
      A[10 * jj] = l;
      for (int r = jj + 1; r < 9; r++) {
        double sacc = A[9 * r + jj];
        for (int k = 0; k < jj; k++)
          sacc -= A[9 * r + k] * A[9 * jj + k];
        A[9 * r + jj] = sacc / l;
      }
    }
    double y[9], x[9];
    for (int r = 0; r < 9; r++) {
      double sacc = A[81 + r];
      for (int k = 0; k < r; k++)
        sacc -= A[9 * r + k] * y[k];
      y[r] = sacc / A[10 * r];
    }
    for (int r = 8; r >= 0; r--) {
      double sacc = y[r];
      for (int k = r + 1; k < 9; k++)
        sacc -= A[9 * k + r] * x[k];
      x[r] = sacc / A[10 * r];
    }
    double gd = 0, dd2 = 0;
    for (int r = 0; r < 9; r++) {
      xs[16 + r] = x[r];
      gd += gm[81 + r] * x[r];
      dd2 += x[r] * x[r];
    }
    double *xg = d.xg + (long long)t * 16;
    for (int r = 0; r < 9; r++)
      xg[r] = x[r];
    xg[10] = gd;
    xg[11] = dd2;
  }
  ok = __shfl_sync(0xffffffffu, ok ? 1 : 0, 0) && ok;
  __syncwarp();
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0)
      atomicExch(&lm.solve_failed, 1);
    return;
  }
  // ---- back-substitution of the separators (right to left)
  double gd = 0, dd2 = 0;
  for (int m = sg.nsep; m >= 1; m--) {
    const double *rf = d.red + ((long long)t * d.P + m) * RED_STRIDE;
    // gather into the FAC layout expected by backsub_node (W_F absent)
    double *f = w; // W_TOTAL >= 720
    for (int idx = lane; idx < 120; idx += 32)
      f[FAC_L + idx] = rf[RED_L + idx];
    for (int idx = lane; idx < 225; idx += 32)
      f[FAC_WO + idx] = rf[RED_WO + idx];
    for (int idx = lane; idx < 135; idx += 32)
      f[FAC_WB + idx] = rf[RED_WB + idx];
    if (lane < 15)
      f[FAC_WG + lane] = rf[RED_WG + lane];
    __syncwarp();
    const double x = backsub_node(f, lane, xs, xs, xs + 16, m < sg.nsep, false);
    __syncwarp();
    const long long s = tr.st_off + (long long)m * sg.stride - 1;
    if (lane < 15) {
      xs[lane] = x;
      d.xsep[((long long)t * d.P + m) * 16 + lane] = x;
      gd += d.Hg[s * HG_STRIDE + lane] * x;
      dd2 += x * x;
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gd += __shfl_down_sync(0xffffffffu, gd, o);
    dd2 += __shfl_down_sync(0xffffffffu, dd2, o);
  }
  if (lane == 0) {
    double *xg = d.xg + (long long)t * 16;
    xg[12] = gd;
    xg[13] = dd2;
    // trial point of the globals
    const double *g0 = d.glob + ((long long)lm.cur * d.T + t) * GLOB_STRIDE;
    double *g1 = d.glob + ((long long)(1 - lm.cur) * d.T + t) * GLOB_STRIDE;
    retract_globals(g0, xg, g1);
  }
}

// ------------------------------------------------------------------------------------------------
// JPLNavState::retract (src/gtsam/JPLNavState.cpp:25-54) of one state row
__device__ inline void retract_state(const double *x0, const double *dx, double *x1) {
  double dq[4], q[4];
  small_quat(dx, dq);
  quat_mul(dq, x0, q);
  x1[0] = q[0]; x1[1] = q[1]; x1[2] = q[2]; x1[3] = q[3];
#pragma unroll
  for (int k = 0; k < 12; k++)
    x1[4 + k] = x0[4 + k] + dx[3 + k];
}

constexpr int BS_WARPS = 4;
constexpr int BS_SMEM_PER_WARP = 720 + 64; // factor row + x_next(16) x_left(16) x_glob(16) delta(16)

// grid: (ceil(P / BS_WARPS), T); warp -> segment j of trial t
__global__ void __launch_bounds__(BS_WARPS * 32) k_seg_backsub(DevPtrs d, int force_all) {
  extern __shared__ double smem[];
  const int t = blockIdx.y;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK || lm.solve_failed))
    return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * BS_WARPS + wid;
  const int n = d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  if (j >= sg.nseg)
    return;
  const TrialDev &tr = d.trials[t];
  double *f = smem + (size_t)wid * BS_SMEM_PER_WARP;
  double *x_next = f + 720, *x_left = f + 736, *x_glob = f + 752, *dl = f + 768;
  const int a0 = j * sg.stride;
  const int b1 = min(n, (j + 1) * sg.stride - 1);
  const bool has_left = (j >= 1);
  const bool has_right = ((j + 1) * sg.stride - 1 < n);
  if (lane < 16) {
    x_next[lane] = (has_right && lane < 15) ? d.xsep[((long long)t * d.P + (j + 1)) * 16 + lane] : 0.0;
    x_left[lane] = (has_left && lane < 15) ? d.xsep[((long long)t * d.P + j) * 16 + lane] : 0.0;
    x_glob[lane] = (lane < 9) ? d.xg[(long long)t * 16 + lane] : 0.0;
  }
  __syncwarp();
  const int cur = lm.cur;
  const double *xc = d.x + (long long)cur * d.S * X_STRIDE;
  double *xo = d.x + (long long)(1 - cur) * d.S * X_STRIDE;
  double gd = 0, dd2 = 0;
  // the separators themselves are retracted by the segment on their right (or by segment nseg-1 for a trailing one)
  for (int k = b1 - 1; k >= a0; k--) {
    const long long s = tr.st_off + k;
    const double *fg = d.fac + s * FAC_STRIDE;
    for (int idx = lane; idx < 720; idx += 32)
      f[idx] = fg[idx];
    __syncwarp();
    const bool has_next = (k < b1 - 1) || has_right;
    const double x = backsub_node(f, lane, x_next, x_left, x_glob, has_next, has_left);
    __syncwarp();
    if (lane < 15) {
      x_next[lane] = x;
      dl[lane] = x;
      d.delta[s * 16 + lane] = x;
      gd += d.Hg[s * HG_STRIDE + lane] * x;
      dd2 += x * x;
    }
    __syncwarp();
    if (lane == 0)
      retract_state(xc + s * X_STRIDE, dl, xo + s * X_STRIDE);
    __syncwarp();
  }
  // retract the separator states: left separator of this segment; a trailing separator (no right segment) by its left segment
  if (lane == 0) {
    if (has_left) {
      const long long s = tr.st_off + a0 - 1;
      for (int k = 0; k < 15; k++)
        d.delta[s * 16 + k] = x_left[k];
      retract_state(xc + s * X_STRIDE, x_left, xo + s * X_STRIDE);
    }
    if (has_right && j == sg.nseg - 1) {
      const long long s = tr.st_off + b1;
      double xr[15];
      for (int k = 0; k < 15; k++) {
        xr[k] = d.xsep[((long long)t * d.P + (j + 1)) * 16 + k];
        d.delta[s * 16 + k] = xr[k];
      }
      retract_state(xc + s * X_STRIDE, xr, xo + s * X_STRIDE);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    gd += __shfl_down_sync(0xffffffffu, gd, o);
    dd2 += __shfl_down_sync(0xffffffffu, dd2, o);
  }
  if (lane == 0) {
    double *so = d.seg + ((long long)t * d.P + j) * SEG_STRIDE;
    so[SEG_DOT] = gd;
    so[SEG_DOT + 1] = dd2;
  }
}

// ------------------------------------------------------------------------------------------------
int solve_configure() {
  const size_t sm1 = sizeof(double) * W_TOTAL * ELIM_WARPS;
  const size_t sm2 = sizeof(double) * (RED_SMEM + 128);
  const size_t sm3 = sizeof(double) * BS_SMEM_PER_WARP * BS_WARPS;
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_seg_eliminate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_reduced, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_seg_backsub, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm3));
  return V2GT_OK;
}
int launch_eliminate(const DevPtrs &d, int force_all, cudaStream_t st) {
  const size_t sm1 = sizeof(double) * W_TOTAL * ELIM_WARPS;
  dim3 g1((d.P + ELIM_WARPS - 1) / ELIM_WARPS, d.T);
  k_seg_eliminate<<<g1, ELIM_WARPS * 32, sm1, st>>>(d, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_reduced(const DevPtrs &d, int force_all, cudaStream_t st) {
  const size_t sm2 = sizeof(double) * (RED_SMEM + 128);
  k_reduced<<<d.T, 32, sm2, st>>>(d, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_backsub(const DevPtrs &d, int force_all, cudaStream_t st) {
  const size_t sm3 = sizeof(double) * BS_SMEM_PER_WARP * BS_WARPS;
  dim3 g3((d.P + BS_WARPS - 1) / BS_WARPS, d.T);
  k_seg_backsub<<<g3, BS_WARPS * 32, sm3, st>>>(d, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_solve(const DevPtrs &d, int force_all, cudaStream_t st) {
  int s = solve_configure();
  if (s != V2GT_OK)
    return s;
  if ((s = launch_eliminate(d, force_all, st)) != V2GT_OK)
    return s;
  if ((s = launch_reduced(d, force_all, st)) != V2GT_OK)
    return s;
  return launch_backsub(d, force_all, st);
}

} // namespace v2
