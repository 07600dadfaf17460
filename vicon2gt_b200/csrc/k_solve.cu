// vicon2gt_b200 kernel (4b): damped normal-equation solve on the chain's block-tridiagonal-plus-arrow
// structure.  Replaces GTSAM's buildDampedSystem + multifrontal Cholesky (called from
// LevenbergMarquardtOptimizer::tryLambda for the graph of ViconGraphSolver.cpp:390-536).
//
// The chain of n 15x15 blocks of one trial is cut into segments separated by single "separator"
// states.  One warp owns one segment and walks it state by state:
//   k_seg_eliminate     per state: D = L L^T as an LDL^T factorisation carried out together with the
//                       Gauss-Jordan inverse L^-1 (both live in registers, distributed in the 8x8 tile
//                       order of the FP64 mma fragments); W = L^-1 [O | F | B | g] and the Schur update
//                       S = W^T W are chains of mma.sync.m8n8k4.f64 on register fragments (no shared
//                       memory traffic in the inner products); the state's normal-equation record is
//                       streamed HBM -> shared memory with cp.async one state ahead.  F is the fill-in
//                       towards the segment's left separator, B the border to the 9 globals, g the rhs.
//   k_reduced_assemble  normal-equation records of the separator chain from the segment outputs
//   k_reduced           one warp per trial: the same elimination over the separator chain, the Schur
//                       complement onto the 9 globals, the 9x9 solve, back-substitution of the separators,
//                       retract of the globals
//   k_seg_backsub       back-substitution through the segment interiors (writes the step delta)
// Nested level 2 (launch_reduced): for P >= 24 the separator chain -- records red_ne, same format as ne -- is itself
// cut into ~sqrt(P) segments and run through k_seg_eliminate / k_reduced_assemble / k_reduced / k_seg_backsub with a
// derived pointer set (k_l2_prepare, k_sep_delta are the glue), so that one warp only walks ~sqrt(P) separators.
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

// ---- per-warp shared memory (doubles): two staged normal-equation records, the broadcast buffers of the
// factorisation (two pivot columns of D [16][2] + two pivot rows of X [2][16], double buffered), two mbarriers, and
// (with ELIM_ACC_SMEM) the persistent Schur sums of the segment
#ifndef ELIM_WARPS_N
#define ELIM_WARPS_N 4
#endif
constexpr int ELIM_WARPS = ELIM_WARPS_N;
#ifndef ELIM_MIN_CTAS
#define ELIM_MIN_CTAS 3
#endif
#ifndef ELIM_ACC_SMEM
#define ELIM_ACC_SMEM 0 // 1: the persistent Schur sums live in shared memory (34 fewer doubles in registers)
#endif
#ifndef ELIM_BLOCKED
#define ELIM_BLOCKED 1 // blocked factorisation: scalar pivots inside the two diagonal 8x8 tiles, everything else on DMMAs
#endif
#ifndef ELIM_PIVOT2
#define ELIM_PIVOT2 0 // 1: 2x2 block pivots (8 dependent rounds, 321 FP64 instructions) instead of 15 scalar pivots (259)
#endif
#ifndef ELIM_BULK
#define ELIM_BULK 1 // 1: record staged by one cp.async.bulk + mbarrier per state (0: cp.async 16-byte chunks per lane)
#endif
#ifndef ELIM_SERIAL
#define ELIM_SERIAL 0 // K > 0: the DMMAs of W and of the Schur tiles are tied into K dependent chains (gaps on the FP64 pipe for the other warps' scalar instructions)
#endif
constexpr int SW_STAGE = 0, SW_CB = 2 * NE_STRIDE, SW_BAR = SW_CB + 128, SW_ACC = SW_BAR + 2;
constexpr int SW_ACC_N = ELIM_ACC_SMEM ? (6 * 64 + 5 * 8) : 0;
constexpr int SW_TOTAL = SW_ACC + SW_ACC_N;

// D[8x8] += A[8x4] B[4x8] on the FP64 tensor path; fragments (lane = 4 gid + tid):
//   a = A[gid][tid], b = B[tid][gid], (c0, c1) = C[gid][2 tid + {0,1}]
__device__ __forceinline__ void dmma(double &c0, double &c1, const double a, const double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
#if ELIM_SERIAL
// a zero that the hardware has to wait for: (high word of prev) & zm with zm = 0 at run time only
__device__ __forceinline__ double dep_zero(const double prev, const int zm) { return __hiloint2double(__double2hiint(prev) & zm, 0); }
// the operand, unchanged, but not before prev exists
__device__ __forceinline__ double dep_on(const double a, const double prev, const int zm) {
  return __hiloint2double(__double2hiint(a) ^ (__double2hiint(prev) & zm), __double2loint(a));
}
#endif
__device__ __forceinline__ void cp_async16(double *smem, const double *gmem) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- bulk asynchronous copy (TMA unit, 1-D) of one record into shared memory, completion on an mbarrier
__device__ __forceinline__ void mbar_init(double *bar, const unsigned count) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(bar);
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(sa), "r"(count) : "memory");
}
__device__ __forceinline__ void bulk_load(double *dst, const double *src, const unsigned bytes, double *bar) {
  const unsigned sd = (unsigned)__cvta_generic_to_shared(dst), sb = (unsigned)__cvta_generic_to_shared(bar);
  // generic-proxy reads of the buffer (the previous state staged there) are ordered before the async-proxy write
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sb), "r"(bytes) : "memory");
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(sd), "l"(src), "r"(bytes),
               "r"(sb)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(double *bar, const unsigned parity) {
  const unsigned sb = (unsigned)__cvta_generic_to_shared(bar);
  unsigned done;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(sb), "r"(parity)
                 : "memory");
  } while (!done);
}

// stream one record into shared memory with 16-byte cp.async chunks over the warp (ELIM_BULK == 0)
__device__ __forceinline__ void stage_record(double *dst, const double *src, int lane) {
#pragma unroll
  for (int k = 0; k < (NE_STRIDE / 2 + 31) / 32; k++) {
    const int ch = lane + 32 * k;
    if (ch < NE_STRIDE / 2)
      cp_async16(dst + 2 * ch, src + 2 * ch);
  }
}

// Running Schur state of a chain elimination (all "+S" = sums of W^T W products).
//   cy[t][u][e] = S[8t+gid][8u+2tid+e] of the last eliminated state (u = O-row tile; tile (0,1) is unused)
//   p15[t]      = sum over the chain of S[15][8t+gid]   (held by lanes tid == 3; F column 0 against F, B, g)
//   pt[i][e]    = sum over the chain of tiles (2,2)(2,3)(2,4)(3,3)(3,4)(4,4): S[8ta+gid][8tb+2tid+e]
// With ELIM_ACC_SMEM the two persistent sums live in the warp's shared memory ([i][lane] double2, [t][gid]).
struct ElimAcc {
  double cy[5][2][2];
#if ELIM_ACC_SMEM
  double *sm;
#else
  double p15[5];
  double pt[6][2];
#endif
};
__device__ __forceinline__ void acc_zero(ElimAcc &A, double *sm_acc, const int lane) {
#pragma unroll
  for (int t = 0; t < 5; t++)
#pragma unroll
    for (int u = 0; u < 2; u++)
      A.cy[t][u][0] = A.cy[t][u][1] = 0.0;
#if ELIM_ACC_SMEM
  A.sm = sm_acc;
#pragma unroll
  for (int i = 0; i < 6; i++)
    reinterpret_cast<double2 *>(sm_acc)[32 * i + lane] = make_double2(0.0, 0.0);
  for (int k = lane; k < 40; k += 32)
    sm_acc[384 + k] = 0.0;
  __syncwarp();
#else
#pragma unroll
  for (int t = 0; t < 5; t++)
    A.p15[t] = 0.0;
#pragma unroll
  for (int i = 0; i < 6; i++)
    A.pt[i][0] = A.pt[i][1] = 0.0;
#endif
}
__device__ __forceinline__ double acc_p15(const ElimAcc &A, const int t, const int gid) {
#if ELIM_ACC_SMEM
  return A.sm[384 + 8 * t + gid];
#else
  return A.p15[t];
#endif
}
__device__ __forceinline__ double2 acc_pt(const ElimAcc &A, const int i, const int lane) {
#if ELIM_ACC_SMEM
  return reinterpret_cast<const double2 *>(A.sm)[32 * i + lane];
#else
  return make_double2(A.pt[i][0], A.pt[i][1]);
#endif
}

// 1/x without the slow path of the IEEE division (x is a positive, normal pivot; anything else is caught by the
// positivity check of the pivots): MUFU.RCP64H seed r0 (relative error e, |e| < 2^-20), then the third-order
// correction r = r0 (1 + e + e^2) (error e^3): rcp_seed gives r0 and p = e + e^2, so that a product c / x is
// fma(c r0, p, c r0) -- c r0 does not wait for the correction, the dependent chain is seed, e, p, product
__device__ __forceinline__ void rcp_seed(const double x, double &r0, double &p) {
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  const double e = fma(-x, r0, 1.0);
  p = fma(e, e, e);
}
__device__ __forceinline__ double fast_rsqrt(const double x) {
  // MUFU.RSQ64H seed r0 (relative error < 2^-20), e = 1 - x r0^2, then the third-order correction
  // 1/sqrt(1 - e) = 1 + e/2 + 3 e^2/8 + O(e^3): five dependent-light FP64 instructions instead of eleven
  double r0;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(x));
  const double y = x * r0;
  const double e = fma(-y, r0, 1.0);
  const double t = fma(0.375, e, 0.5) * e;
  return fma(r0, t, r0);
}

// predicated shared-memory stores (one STS under a predicate: no branch, and the lanes that do not hold the value
// generate no shared-memory traffic -- the select-an-address form stored from all 32 lanes, ~4 wavefronts per store)
__device__ __forceinline__ void sts_if(const bool p, double *dst, const double v) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.f64 [%0], %1;\n\t}" ::"r"(sa), "d"(v), "r"((unsigned)p) : "memory");
}
__device__ __forceinline__ void sts2_if(const bool p, double *dst, const double v0, const double v1) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %3, 0;\n\t@q st.shared.v2.f64 [%0], {%1, %2};\n\t}" ::"r"(sa), "d"(v0), "d"(v1),
               "r"((unsigned)p)
               : "memory");
}
__device__ __forceinline__ double shfl_d(const double v, const int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double2 lds2(const double *p) { return *reinterpret_cast<const double2 *>(p); }
// sign flip on the integer pipe: the FP64 pipe is the contended one, a DADD with a negated operand
// costs a warp about as much as a DMMA there
__device__ __forceinline__ double neg_d(const double x) { return __hiloint2double(__double2hiint(x) ^ 0x80000000, __double2loint(x)); }

// LDL^T of the 15x15 block held as the tiles d00 = D[gid][2tid+e], d10 = D[8+gid][2tid+e],
// d11 = D[8+gid][8+2tid+e] together with the Gauss-Jordan inverse of its unit-lower factor; returns
// L^-1 = D^-1/2 L'^-1 (L = Cholesky factor) in the same tile order.  Per pivot j every lane needs the pivot, the
// entries of column j of D in its own row and in its two columns (D is symmetric), and row j of X in its two columns:
// the holders publish column and row through cb (2 x 32 doubles of shared memory, double buffered) with predicated
// stores.  Fifteen scalar pivots: 259 FP64 instructions per state block, against 321 for the 2x2 block-pivot form
// below (ELIM_PIVOT2) -- under load every scalar FP64 instruction costs a warp about as much time as a DMMA (it queues
// behind the other warps' DMMAs on the one FP64 pipe of the sub-partition), so the instruction count decides.
// Blocked form of the same factorisation (ELIM_BLOCKED): D = [A B^T; B C] in its 8x8 tiles,
//   L00 = chol(A)                 eight scalar pivots on tile (0,0) alone, with the Gauss-Jordan inverse of that tile
//   L10 = B L00^-T                2 DMMAs  (tile products T V^T come straight from the fragment registers: the A operand
//   S   = C - L10 L10^T           2 DMMAs   of a k-step is register e of T, the B operand register e of V)
//   L11 = chol(S)                 seven scalar pivots on tile (1,1) alone
//   (L^-1)10 = -L11^-1 L10 L00^-1   4 DMMAs and two 8x8 tile transposes (warp shuffles)
// i.e. the updates of the off-diagonal tile and of the trailing tile, which the scalar form carries through every one of
// the first eight pivots, become eight DMMAs: 142 scalar FP64 instructions per state block instead of 259.  Under load a
// scalar FP64 instruction costs a warp about as much time as a DMMA (both queue on the one FP64 pipe of the
// sub-partition), so the instruction count is what decides.
#if ELIM_BLOCKED
// transpose of an 8x8 tile held as T[gid][2tid+e]: element (2tid+e, gid) sits on lane (2tid+e, gid >> 1), register gid & 1
__device__ __forceinline__ void tile_transpose(const double (&t)[2], double (&o)[2], const int gid, const int tid) {
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int src = 4 * (2 * tid + e) + (gid >> 1);
    const double v0 = shfl_d(t[0], src), v1 = shfl_d(t[1], src);
    o[e] = (gid & 1) ? v1 : v0;
  }
}
// scalar pivots on ONE 8x8 tile d (symmetric, full) with the Gauss-Jordan inverse x of its unit-lower factor; n pivots;
// returns the pivot of the lane's row in p
__device__ __forceinline__ void tile_ldl(double (&d)[2], double (&x)[2], double &p, const int n, double *cb, const int gid, const int tid) {
  x[0] = (gid == 2 * tid) ? 1.0 : 0.0;
  x[1] = (gid == 2 * tid + 1) ? 1.0 : 0.0;
  p = 1.0;
#pragma unroll
  for (int j = 0; j < 8; j++) {
    if (j < n) {
      double *buf = cb + 32 * (j & 1);
      const int h = j >> 1, e = j & 1;
      sts_if(tid == h, buf + gid, d[e]);             // column j of the tile
      sts2_if(gid == j, buf + 8 + 2 * tid, x[0], x[1]); // row j of X
      __syncwarp();
      const double piv = buf[j];
      double r0, pc;
      rcp_seed(piv, r0, pc);
      p = (gid == j) ? piv : p;
      const double a0 = (gid > j) ? buf[gid] * r0 : 0.0;
      const double mr = fma(a0, pc, a0);
      const double2 c = lds2(buf + 2 * tid), xr = lds2(buf + 8 + 2 * tid);
      d[0] = fma(-mr, c.x, d[0]);
      d[1] = fma(-mr, c.y, d[1]);
      x[0] = fma(-mr, xr.x, x[0]);
      x[1] = fma(-mr, xr.y, x[1]);
    }
  }
}
__device__ __forceinline__ bool ldl_inverse(double (&d00)[2], double (&d10)[2], double (&d11)[2], double (&li00)[2],
                                            double (&li10)[2], double (&li11)[2], double *cb, const int gid, const int tid) {
  double x00[2], x11[2], p0, p1;
  tile_ldl(d00, x00, p0, 8, cb, gid, tid);
  const double s0 = fast_rsqrt(p0);
  li00[0] = s0 * x00[0];
  li00[1] = s0 * x00[1];
  // L10 = B L00^-T, S = C - L10 L10^T
  double g[2] = {0.0, 0.0};
  dmma(g[0], g[1], d10[0], li00[0]);
  dmma(g[0], g[1], d10[1], li00[1]);
  dmma(d11[0], d11[1], g[0], neg_d(g[0])); // accumulated straight onto C with one operand negated
  dmma(d11[0], d11[1], g[1], neg_d(g[1]));
  __syncwarp(); // the broadcast buffers of the first tile are free again
  tile_ldl(d11, x11, p1, 7, cb, gid, tid);
  const bool ok = __all_sync(0xffffffffu, (p0 > 0.0) && ((gid == 7) || (p1 > 0.0)));
  const double s1 = (gid < 7) ? fast_rsqrt(p1) : 0.0; // row 15 is padding
  li11[0] = s1 * x11[0];
  li11[1] = s1 * x11[1];
  // (L^-1)10 = -(L11^-1 L10) L00^-1
  double gt[2], lt[2], u[2] = {0.0, 0.0}, nn[2] = {0.0, 0.0};
  tile_transpose(g, gt, gid, tid);
  tile_transpose(li00, lt, gid, tid);
  dmma(u[0], u[1], li11[0], gt[0]);
  dmma(u[0], u[1], li11[1], gt[1]);
  dmma(nn[0], nn[1], u[0], lt[0]);
  dmma(nn[0], nn[1], u[1], lt[1]);
  li10[0] = neg_d(nn[0]);
  li10[1] = neg_d(nn[1]);
  return ok;
}
#elif !ELIM_PIVOT2
__device__ __forceinline__ bool ldl_inverse(double (&d00)[2], double (&d10)[2], double (&d11)[2], double (&li00)[2],
                                            double (&li10)[2], double (&li11)[2], double *cb, const int gid, const int tid) {
  double x00[2], x10[2], x11[2];
  x00[0] = (gid == 2 * tid) ? 1.0 : 0.0;
  x00[1] = (gid == 2 * tid + 1) ? 1.0 : 0.0;
  x10[0] = x10[1] = 0.0;
  x11[0] = x00[0];
  x11[1] = x00[1];
  double p0 = 1.0, p1 = 1.0;
#pragma unroll
  for (int j = 0; j < 15; j++) {
    const int jj = j & 7, h = jj >> 1, e = jj & 1; // holders of column j: lanes with tid == h, register e
    double piv, ca_g = 0.0, cb_g, ca0 = 0.0, ca1 = 0.0, cb0, cb1, xr0, xr1, xs0 = 0.0, xs1 = 0.0;
    double *buf = cb + 32 * (j & 1);
    // the holders publish column j of D (rows 0..15) and row j of X (columns 0..15)
    if (j < 8) {
      const bool hc = (tid == h), hr = (gid == j);
      sts_if(hc, buf + gid, d00[e]);
      sts_if(hc, buf + 8 + gid, d10[e]);
      sts2_if(hr, buf + 16 + 2 * tid, x00[0], x00[1]);
    } else {
      const bool hc = (tid == h), hr = (gid == jj);
      sts_if(hc, buf + 8 + gid, d11[e]);
      sts2_if(hr, buf + 16 + 2 * tid, x10[0], x10[1]);
      sts2_if(hr, buf + 24 + 2 * tid, x11[0], x11[1]);
    }
    __syncwarp();
    piv = buf[j];
    cb_g = buf[8 + gid];
    cb0 = buf[8 + 2 * tid];
    cb1 = buf[9 + 2 * tid];
    xr0 = buf[16 + 2 * tid];
    xr1 = buf[17 + 2 * tid];
    if (j < 8) {
      ca_g = buf[gid];
      ca0 = buf[2 * tid];
      ca1 = buf[2 * tid + 1];
    } else {
      xs0 = buf[24 + 2 * tid];
      xs1 = buf[25 + 2 * tid];
    }
    double r0, pc;
    rcp_seed(piv, r0, pc);
    if (j < 8) {
      p0 = (gid == j) ? piv : p0;
      const double a0 = (gid > j) ? ca_g * r0 : 0.0, a1 = cb_g * r0;
      const double mr0 = fma(a0, pc, a0), mr1 = fma(a1, pc, a1);
      d00[0] = fma(-mr0, ca0, d00[0]);
      d00[1] = fma(-mr0, ca1, d00[1]);
      d10[0] = fma(-mr1, ca0, d10[0]);
      d10[1] = fma(-mr1, ca1, d10[1]);
      d11[0] = fma(-mr1, cb0, d11[0]);
      d11[1] = fma(-mr1, cb1, d11[1]);
      x00[0] = fma(-mr0, xr0, x00[0]);
      x00[1] = fma(-mr0, xr1, x00[1]);
      x10[0] = fma(-mr1, xr0, x10[0]);
      x10[1] = fma(-mr1, xr1, x10[1]);
    } else {
      p1 = (8 + gid == j) ? piv : p1;
      const double a1 = (8 + gid > j) ? cb_g * r0 : 0.0;
      const double mr1 = fma(a1, pc, a1);
      d11[0] = fma(-mr1, cb0, d11[0]);
      d11[1] = fma(-mr1, cb1, d11[1]);
      x10[0] = fma(-mr1, xr0, x10[0]);
      x10[1] = fma(-mr1, xr1, x10[1]);
      x11[0] = fma(-mr1, xs0, x11[0]);
      x11[1] = fma(-mr1, xs1, x11[1]);
    }
  }
  // every pivot is held by exactly one (gid, *) group: positive-definiteness check in one vote
  const bool ok = __all_sync(0xffffffffu, (p0 > 0.0) && (p1 > 0.0));
  const double s0 = fast_rsqrt(p0);
  const double s1 = (gid < 7) ? fast_rsqrt(p1) : 0.0; // row 15 is padding
  li00[0] = s0 * x00[0];
  li00[1] = s0 * x00[1];
  li10[0] = s1 * x10[0];
  li10[1] = s1 * x10[1];
  li11[0] = s1 * x11[0];
  li11[1] = s1 * x11[1];
  return ok;
}
#else
// Block LDL^T of the 15x15 block (padded to 16 by a unit pivot) with 2x2 pivot blocks, held as the tiles
// d00 = D[gid][2tid+e], d10 = D[8+gid][2tid+e], d11 = D[8+gid][8+2tid+e], carried out together with the Gauss-Jordan
// inverse X of its unit block-lower factor:  D = Lb B Lb^T, B = blockdiag(P_0 .. P_7), P_k = [a b; b c].
// Returns L^-1 = C^-1 Lb^-1 in the same tile order, where P_k = C_k C_k^T is the 2x2 Cholesky factor of every pivot
// block -- C^-1 Lb^-1 is lower triangular with Gram matrix D^-1, i.e. the inverse of THE Cholesky factor of D.
// Eight dependent rounds instead of fifteen: per round the lanes holding columns 2k, 2k+1 of D and rows 2k, 2k+1 of X
// publish them through cb (2 x 64 doubles of shared memory, double buffered), every lane forms the determinant, ONE
// reciprocal (rcp_seed) and its two multipliers per row, and applies the rank-2 update.  The fixed cost of a round
// (shared-memory round trip, reciprocal) is what a state's elimination mostly waits for; it is paid 8 times, not 15.
__device__ __forceinline__ bool ldl_inverse(double (&d00)[2], double (&d10)[2], double (&d11)[2], double (&li00)[2],
                                            double (&li10)[2], double (&li11)[2], double *cb, const int gid, const int tid) {
  double x00[2], x10[2], x11[2];
  x00[0] = (gid == 2 * tid) ? 1.0 : 0.0;
  x00[1] = (gid == 2 * tid + 1) ? 1.0 : 0.0;
  x10[0] = x10[1] = 0.0;
  x11[0] = x00[0];
  x11[1] = x00[1];
  // pivot block of the lane's own rows gid (tile 0) and 8 + gid (tile 1): a, b, det
  double sa0 = 1.0, sb0 = 0.0, sd0 = 1.0, sa1 = 1.0, sb1 = 0.0, sd1 = 1.0;
#pragma unroll
  for (int k = 0; k < 8; k++) {
    double *col = cb + 64 * (k & 1); // [16][2]: (D[r][p], D[r][q]) of the pivot columns p = 2k, q = 2k + 1
    double *xb = col + 32;           // [2][16]: rows p, q of X
    const int kk = k & 3;
    const bool hc = (tid == kk);
    if (k < 4) {
      const bool hp = (gid == 2 * kk), hq = (gid == 2 * kk + 1);
      sts2_if(hc, col + 2 * gid, d00[0], d00[1]);
      sts2_if(hc, col + 2 * (8 + gid), d10[0], d10[1]);
      sts2_if(hp, xb + 2 * tid, x00[0], x00[1]);
      sts2_if(hq, xb + 16 + 2 * tid, x00[0], x00[1]);
    } else {
      const bool hp = (gid == 2 * kk), hq = (gid == 2 * kk + 1);
      sts2_if(hc, col + 2 * (8 + gid), d11[0], d11[1]);
      sts2_if(hp, xb + 2 * tid, x10[0], x10[1]);
      sts2_if(hp, xb + 8 + 2 * tid, x11[0], x11[1]);
      sts2_if(hq, xb + 16 + 2 * tid, x10[0], x10[1]);
      sts2_if(hq, xb + 24 + 2 * tid, x11[0], x11[1]);
    }
    __syncwarp();
    const int p = 2 * k;
    const double2 cp = lds2(col + 2 * p), cq = lds2(col + 2 * p + 2); // (a, b), (b, c)
    const double a = cp.x, b = cq.x, c = cq.y;
    const double det = fma(a, c, -(b * b));
    double r0, pc;
    rcp_seed(det, r0, pc);
    const double2 r1 = lds2(col + 2 * (8 + gid)); // (D[8+gid][p], D[8+gid][q])
    // multipliers of a row (u, v) = (D[g][p], D[g][q]):  (m_p, m_q) = (u, v) P^-1 = (u c - v b, v a - u b) / det
    const bool act1 = (k < 4) || (gid > 2 * kk + 1);
    const double t1p = fma(r1.x, c, -(r1.y * b)) * r0, t1q = fma(r1.y, a, -(r1.x * b)) * r0;
    const double m1p = act1 ? fma(t1p, pc, t1p) : 0.0, m1q = act1 ? fma(t1q, pc, t1q) : 0.0;
    const double2 b0 = lds2(col + 2 * (8 + 2 * tid)), b1 = lds2(col + 2 * (9 + 2 * tid)); // columns 8+2tid+e: (D[.][p], D[.][q])
    d11[0] = fma(-m1p, b0.x, fma(-m1q, b0.y, d11[0]));
    d11[1] = fma(-m1p, b1.x, fma(-m1q, b1.y, d11[1]));
    const double2 xp0 = lds2(xb + 2 * tid), xq0 = lds2(xb + 16 + 2 * tid); // X[p][2tid+e], X[q][2tid+e]
    x10[0] = fma(-m1p, xp0.x, fma(-m1q, xq0.x, x10[0]));
    x10[1] = fma(-m1p, xp0.y, fma(-m1q, xq0.y, x10[1]));
    if (k < 4) {
      const double2 r0w = lds2(col + 2 * gid); // (D[gid][p], D[gid][q])
      const bool act0 = gid > 2 * kk + 1;
      const double t0p = fma(r0w.x, c, -(r0w.y * b)) * r0, t0q = fma(r0w.y, a, -(r0w.x * b)) * r0;
      const double m0p = act0 ? fma(t0p, pc, t0p) : 0.0, m0q = act0 ? fma(t0q, pc, t0q) : 0.0;
      const double2 a0 = lds2(col + 2 * (2 * tid)), a1 = lds2(col + 2 * (2 * tid + 1)); // columns 2tid+e
      d00[0] = fma(-m0p, a0.x, fma(-m0q, a0.y, d00[0]));
      d00[1] = fma(-m0p, a1.x, fma(-m0q, a1.y, d00[1]));
      d10[0] = fma(-m1p, a0.x, fma(-m1q, a0.y, d10[0]));
      d10[1] = fma(-m1p, a1.x, fma(-m1q, a1.y, d10[1]));
      x00[0] = fma(-m0p, xp0.x, fma(-m0q, xq0.x, x00[0]));
      x00[1] = fma(-m0p, xp0.y, fma(-m0q, xq0.y, x00[1]));
      const bool mine = (gid >> 1) == kk;
      sa0 = mine ? a : sa0;
      sb0 = mine ? b : sb0;
      sd0 = mine ? det : sd0;
    } else {
      const double2 xp1 = lds2(xb + 8 + 2 * tid), xq1 = lds2(xb + 24 + 2 * tid); // X[p][8+2tid+e], X[q][8+2tid+e]
      x11[0] = fma(-m1p, xp1.x, fma(-m1q, xq1.x, x11[0]));
      x11[1] = fma(-m1p, xp1.y, fma(-m1q, xq1.y, x11[1]));
      const bool mine = (gid >> 1) == kk;
      sa1 = mine ? a : sa1;
      sb1 = mine ? b : sb1;
      sd1 = mine ? det : sd1;
    }
  }
  // positive definite <=> every leading pivot a > 0 and every block determinant > 0
  const bool ok = __all_sync(0xffffffffu, (sa0 > 0.0) && (sd0 > 0.0) && (sa1 > 0.0) && (sd1 > 0.0));
  // C_k^-1 = [1/sqrt(a) 0; -(b/a)/sqrt(s) 1/sqrt(s)], s = det / a: the even row of a block is scaled, the odd row first
  // takes -(b/a) times the even row (held by the lanes of gid - 1)
  const bool odd = gid & 1;
  const int src = ((gid & ~1) << 2) | tid; // lane holding the even row of the block, same columns
  double ra0, pa0, ra1, pa1;
  rcp_seed(sa0, ra0, pa0);
  rcp_seed(sa1, ra1, pa1);
  ra0 = fma(ra0, pa0, ra0);
  ra1 = fma(ra1, pa1, ra1);
  const double f0 = sb0 * ra0, f1 = sb1 * ra1;
  const double s0 = fast_rsqrt(odd ? sd0 * ra0 : sa0);
  const double s1 = (gid < 7) ? fast_rsqrt(odd ? sd1 * ra1 : sa1) : 0.0; // row 15 is padding
  const double e00 = shfl_d(x00[0], src), e01 = shfl_d(x00[1], src);
  const double e10 = shfl_d(x10[0], src), e11 = shfl_d(x10[1], src);
  const double e20 = shfl_d(x11[0], src), e21 = shfl_d(x11[1], src);
  li00[0] = s0 * (odd ? fma(-f0, e00, x00[0]) : x00[0]);
  li00[1] = s0 * (odd ? fma(-f0, e01, x00[1]) : x00[1]);
  li10[0] = s1 * (odd ? fma(-f1, e10, x10[0]) : x10[0]);
  li10[1] = s1 * (odd ? fma(-f1, e11, x10[1]) : x10[1]);
  li11[0] = s1 * (odd ? fma(-f1, e20, x11[0]) : x11[0]);
  li11[1] = s1 * (odd ? fma(-f1, e21, x11[1]) : x11[1]);
  return ok;
}
#endif // ELIM_PIVOT2

// Eliminate one state.  st: its staged normal-equation record (shared memory); Finit: for the first state of a
// segment with a left separator, the record of that separator (global; F = O_sep^T), else nullptr.
// On return A.cy holds this state's Schur tiles and the persistent sums are updated; the factor record is stored.
template <bool HAS_F>
__device__ __forceinline__ bool elim_node(const double *st, const double *__restrict__ Finit, const double lambda, ElimAcc &A,
                                          double *cb, double *__restrict__ fac, const int lane) {
  const int gid = lane >> 2, tid = lane & 3;
  // ---- D = record + lambda I - carry, as tiles (the record holds the lower triangle: element (r, c) at tri(max, min))
  double d00[2], d10[2], d11[2];
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int c = 2 * tid + e;
    d00[e] = st[NE_D + (gid >= c ? tri(gid, c) : tri(c, gid))] - A.cy[0][0][e] + ((gid == c) ? lambda : 0.0);
    d10[e] = (gid < 7) ? st[NE_D + tri(8 + gid, c)] - A.cy[1][0][e] : 0.0;
    const int c1 = 8 + c;
    if (gid < 7 && c1 < 15)
      d11[e] = st[NE_D + (gid >= c ? tri(8 + gid, c1) : tri(c1, 8 + gid))] - A.cy[1][1][e] + ((gid == c) ? lambda : 0.0);
    else
      d11[e] = (gid == 7 && c1 == 15) ? 1.0 : 0.0;
  }
  // ---- M = [O | F | B | g] (16x40, row 15 zero) in fragment order m[kk][t] = M[8u+2tid+e][8t+gid], kk = 2u+e
  double m[4][5];
#pragma unroll
  for (int kk = 0; kk < 4; kk++) {
    const int u = kk >> 1, e = kk & 1;
    const int r = 8 * u + 2 * tid + e;
    const bool rv = r < 15;
    const double *orow = st + NE_O + 15 * (rv ? r : 0);
    const double *brow = st + NE_BG + NE_BGW * (rv ? r : 0);
    m[kk][0] = rv ? orow[gid] : 0.0;
    double v1, v2, v3;
    if (HAS_F) {
      if (Finit) {
        const double *fi = Finit + NE_O + (rv ? r : 0);
        v1 = __ldg(fi);                     // F[r][0]     = O_sep[0][r]
        v2 = __ldg(fi + 15 * (1 + gid));    // F[r][1+gid]
        v3 = (gid < 6) ? __ldg(fi + 15 * (9 + gid)) : 0.0;
      } else {
        v1 = neg_d(A.cy[1][u][e]);
        v2 = neg_d(A.cy[2][u][e]);
        v3 = neg_d(A.cy[3][u][e]);
      }
    } else {
      v1 = v2 = v3 = 0.0;
    }
    m[kk][1] = rv ? ((gid < 7) ? orow[8 + gid] : v1) : 0.0;
    m[kk][2] = rv ? v2 : 0.0;
    m[kk][3] = rv ? ((gid < 6) ? v3 : brow[gid - 6] - A.cy[3][u][e]) : 0.0;
    m[kk][4] = rv ? brow[2 + gid] - A.cy[4][u][e] : 0.0;
  }
  __syncwarp(); // every lane has consumed the staged record and the carry
  // ---- factorisation + inverse
  double li00[2], li10[2], li11[2];
  const bool ok = ldl_inverse(d00, d10, d11, li00, li10, li11, cb, gid, tid);
  // ---- W^T = M^T L^-T  (row tile u of W from the kk-steps its triangle reaches)
  double w[4][5];
#if ELIM_SERIAL
  const int zm = (int)blockIdx.z; // 0 at run time, unknown to the compiler
  double tok[ELIM_SERIAL];
#pragma unroll
  for (int q = 0; q < ELIM_SERIAL; q++)
    tok[q] = li11[0];
  int nch = 0;
#endif
#pragma unroll
  for (int t = 0; t < 5; t++) {
    if (!HAS_F && t == 2) {
      w[0][t] = w[1][t] = w[2][t] = w[3][t] = 0.0;
      continue;
    }
#if ELIM_SERIAL
    double a0 = dep_zero(tok[t % ELIM_SERIAL], zm), a1 = 0.0, b0 = 0.0, b1 = 0.0;
    dmma(a0, a1, m[0][t], li00[0]);
    dmma(a0, a1, m[1][t], li00[1]);
    b0 = dep_zero(a0, zm);
    dmma(b0, b1, m[0][t], li10[0]);
    dmma(b0, b1, m[1][t], li10[1]);
    dmma(b0, b1, m[2][t], li11[0]);
    dmma(b0, b1, m[3][t], li11[1]);
    tok[t % ELIM_SERIAL] = b0;
#else
    double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
    dmma(a0, a1, m[0][t], li00[0]);
    dmma(b0, b1, m[0][t], li10[0]);
    dmma(a0, a1, m[1][t], li00[1]);
    dmma(b0, b1, m[1][t], li10[1]);
    dmma(b0, b1, m[2][t], li11[0]);
    dmma(b0, b1, m[3][t], li11[1]);
#endif
    w[0][t] = a0;
    w[1][t] = a1;
    w[2][t] = b0;
    w[3][t] = b1;
  }
  // ---- factor record (fragment order without the structurally zero fragments; coalesced 16- and 8-byte stores)
  {
    const bool lo = (2 * tid <= gid);
    if (lo)
      *reinterpret_cast<double2 *>(fac + FAC_LI00 + 2 * tri_slot(gid, tid)) = make_double2(li00[0], li00[1]);
    if (gid < 7) {
      *reinterpret_cast<double2 *>(fac + FAC_LI10 + 2 * lane) = make_double2(li10[0], li10[1]);
      if (lo)
        *reinterpret_cast<double2 *>(fac + FAC_LI11 + 2 * tri_slot(gid, tid)) = make_double2(li11[0], li11[1]);
    }
#pragma unroll
    for (int t = 0; t < 5; t++) {
      *reinterpret_cast<double2 *>(fac + FAC_W0 + 64 * t + 2 * lane) = make_double2(w[0][t], w[1][t]);
      fac[FAC_W1 + 56 * t + lane] = w[2][t];
      if (tid < 3)
        fac[FAC_W1 + 56 * t + 32 + 3 * gid + tid] = w[3][t];
    }
  }
  // ---- Schur tiles S = W^T W
#pragma unroll
  for (int t = 0; t < 5; t++)
#pragma unroll
    for (int u = 0; u < 2; u++) {
      double c0 = 0.0, c1 = 0.0;
      if (!(t == 0 && u == 1) && (HAS_F || t != 2)) {
#if ELIM_SERIAL
        c0 = dep_zero(tok[nch % ELIM_SERIAL], zm);
#endif
#pragma unroll
        for (int kk = 0; kk < 4; kk++)
          dmma(c0, c1, w[kk][t], w[kk][u]);
#if ELIM_SERIAL
        tok[nch % ELIM_SERIAL] = c0;
        nch++;
#endif
      }
      A.cy[t][u][0] = c0;
      A.cy[t][u][1] = c1;
    }
  if (HAS_F) {
#if ELIM_ACC_SMEM
    if (tid == 3) {
#pragma unroll
      for (int t = 1; t < 5; t++)
        A.sm[384 + 8 * t + gid] += A.cy[t][1][1];
    }
#else
#pragma unroll
    for (int t = 1; t < 5; t++)
      A.p15[t] += A.cy[t][1][1];
#endif
  }
  {
    int i = 0;
#pragma unroll
    for (int ta = 2; ta < 5; ta++)
#pragma unroll
      for (int tb = ta; tb < 5; tb++) {
        if (HAS_F || ta != 2) {
#if ELIM_ACC_SMEM
          double c0 = 0.0, c1 = 0.0;
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            dmma(c0, c1, w[kk][ta], w[kk][tb]);
          double2 *q = reinterpret_cast<double2 *>(A.sm) + 32 * i + lane;
          const double2 o = *q;
          *q = make_double2(o.x + c0, o.y + c1);
#else
#if ELIM_SERIAL
          dmma(A.pt[i][0], A.pt[i][1], dep_on(w[0][ta], tok[nch % ELIM_SERIAL], zm), w[0][tb]);
#pragma unroll
          for (int kk = 1; kk < 4; kk++)
            dmma(A.pt[i][0], A.pt[i][1], w[kk][ta], w[kk][tb]);
          tok[nch % ELIM_SERIAL] = A.pt[i][0];
          nch++;
#else
#pragma unroll
          for (int kk = 0; kk < 4; kk++)
            dmma(A.pt[i][0], A.pt[i][1], w[kk][ta], w[kk][tb]);
#endif
#endif
        }
        i++;
      }
  }
  return ok;
}

// Eliminate `count` consecutive states whose records start at ne0 (stride NE_STRIDE); factor records to fac0.
// w: the warp's shared memory; the state's record is fetched one state ahead.
template <bool HAS_F>
__device__ __forceinline__ bool elim_chain(const double *__restrict__ ne0, const int count, const double *__restrict__ left_sep_ne,
                                           const double lambda, double *__restrict__ fac0, ElimAcc &A, double *w, const int lane) {
  double *stage = w + SW_STAGE;
  double *cb = w + SW_CB;
  bool ok = true;
#if ELIM_BULK
  double *bar = w + SW_BAR;
  constexpr unsigned REC_BYTES = NE_STRIDE * sizeof(double);
  if (lane == 0) {
    mbar_init(bar, 1);
    mbar_init(bar + 1, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();
  if (count > 0 && lane == 0)
    bulk_load(stage, ne0, REC_BYTES, bar);
  for (int k = 0; k < count; k++) {
    // the buffer of state k + 1 was read by state k - 1, whose lanes have all passed the __syncwarp of elim_node
    if (k + 1 < count && lane == 0)
      bulk_load(stage + NE_STRIDE * ((k + 1) & 1), ne0 + (long long)(k + 1) * NE_STRIDE, REC_BYTES, bar + ((k + 1) & 1));
    mbar_wait(bar + (k & 1), (unsigned)(k >> 1) & 1u);
    ok = elim_node<HAS_F>(stage + NE_STRIDE * (k & 1), (HAS_F && k == 0) ? left_sep_ne : nullptr, lambda, A, cb,
                          fac0 + (long long)k * FAC_STRIDE, lane) && ok;
  }
#else
  if (count > 0)
    stage_record(stage, ne0, lane);
  cp_async_commit();
  for (int k = 0; k < count; k++) {
    if (k + 1 < count)
      stage_record(stage + NE_STRIDE * ((k + 1) & 1), ne0 + (long long)(k + 1) * NE_STRIDE, lane);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    ok = elim_node<HAS_F>(stage + NE_STRIDE * (k & 1), (HAS_F && k == 0) ? left_sep_ne : nullptr, lambda, A, cb,
                          fac0 + (long long)k * FAC_STRIDE, lane) && ok;
  }
  cp_async_wait<0>();
#endif
  return ok;
}

// ------------------------------------------------------------------------------------------------
// grid: (ceil(P / ELIM_WARPS), T); block: ELIM_WARPS warps; warp -> segment j of trial t
__global__ void __launch_bounds__(ELIM_WARPS * 32, ELIM_MIN_CTAS) k_seg_eliminate(DevPtrs d, int force_all) {
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.y;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int j = blockIdx.x * ELIM_WARPS + wid + (d.chain ? d.c_j0 : 0);
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  if (j >= (d.chain ? min(sg.nseg, d.c_j1) : sg.nseg))
    return;
  double *w = smem + (size_t)wid * SW_TOTAL;
  const TrialDev &tr = d.trials[t];
  const double lambda = lm.lambda;
  const int a0 = j * sg.stride;
  const int b1 = min(n, (j + 1) * sg.stride - 1); // exclusive end of the interior
  const bool has_left = (j >= 1);
  const long long s0 = tr.st_off + a0 - (d.chain ? d.c_goff : 0);
  ElimAcc A;
  acc_zero(A, w + SW_ACC, lane);
  bool ok;
  if (has_left)
    ok = elim_chain<true>(d.ne + s0 * NE_STRIDE, b1 - a0, d.ne + (s0 - 1) * NE_STRIDE, lambda, d.fac + s0 * FAC_STRIDE, A, w, lane);
  else
    ok = elim_chain<false>(d.ne + s0 * NE_STRIDE, b1 - a0, nullptr, lambda, d.fac + s0 * FAC_STRIDE, A, w, lane);
  // ---- segment outputs
  const int gid = lane >> 2, tid = lane & 3;
  double *so = d.seg + ((long long)t * d.P + j) * SEG_STRIDE;
  double *cyo = so + SEG_CY, *fbo = so + SEG_FB;
#pragma unroll
  for (int tt = 0; tt < 5; tt++)
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int e = 0; e < 2; e++) {
        if (tt == 0 && u == 1)
          continue;
        const int r = 8 * u + 2 * tid + e, c = 8 * tt + gid;
        if (r < 15) {
          cyo[40 * r + c] = A.cy[tt][u][e];
          if (tt == 1 && u == 0 && c < 15)
            cyo[40 * c + r] = A.cy[tt][u][e]; // mirror of the D part
        }
      }
#if ELIM_ACC_SMEM
  __syncwarp();
#endif
  if (tid == 3) {
#pragma unroll
    for (int tt = 1; tt < 5; tt++) {
      const int b = 8 * tt + gid - 15;
      if (b >= 0) {
        const double v = acc_p15(A, tt, gid);
        fbo[b] = v;
        fbo[25 * b] = v;
      }
    }
  }
  {
    int i = 0;
#pragma unroll
    for (int ta = 2; ta < 5; ta++)
#pragma unroll
      for (int tb = ta; tb < 5; tb++) {
        const double2 v = acc_pt(A, i, lane);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int a = 8 * ta + gid - 15, b = 8 * tb + 2 * tid + e - 15;
          fbo[25 * a + b] = e ? v.y : v.x;
          fbo[25 * b + a] = e ? v.y : v.x;
        }
        i++;
      }
  }
  if (!ok && lane == 0)
    atomicExch(&lm.solve_failed, 1);
}

// ------------------------------------------------------------------------------------------------
// Normal-equation records of the separator chain.  grid (P, T), 128 threads; CTA -> separator m = blockIdx.x + 1.
__global__ void __launch_bounds__(128) k_reduced_assemble(DevPtrs d, int force_all) {
  const int t = blockIdx.y;
  const LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  const int m = blockIdx.x + 1;
  if (m > sg.nsep)
    return;
  const TrialDev &tr = d.trials[t];
  const long long s = tr.st_off + (long long)m * sg.stride - 1;
  const double *ne = d.chain ? d.c_sepne + (long long)m * NE_STRIDE : d.ne + s * NE_STRIDE;
  const double *segl = d.seg + ((long long)t * d.P + (m - 1)) * SEG_STRIDE; // this separator closes segment m-1
  const bool has_r = (m < sg.nseg);                                         // ... and opens segment m
  const double *segr = d.seg + ((long long)t * d.P + m) * SEG_STRIDE;
  const bool has_next = has_r && (m + 1 <= sg.nsep);
  double *out = d.red_ne + ((long long)t * (d.P + 1) + m) * NE_STRIDE;
  for (int idx = threadIdx.x; idx < NE_STRIDE; idx += blockDim.x) {
    double v;
    if (idx < NE_O) {
      // packed lower triangle: idx = tri(i, jn), i >= jn
      int i = (int)((sqrt(8.0 * idx + 1.0) - 1.0) * 0.5);
      while (tri(i + 1, 0) <= idx)
        i++;
      while (tri(i, 0) > idx)
        i--;
      const int jn = idx - tri(i, 0);
      v = ne[idx] - segl[SEG_CY + 40 * i + jn];
      if (has_r)
        v -= segr[SEG_FB + 25 * i + jn];
    } else if (idx < NE_BG) {
      const int q = idx - NE_O;
      const int i = q / 15, c = q - 15 * i;
      // coupling separator m -> m+1 through segment m:  H[m][m+1](i, c) = -S(rows: O index c, column F index i)
      v = has_next ? -segr[SEG_CY + 40 * c + XC_F + i] : 0.0;
    } else if (idx < NE_BG + 15 * NE_BGW) {
      const int q = idx - NE_BG;
      const int i = q / NE_BGW, k = q - NE_BGW * i; // border columns 0..8, rhs 9
      v = ne[idx] - segl[SEG_CY + 40 * i + XC_B + k];
      if (has_r)
        v -= segr[SEG_FB + 25 * i + 15 + k];
    } else {
      v = 0.0; // pad
    }
    out[idx] = v;
  }
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

// Back-substitution state of one chain: the lane's slice xv[8t+gid] of the extended vector
// [x_next (15) | x_left (15) | x_glob (9) | -1]; slices 0 and 1 follow the chain, the rest is fixed.
struct BackVec {
  double xv[5];
};
__device__ __forceinline__ void backvec_init(BackVec &V, const double *x_next, const double *x_left, const double *x_glob,
                                             const int gid) {
  V.xv[0] = x_next[gid];
  V.xv[1] = (gid < 7) ? x_next[8 + gid] : x_left[0];
  V.xv[2] = x_left[1 + gid];
  V.xv[3] = (gid < 6) ? x_left[9 + gid] : x_glob[gid - 6];
  V.xv[4] = (gid < 7) ? x_glob[2 + gid] : -1.0;
}
struct FacRegs {
  double2 li[3];
  double2 w[10];
};
// the packed factor record back into fragment registers (the fragments that are zero by structure are not stored)
__device__ __forceinline__ void fac_load(FacRegs &F, const double *__restrict__ fac, const int lane) {
  const int gid = lane >> 2, tid = lane & 3;
  const bool lo = (2 * tid <= gid);
  const double2 z = make_double2(0.0, 0.0);
  F.li[0] = lo ? __ldcs(reinterpret_cast<const double2 *>(fac + FAC_LI00) + tri_slot(gid, tid)) : z;
  F.li[1] = (gid < 7) ? __ldcs(reinterpret_cast<const double2 *>(fac + FAC_LI10) + lane) : z;
  F.li[2] = (gid < 7 && lo) ? __ldcs(reinterpret_cast<const double2 *>(fac + FAC_LI11) + tri_slot(gid, tid)) : z;
#pragma unroll
  for (int t = 0; t < 5; t++) {
    F.w[t] = __ldcs(reinterpret_cast<const double2 *>(fac + FAC_W0 + 64 * t) + lane);
    F.w[5 + t].x = __ldcs(fac + FAC_W1 + 56 * t + lane);
    F.w[5 + t].y = (tid < 3) ? __ldcs(fac + FAC_W1 + 56 * t + 32 + 3 * gid + tid) : 0.0;
  }
}
// x = L^-T (w_g - W_O x_next - W_F x_left - W_B x_glob); on return every lane holds x[2tid+e] in y0[e] and
// x[8+2tid+e] in y1[e], and V follows the chain.
__device__ __forceinline__ void backsub_node(const FacRegs &F, BackVec &V, double (&y0)[2], double (&y1)[2], const double x_left0,
                                             const int gid, const int tid) {
  double p[4];
#pragma unroll
  for (int u = 0; u < 2; u++) {
    double a = 0.0, b = 0.0;
#pragma unroll
    for (int t = 0; t < 5; t++) {
      a = fma(F.w[u * 5 + t].x, V.xv[t], a);
      b = fma(F.w[u * 5 + t].y, V.xv[t], b);
    }
    p[2 * u] = a;
    p[2 * u + 1] = b;
  }
#pragma unroll
  for (int o = 4; o < 32; o <<= 1)
#pragma unroll
    for (int k = 0; k < 4; k++)
      p[k] += __shfl_xor_sync(0xffffffffu, p[k], o);
  // rhs[gid], rhs[8+gid] live on the lanes with tid' = gid / 2, register kk = (gid & 1) (+2)
  const int src = gid >> 1;
  const double s0 = __shfl_sync(0xffffffffu, p[0], src), s1 = __shfl_sync(0xffffffffu, p[1], src);
  const double s2 = __shfl_sync(0xffffffffu, p[2], src), s3 = __shfl_sync(0xffffffffu, p[3], src);
  const double r_lo = -((gid & 1) ? s1 : s0), r_hi = -((gid & 1) ? s3 : s2);
  y0[0] = fma(F.li[0].x, r_lo, F.li[1].x * r_hi);
  y0[1] = fma(F.li[0].y, r_lo, F.li[1].y * r_hi);
  y1[0] = F.li[2].x * r_hi;
  y1[1] = F.li[2].y * r_hi;
#pragma unroll
  for (int o = 4; o < 32; o <<= 1) {
    y0[0] += __shfl_xor_sync(0xffffffffu, y0[0], o);
    y0[1] += __shfl_xor_sync(0xffffffffu, y0[1], o);
    y1[0] += __shfl_xor_sync(0xffffffffu, y1[0], o);
    y1[1] += __shfl_xor_sync(0xffffffffu, y1[1], o);
  }
  const double t0 = __shfl_sync(0xffffffffu, y0[0], src), t1 = __shfl_sync(0xffffffffu, y0[1], src);
  const double t2 = __shfl_sync(0xffffffffu, y1[0], src), t3 = __shfl_sync(0xffffffffu, y1[1], src);
  V.xv[0] = (gid & 1) ? t1 : t0;
  V.xv[1] = (gid < 7) ? ((gid & 1) ? t3 : t2) : x_left0;
}
// store x (15 values) of a state: lanes gid == 0 hold all of it between them
__device__ __forceinline__ void store_x(double *__restrict__ dst, const double (&y0)[2], const double (&y1)[2], const int gid,
                                        const int tid) {
  if (gid == 0) {
    *reinterpret_cast<double2 *>(dst + 2 * tid) = make_double2(y0[0], y0[1]);
    *reinterpret_cast<double2 *>(dst + 8 + 2 * tid) = make_double2(y1[0], (tid == 3) ? 0.0 : y1[1]);
  }
}

constexpr int RED_SMEM = SW_TOTAL + 160; // + 9x9 system (90) + x_glob (16) + x scratch (32)

// grid: T CTAs of one warp.
__global__ void __launch_bounds__(32) k_reduced(DevPtrs d, int force_all) {
  extern __shared__ __align__(16) double smem[];
  const int t = blockIdx.x;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int lane = threadIdx.x;
  const int gid = lane >> 2, tid = lane & 3;
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  const TrialDev &tr = d.trials[t];
  const double lambda = lm.lambda;
  double *w = smem;
  double *G9 = smem + SW_TOTAL;       // 9x9 + rhs 9
  double *xgs = smem + SW_TOTAL + 96; // x_glob (16)
  double *xz = smem + SW_TOTAL + 112; // zeros / x_next scratch (32)
  // ---- forward elimination over the separators m = 1..nsep
  ElimAcc A;
  acc_zero(A, w + SW_ACC, lane);
  const long long r0 = (long long)t * (d.P + 1) + 1;
  bool ok = elim_chain<false>(d.red_ne + r0 * NE_STRIDE, sg.nsep, nullptr, lambda, d.red_fac + r0 * FAC_STRIDE, A, w, lane);
  // ---- global 9x9: Gamma + lambda I - sum_seg S_BB - S_BB(separator chain); rhs likewise
  double *fbr = d.red_fb + (long long)t * 640; // 25x25 scratch in global (only the B,g rows/columns are meaningful)
  {
    int i = 0;
#pragma unroll
    for (int ta = 2; ta < 5; ta++)
#pragma unroll
      for (int tb = ta; tb < 5; tb++) {
        const double2 v = acc_pt(A, i, lane);
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const int a = 8 * ta + gid - 15, b = 8 * tb + 2 * tid + e - 15;
          fbr[25 * a + b] = e ? v.y : v.x;
          fbr[25 * b + a] = e ? v.y : v.x;
        }
        i++;
      }
  }
  __syncwarp();
  const double *segs = d.seg + (long long)t * d.P * SEG_STRIDE;
  const double *gm = d.gamma + (long long)t * 96;
  for (int idx = lane; idx < 90; idx += 32) {
    const int r = (idx < 81) ? idx / 9 : idx - 81;
    const int c = (idx < 81) ? idx - 9 * r : 9;
    double v = gm[idx] - fbr[25 * (15 + r) + 15 + c];
    if (d.gamma_sub)
      v -= d.gamma_sub[(long long)t * 96 + idx];
    for (int jj = 0; jj < sg.nseg; jj++)
      v -= segs[(long long)jj * SEG_STRIDE + SEG_FB + 25 * (15 + r) + 15 + c];
    if (idx < 81 && r == c)
      v += lambda;
    G9[idx] = v;
  }
  for (int idx = lane; idx < 48; idx += 32)
    xgs[idx] = 0.0; // x_glob + scratch zeros
  __syncwarp();
  double *Am = G9;
  if (lane == 0) {
    // symmetrise, Cholesky, solve (tiny, sequential)
    for (int r = 0; r < 9; r++)
      for (int c = 0; c < r; c++) {
        const double v = 0.5 * (Am[9 * r + c] + Am[9 * c + r]);
        Am[9 * r + c] = v;
        Am[9 * c + r] = v;
      }
    for (int jj = 0; jj < 9; jj++) {
      double dd = Am[10 * jj];
      for (int k = 0; k < jj; k++)
        dd -= Am[9 * jj + k] * Am[9 * jj + k];
      if (!(dd > 0.0))
        ok = false;
      const double l = sqrt(dd);
      Am[10 * jj] = l;
      for (int r = jj + 1; r < 9; r++) {
        double sacc = Am[9 * r + jj];
        for (int k = 0; k < jj; k++)
          sacc -= Am[9 * r + k] * Am[9 * jj + k];
        Am[9 * r + jj] = sacc / l;
      }
    }
    double y[9], x[9];
    for (int r = 0; r < 9; r++) {
      double sacc = Am[81 + r];
      for (int k = 0; k < r; k++)
        sacc -= Am[9 * r + k] * y[k];
      y[r] = sacc / Am[10 * r];
    }
    for (int r = 8; r >= 0; r--) {
      double sacc = y[r];
      for (int k = r + 1; k < 9; k++)
        sacc -= Am[9 * k + r] * x[k];
      x[r] = sacc / Am[10 * r];
    }
    double gd = 0, dd2 = 0;
    for (int r = 0; r < 9; r++) {
      xgs[r] = x[r];
      gd += gm[81 + r] * x[r];
      dd2 += x[r] * x[r];
    }
    double *xg = d.xg + (long long)t * 16;
    for (int r = 0; r < 9; r++)
      xg[r] = x[r];
    xg[10] = gd;
    xg[11] = dd2;
  }
  __syncwarp();
  if (!__all_sync(0xffffffffu, ok)) {
    if (lane == 0)
      atomicExch(&lm.solve_failed, 1);
    return;
  }
  // ---- back-substitution of the separators (right to left)
  BackVec V;
  backvec_init(V, xz, xz, xgs, gid);
  for (int m = sg.nsep; m >= 1; m--) {
    FacRegs F;
    fac_load(F, d.red_fac + ((long long)t * (d.P + 1) + m) * FAC_STRIDE, lane);
    double y0[2], y1[2];
    backsub_node(F, V, y0, y1, 0.0, gid, tid);
    store_x(d.xsep + ((long long)t * (d.P + 1) + m) * 16, y0, y1, gid, tid);
    // the separator's own step: in chain-sharded mode only where the state has a local slot (own or halo)
    const long long slot = (long long)m * sg.stride - 1 - (d.chain ? d.c_goff : 0);
    if (!d.chain || (slot >= 0 && slot < d.n_states[t]))
      store_x(d.delta + (tr.st_off + slot) * 16, y0, y1, gid, tid);
  }
  if (lane == 0) {
    // trial point of the globals
    const double *xg = d.xg + (long long)t * 16;
    const double *g0 = d.glob + ((long long)lm.cur * d.T + t) * GLOB_STRIDE;
    double *g1 = d.glob + ((long long)(1 - lm.cur) * d.T + t) * GLOB_STRIDE;
    retract_globals(g0, xg, g1);
  }
}

// ------------------------------------------------------------------------------------------------
constexpr int BS_WARPS = 4;

// grid: (ceil(P / BS_WARPS), T); warp -> segment j of trial t
__global__ void __launch_bounds__(BS_WARPS * 32) k_seg_backsub(DevPtrs d, int force_all) {
  __shared__ double s_x[BS_WARPS][48];
  const int t = blockIdx.y;
  LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK || lm.solve_failed))
    return;
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int gid = lane >> 2, tid = lane & 3;
  const int j = blockIdx.x * BS_WARPS + wid + (d.chain ? d.c_j0 : 0);
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  if (j >= (d.chain ? min(sg.nseg, d.c_j1) : sg.nseg))
    return;
  const TrialDev &tr = d.trials[t];
  const long long base = tr.st_off - (d.chain ? d.c_goff : 0); // slot of global state 0
  const int a0 = j * sg.stride;
  const int b1 = min(n, (j + 1) * sg.stride - 1);
  const bool has_left = (j >= 1);
  const bool has_right = ((j + 1) * sg.stride - 1 < n);
  double *xs = s_x[wid]; // [0,16) x_next, [16,32) x_left, [32,48) x_glob
  if (lane < 16) {
    xs[lane] = (has_right && lane < 15) ? d.xsep[((long long)t * (d.P + 1) + (j + 1)) * 16 + lane] : 0.0;
    xs[16 + lane] = (has_left && lane < 15) ? d.xsep[((long long)t * (d.P + 1) + j) * 16 + lane] : 0.0;
    xs[32 + lane] = (lane < 9) ? d.xg[(long long)t * 16 + lane] : 0.0;
  }
  __syncwarp();
  BackVec V;
  backvec_init(V, xs, xs + 16, xs + 32, gid);
  const double x_left0 = xs[16];
  if (b1 <= a0)
    return;
  FacRegs F, Fn;
  fac_load(F, d.fac + (base + b1 - 1) * FAC_STRIDE, lane);
  for (int k = b1 - 1; k >= a0; k--) {
    const long long s = base + k;
    if (k > a0)
      fac_load(Fn, d.fac + (s - 1) * FAC_STRIDE, lane);
    double y0[2], y1[2];
    backsub_node(F, V, y0, y1, x_left0, gid, tid);
    store_x(d.delta + s * 16, y0, y1, gid, tid);
    F = Fn;
  }
}

// ------------------------------------------------------------------------------------------------
int elim_warps_per_cta() { return ELIM_WARPS; }
int elim_ctas_per_sm() { return ELIM_MIN_CTAS; }
int solve_configure() {
  const size_t sm1 = sizeof(double) * SW_TOTAL * ELIM_WARPS;
  const size_t sm2 = sizeof(double) * RED_SMEM;
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_seg_eliminate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm1));
  V2_CUDA_CHECK(cudaFuncSetAttribute(k_reduced, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm2));
  return V2GT_OK;
}
int launch_eliminate(const DevPtrs &d, int force_all, cudaStream_t st) {
  const size_t sm1 = sizeof(double) * SW_TOTAL * ELIM_WARPS;
  dim3 g1((d.P + ELIM_WARPS - 1) / ELIM_WARPS, d.T);
  k_seg_eliminate<<<g1, ELIM_WARPS * 32, sm1, st>>>(d, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
// ------------------------------------------------------------------------------------------------
// Level 2 of the nested elimination.  grid T, 96 threads: the Schur terms of the level-1 segments on the global
// block (summed in segment order), and the description of the level-1 separator chain as a chain of "states".
__global__ void __launch_bounds__(96) k_l2_prepare(DevPtrs d, int force_all) {
  const int t = blockIdx.x;
  const LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK))
    return;
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  if (threadIdx.x == 0) {
    d.l2_nstates[t] = sg.nsep;
    d.l2_trials[t].st_off = (long long)t * (d.P + 1) + 1;
    d.l2_trials[t].n_raw = sg.nsep;
  }
  const int idx = threadIdx.x;
  if (idx < 90) {
    const int r = (idx < 81) ? idx / 9 : idx - 81;
    const int c = (idx < 81) ? idx - 9 * r : 9;
    const double *segs = d.seg + (long long)t * d.P * SEG_STRIDE;
    double v = 0.0;
    for (int jj = 0; jj < sg.nseg; jj++)
      v += segs[(long long)jj * SEG_STRIDE + SEG_FB + 25 * (15 + r) + 15 + c];
    d.l2_gamma_sub[(long long)t * 96 + idx] = v;
  }
}

// grid (P, T), 32 threads: step of separator m = blockIdx.x + 1 from xsep into the delta slot of its state
__global__ void __launch_bounds__(32) k_sep_delta(DevPtrs d, int force_all) {
  const int t = blockIdx.y;
  const LmState &lm = d.lm[t];
  if (!force_all && (lm.lm_state != 0 || lm.status != V2GT_OK || lm.solve_failed))
    return;
  const int n = d.chain ? d.c_n : d.n_states[t];
  const SegGeom sg = seg_geom(n, d.P);
  const int m = blockIdx.x + 1;
  if (m > sg.nsep || threadIdx.x >= 16)
    return;
  const long long slot = (long long)m * sg.stride - 1 - (d.chain ? d.c_goff : 0);
  if (!d.chain || (slot >= 0 && slot < d.n_states[t]))
    d.delta[(d.trials[t].st_off + slot) * 16 + threadIdx.x] = d.xsep[((long long)t * (d.P + 1) + m) * 16 + threadIdx.x];
}

// level-2 segment count for a level-1 chain cut into P segments (0: solve the separator chain directly)
__host__ int l2_segments(int P, int l2_max) {
  if (l2_max < 2 || P < 24)
    return 0;
  int p2 = (int)(sqrt((double)P) + 0.5);
  p2 = p2 < 2 ? 2 : p2;
  return p2 > l2_max ? l2_max : p2;
}

int launch_reduced(const DevPtrs &d, int force_all, cudaStream_t st) {
  const size_t sm1 = sizeof(double) * SW_TOTAL * ELIM_WARPS;
  const size_t sm2 = sizeof(double) * RED_SMEM;
  if (d.P > 1) {
    dim3 g0(d.P - 1, d.T);
    k_reduced_assemble<<<g0, 128, 0, st>>>(d, force_all);
  }
  const int P2 = l2_segments(d.P, d.l2_P);
  if (P2 == 0) {
    k_reduced<<<d.T, 32, sm2, st>>>(d, force_all);
    V2_CUDA_CHECK(cudaGetLastError());
    return V2GT_OK;
  }
  // ---- nested: the separator chain (records red_ne, one per separator) is eliminated like a state chain
  k_l2_prepare<<<d.T, 96, 0, st>>>(d, force_all);
  DevPtrs d2 = d;
  d2.chain = 0;
  d2.P = P2;
  d2.trials = d.l2_trials;
  d2.n_states = d.l2_nstates;
  d2.ne = d.red_ne;
  d2.fac = d.red_fac;
  d2.seg = d.l2_seg;
  d2.red_ne = d.l2_red_ne;
  d2.red_fac = d.l2_red_fac;
  d2.xsep = d.l2_xsep;
  d2.delta = d.xsep; // the "states" of level 2 are the separators of level 1
  d2.gamma_sub = d.l2_gamma_sub;
  d2.l2_P = 0;
  dim3 g1((P2 + ELIM_WARPS - 1) / ELIM_WARPS, d.T);
  k_seg_eliminate<<<g1, ELIM_WARPS * 32, sm1, st>>>(d2, force_all);
  dim3 g2(P2 - 1, d.T);
  k_reduced_assemble<<<g2, 128, 0, st>>>(d2, force_all);
  k_reduced<<<d.T, 32, sm2, st>>>(d2, force_all);
  dim3 g3((P2 + BS_WARPS - 1) / BS_WARPS, d.T);
  k_seg_backsub<<<g3, BS_WARPS * 32, 0, st>>>(d2, force_all);
  dim3 g4(d.P, d.T);
  k_sep_delta<<<g4, 32, 0, st>>>(d, force_all);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_backsub(const DevPtrs &d, int force_all, cudaStream_t st) {
  dim3 g3((d.P + BS_WARPS - 1) / BS_WARPS, d.T);
  k_seg_backsub<<<g3, BS_WARPS * 32, 0, st>>>(d, force_all);
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
