// vicon2gt_b200: shared host/device definitions for the batch solver (sm_100a, fp64).
//
// Data layout in HBM (all arrays are concatenated over the trials of a batch; a trial
// owns the slot range [st_off, st_off + n_raw) of every per-state array, of which the first
// n_states[t] slots are live after compaction):
//   imu_t   [M]        sample times (dense, for the batched binary search)
//   imu_s   [M][8]     (t, w3, a3, pad)            one 64-byte line per sample
//   vic_t   [V]        pose times (sorted, unique)
//   vic_s   [V][8]     (q4, p3, pad)               one 64-byte line per pose
//   vic_c   [V][18]    (R_q 3x3, R_p 3x3) or absent when the trial has one constant covariance
//   st_t    [S]        state times
//   x       [2][S][16] JPLNavState rows (q4 bg3 v3 ba3 p3), double buffered (current / trial point)
//   glob    [2][T][16] q_BtoI4 p_BinI3 t_off theta_x theta_y (+pad), double buffered
//   pre     [S/32][64][32]   preintegration constants of interval slot s -> s+1 (PRE_* offsets below)
//   info    [S/32][120][32]  P^-1 of that interval, packed lower triangle (row-major: r*(r+1)/2 + c)
//   fpi     [S/32][64][32]   per-factor linearisation pieces of the IMU factor (FPI_* offsets)
//   fpv     [S/32][88][32]   per-factor pieces of the vicon factor (FPV_* offsets)
//                      (these four are TILED struct-of-arrays, soa_base() below: element e of slot s at
//                       [(s / 32) * E * 32 + e * 32 + s % 32], so the one-thread-per-state kernels read and write them
//                       fully coalesced and a tile is one contiguous block)
//   ne      [S][496]   normal-equation record per state: D packed lower triangle (120) | O 15x15 (225, coupling k,k+1) |
//                      BG 15x10 = [B 9 | g 1] (150) | pad 1 -- 495 doubles is what the elimination has to read; one
//                      contiguous, 16-byte aligned record of 3968 bytes, fetched by ONE bulk asynchronous copy
//                      (cp.async.bulk + mbarrier) per state
//   grad    [S][16]    compact copy of g
//   fac     [S][728]   elimination factors per state in mma-fragment order, zero fragments and the padding row left
//                      out: L^-1 (40 + 56 + 32) | W (15x40: 320 + 280)
//   delta   [S][16]    LM step per state
#pragma once
#include "../../include/vicon2gt_b200.h"
#include "v2gt_math.cuh"
#include <cuda_runtime.h>
#include <stdint.h>

namespace v2 {

// ---- preintegration constants row (64 doubles)
constexpr int PRE_DT = 0, PRE_ALPHA = 1, PRE_BETA = 4, PRE_Q = 7, PRE_BGLIN = 11, PRE_BALIN = 14, PRE_JQ = 17, PRE_JB = 26,
              PRE_JA = 35, PRE_HB = 44, PRE_HA = 53, PRE_NSTEPS = 62, PRE_STRIDE = 64;
constexpr int INFO_STRIDE = 120;
constexpr int X_STRIDE = 16, GLOB_STRIDE = 16;
// globals row
constexpr int GL_Q = 0, GL_P = 4, GL_TOFF = 7, GL_THX = 8, GL_THY = 9;

// ---- IMU factor pieces (64 doubles): residual + the non-trivial Jacobian blocks
//  e[15]; Tth[9] = H1(0:3,0:3); Tb[9] = H1(0:3,3:6); sv[3], sp[3] (H1(6:9,0:3) = [sv]x, H1(12:15,0:3) = [sp]x);
//  Rk[9] = quat_2_Rot(q_k); Jth2[9] = H2(0:3,0:3); valid flag
constexpr int FPI_E = 0, FPI_TTH = 15, FPI_TB = 24, FPI_SV = 33, FPI_SP = 36, FPI_RK = 39, FPI_JTH2 = 48, FPI_VALID = 57,
              FPI_STRIDE = 64;
// ---- vicon factor pieces (88 doubles): Huber-reweighted whitened system
//  r[6] = sqrt(w) W e ; A[6][6] = sqrt(w) W H1(:, [0:3, 12:15]) ; G[6][7] = sqrt(w) W [H2 H3 H4] ; has_vicon ; idx0 ; idx1 ; w
constexpr int FPV_R = 0, FPV_A = 6, FPV_G = 42, FPV_HAS = 84, FPV_IDX0 = 85, FPV_IDX1 = 86, FPV_W = 87, FPV_STRIDE = 88;

// ---- normal-equation record of one state (row-major pieces, each piece contiguous so that the assembly can write
//  it as whole 3-row strips):  D, symmetric, as its packed lower triangle (row r at r (r + 1) / 2) at [0, 120);
//  O 15x15 at [120, 345) (coupling to the next state);  BG 15x10 at [345, 495): columns [0,9) = B (border to the 9
//  globals), 9 = g (right-hand side);  one pad double -> 496 (the record is a whole number of 16-byte units)
constexpr int NE_D = 0, NE_O = 120, NE_BG = 345, NE_BGW = 10, NE_STRIDE = 496;
// ---- elimination factor record of one state, in the register order of the mma.m8n8k4 fragments
//  (lane = 4 gid + tid), without the fragments that are zero by structure:
//  L^-1 = [T00 0; T10 T11], element (gid, 2 tid + e) of a tile on lane (gid, tid):
//    T00  lanes with 2 tid <= gid, double2 at slot tri_slot(gid, tid)                         (20 x 2 = 40)
//    T10  rows 8..14: lanes with gid < 7, double2 at slot lane                                (28 x 2 = 56)
//    T11  rows / columns 8..14: lanes with 2 tid <= gid < 7, double2 at slot tri_slot         (16 x 2 = 32)
//  W = L^-1 [O F B g] (15x40), lane holds W[8u+2tid+e][8t+gid] for column tile t:
//    u = 0: [t][lane] double2                                                                 (5 x 64 = 320)
//    u = 1: row 15 does not exist: [t][e = 0: lane (32) | e = 1: 3 gid + tid, tid < 3 (24)]   (5 x 56 = 280)
constexpr int FAC_LI00 = 0, FAC_LI10 = 40, FAC_LI11 = 96, FAC_W0 = 128, FAC_W1 = 448, FAC_STRIDE = 728;
// slot of lane (gid, tid) among the lanes with 2 tid <= gid, in lane order: 0 1 | 2 3 | 4 5 | 6 7 8 | ...
__host__ __device__ inline int tri_slot(int gid, int tid) { return ((gid + 1) >> 1) * ((gid >> 1) + 1) + tid; }
// ---- extended column order of one elimination step: [O 0..14 | F 15..29 | B 30..38 | g 39]
constexpr int XC_F = 15, XC_B = 30, XC_G = 39, XC_N = 40;
// ---- per-segment outputs, all "+S" (the consumer subtracts):
//  CY 15x40: S[r][c] = sum W_O^T [W_O W_F W_B w_g] after the last interior state (Schur terms of the right
//  separator: D, coupling to the left separator, border, rhs);  FB 25x25 (symmetric): sum over the segment
//  of [W_F W_B w_g]^T [W_F W_B w_g] (Schur terms of the left separator and the globals)
constexpr int SEG_CY = 0, SEG_FB = 600, SEG_STRIDE = 1232;
// ---- per-CTA partial sums of the assembly / cost kernels
//  assembly: Gamma vicon 7x7 (49) + g (7) + gravity 2x2 (4) + g (2) + lin_error0 (1) -> 63 (+pad)
constexpr int PART_STRIDE = 64;
// ---- chain-sharded mode: per-rank record of the second exchange of a damped solve
constexpr int CSUM_HALO = 8, CSUM_N = 24;

__host__ __device__ inline int tri(int r, int c) { return r * (r + 1) / 2 + c; } // packed lower, r >= c

// Per-trial description (device copy of the host staging result)
struct TrialDev {
  long long imu_off, n_imu;
  long long vic_off, n_vic;
  long long st_off, n_raw;
  int vic_cov_const; // 1: vic_Rconst applies to every pose
  int imu_unsorted;  // 1: the IMU store is not in time order: intervals are selected by the reference's scan from sample 0
  double vic_Rconst[18];
  v2gt_params prm;
};

// Device-resident Levenberg-Marquardt state of one trial
struct LmState {
  double lambda;
  double error;          // cost at the current estimate
  double current_error;  // cost at the start of the current outer iteration (defaultOptimize's currentError)
  double lin_error0;     // 1/2 |b|^2 of the current linearisation
  double new_error;      // scratch: cost at the trial point
  double lin_change;     // scratch
  double initial_error;
  int iterations, tries, linearizations;
  int lm_state;          // 0 running, 1 max iterations, 2 converged, 3 lambda gave up, 4 try cap
  int need_lin;          // 1: the next try starts a new outer iteration (relinearise)
  int solve_failed;      // set by the elimination kernels (non-positive pivot)
  int cur;               // which half of x/glob holds the current estimate
  int status;            // V2GT_OK or error code of the build stage
  int n_trace;
  int frozen;            // 1: the trial has used up its num_loop_relin budget; later build / optimise passes of the batch skip it
};
constexpr int TRACE_MAX = 1024; // rows of [lambda, cost_before, cost_after, lin_change, accepted] kept per trial

struct DevPtrs {
  int T;
  long long S, M, V; // totals
  int P;             // max segments per trial
  const TrialDev *trials;
  int *n_states;     // [T] live states per trial
  int *trial_of;     // [S] trial index of a state slot
  const double *imu_t, *imu_s, *vic_t, *vic_s, *vic_c;
  const long long *vic_c_off; // [T] offset into vic_c (or -1)
  double *st_t, *st_t_tmp;
  double *x, *glob;  // double buffered
  int *valid;        // [S] scratch flags of the build stage; afterwards the last vicon bracket of every state (a search hint)
  double *pre, *info, *cov;
  double *fpi, *fpv;
  double *ne;        // [S][NE_STRIDE]
  double *grad;      // [S][16] compact copy of the gradient g of every state (read by the trial-point evaluation)
  double *fac, *seg; // [S][FAC_STRIDE], [T][P][SEG_STRIDE]
  double *red_ne, *red_fac, *red_fb; // reduced (separator) chain: [T][P+1][NE_STRIDE], [T][P+1][FAC_STRIDE], [T][640]
  double *xsep;      // [T][P+1][16] separator steps (slot m = separator m, 1-based)
  double *delta;
  double *part_asm;  // [T][nchunk][PART_STRIDE]
  double *part_cost; // [T][nchunk][4]: cost, lin_error0, g.delta, |delta|^2
  double *gamma;     // [T][96]: Gamma 81 + g 9 (assembled, undamped)
  const double *gamma_sub; // [T][96] or null: Schur terms of a finer level, subtracted from gamma by k_reduced
  // ---- second level of the nested elimination: the separator chain of level 1 (red_ne) is itself cut into l2_P
  // segments and eliminated by the same kernels (launch_reduced builds the derived pointer set from these)
  int l2_P;                 // max level-2 segments per trial (0: level 2 not allocated)
  TrialDev *l2_trials;      // [T] st_off = slot of separator 1 of the trial in red_ne / xsep
  int *l2_nstates;          // [T] separators of level 1 = "states" of level 2
  double *l2_seg, *l2_red_ne, *l2_red_fac, *l2_xsep, *l2_gamma_sub;
  double *xg;        // [T][16]: global step
  LmState *lm;
  double *trace;     // [T][TRACE_MAX][5]
  int *n_active;     // device counter of running trials
  int *removed_no_vicon; // [T]
  int nchunk;        // chunks (CTAs) per trial of the per-state kernels
  // ---- chain-sharded mode (ONE long sequence split across ranks; T == 1).  This handle holds the states
  // [c_goff, c_goff + n_states[0]) of a chain of c_n states; it owns the global segments [c_j0, c_j1) and the
  // separators in front of them = local slots [c_lo, c_hi); the slots outside are halo copies of the neighbours'
  // boundary states.  The segment / separator buffers (seg, red_*, xsep, c_sepne) are indexed GLOBALLY.
  int chain;         // 0: ordinary batch
  int c_n, c_goff, c_j0, c_j1, c_lo, c_hi, c_rank, c_world;
  double *c_sepne;     // [P][NE_STRIDE] normal-equation records of every separator of the chain (all-gathered)
  double *c_sums;      // [8] cost, lin_error0, g.delta, |delta|^2, solve_failed of the whole chain (summed in rank order)
  double *c_gamma_loc; // [96] this rank's part of the global block (summed in rank order into gamma by k_chain_unpack)
  // second exchange of a damped solve: [5 local sums | pad | step of the last own state] per rank, all-gathered
  double *c_sums_send, *c_sums_all; // [CSUM_N], [c_world][CSUM_N]
  // first exchange of a damped solve: every rank contributes [its segments' outputs | its separators' records |
  // its part of the global block] (c_pack_send), and unpacks everybody's (c_pack_all, rank-major)
  double *c_pack_send, *c_pack_all;
};

// ---- tiled struct-of-arrays layout of the per-slot planes (pre, info, fpi, fpv): slots are grouped in tiles of
// SOA_W = 32; inside a tile the layout is [element][slot], so the one-thread-per-state kernels read 256 contiguous
// bytes per element and warp, and the E elements of a tile are ONE contiguous block of E x 256 bytes (an [element][S]
// layout spread a warp's reads over E streams 30 MB apart: poor DRAM row locality, a multiply per access)
constexpr int SOA_W = 32;
__host__ __device__ inline long long soa_base(long long s, int E) { return (s / SOA_W) * ((long long)E * SOA_W) + (s % SOA_W); }
__host__ __device__ inline long long soa_alloc(long long S, int E) { return ((S + SOA_W - 1) / SOA_W) * (long long)E * SOA_W; }

#if defined(__CUDACC__)
// views of one slot: element e lives at p[e * SOA_W]
struct SoaR {
  const double *p;
  __device__ __forceinline__ SoaR operator+(int e) const { return SoaR{p + e * SOA_W}; }
  __device__ __forceinline__ double operator[](int e) const { return __ldg(p + e * SOA_W); }
};
// the same view of a shared-memory tile ([element][column], column stride S): plain loads
struct SoaS {
  const double *p;
  int S;
  __device__ __forceinline__ SoaS operator+(int e) const { return SoaS{p + e * S, S}; }
  __device__ __forceinline__ double operator[](int e) const { return p[e * S]; }
};
struct SoaW {
  double *p;
  __device__ __forceinline__ double &operator[](int e) const { return p[e * SOA_W]; }
};
__device__ __forceinline__ SoaR soa_r(const double *plane, long long s, int E) { return SoaR{plane + soa_base(s, E)}; }
__device__ __forceinline__ SoaW soa_w(double *plane, long long s, int E) { return SoaW{plane + soa_base(s, E)}; }
// 32-byte read-only load as two 16-byte __ldg (there is no __ldg overload for double4)
__device__ __forceinline__ double4 ldg4(const double *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p));
  const double2 b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
}
#endif

#define V2GT_CHUNK 128 // states per CTA chunk in the per-state kernels (trial aligned)

} // namespace v2

#define V2_CUDA_CHECK(expr)                                                                                                      \
  do {                                                                                                                           \
    cudaError_t _e = (expr);                                                                                                     \
    if (_e != cudaSuccess) {                                                                                                     \
      fprintf(stderr, "[vicon2gt_b200] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(_e), __FILE__, __LINE__,                  \
              cudaGetErrorString(_e));                                                                                           \
      return V2GT_ERR_CUDA;                                                                                                      \
    }                                                                                                                            \
  } while (0)
