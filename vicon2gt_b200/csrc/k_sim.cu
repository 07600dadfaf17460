// vicon2gt_b200 GPU simulator: the reference's src/sim (Simulator + BsplineSE3) for a whole Monte-Carlo batch at once
// (SURVEY 8(f) row 1).  The per-seed measurement streams are the ones a GCC build of the reference draws -- the same
// std::mt19937(seed) words, the same libstdc++ polar normal_distribution -- generated in parallel on the device:
//
//   host (once per batch, cheap)   control points of the spline (BsplineSE3.cpp:21-83), the three-clock schedule of
//                                  Simulator.cpp:135-262 (which sensor fires when: identical for every seed), the
//                                  random truth of every seed (Simulator.cpp:29-38: 9 draws) and the state-time grids
//                                  (run_simulation.cpp:149-158)
//   k_sim_spline_imu / _vic        thread per clock stamp: cumulative cubic B-spline pose / velocity / acceleration
//                                  (BsplineSE3.cpp:85-240, csrc/v2gt_spline.cuh -- the statements the host simulator runs);
//                                  evaluated ONCE per stamp, every trial of the batch shares the trajectory
//   k_sim_noise                    CTA per trial: MT19937 regenerated 624 words at a time (three dependent thirds, each
//                                  parallel), tempered, turned into polar-method attempts (4 words each, so attempt k
//                                  always sits at words 4k..4k+3 whether or not earlier attempts were rejected); the
//                                  accepted pairs are compacted in stream order.  Every distribution object of the
//                                  reference is created fresh for its 12 (IMU sample), 6 (vicon pose) or 9 (truth)
//                                  draws, so sample i owns accepted pairs [6i, 6i+6) and pose j pairs [3j, 3j+3) of
//                                  the same per-seed stream (both generators are seeded with `seed`, Simulator.cpp:85-88)
//   k_sim_bias                     CTA per trial: the bias random walks, summed in sample order (tiles staged in shared memory)
//   k_sim_imu / k_sim_vic          thread per (trial, sample): measurement = spline value (+ bias) + noise
//                                  (Simulator.cpp:170-185, 228-258), written in the batch solver's device layout
//   k_sim_state_in_vicon           thread per query: Simulator::get_state_in_vicon (Simulator.cpp:91-133) with a
//                                  bisection for the bias bracket instead of the reference's linear scan
// The arithmetic of the draws (uniform -> x, y, r2, accept) is exact in IEEE double and written with explicit
// round-to-nearest operations (no FMA contraction), so the accept / reject pattern and with it the stream positions are
// those of the host; the normals themselves agree to the last bits of log().
#include "../host/v2gt_sim.h"
#include "v2gt_common.cuh"
#include "v2gt_spline.cuh"
#include <algorithm>
#include <array>
#include <cmath>
#include <cstddef>
#include <cstdio>
#include <cstring>
#include <random>
#include <string>
#include <thread>
#include <vector>

namespace v2 {

constexpr int MT_N = 624;
constexpr int SIM_NOISE_THREADS = 256;
constexpr int SIM_ATTEMPTS = MT_N / 4; // polar attempts per regenerated block

struct SimTrialDev {
  double R_BtoI[9], p_BinI[3], viconimu_dt, R_GtoV[9];
  double sigma_vicon[6];
  double k_w, k_a, k_wb, k_ab; // sigma_w / sqrt(dt), sigma_a / sqrt(dt), sigma_wb sqrt(dt), sigma_ab sqrt(dt)
  unsigned seed;
  int pad;
};

__device__ __forceinline__ unsigned mt_twist(const unsigned a, const unsigned b, const unsigned c) {
  const unsigned y = (a & 0x80000000u) | (b & 0x7fffffffu);
  return c ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
}
__device__ __forceinline__ unsigned mt_temper(unsigned y) {
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  return y;
}
// std::generate_canonical<double, 53>(mt19937): two 32-bit words, low word first
__device__ __forceinline__ double canonical53(const unsigned w0, const unsigned w1) {
  const double sum = __dadd_rn((double)w0, __dmul_rn((double)w1, 4294967296.0));
  const double c = __dmul_rn(sum, 0x1p-64);
  return (c >= 1.0) ? 0x1.fffffffffffffp-1 : c;
}

// grid T, block 256.  normals: [T][2 n_pairs]; pair p of a trial = (first returned value, saved value).
__global__ void __launch_bounds__(SIM_NOISE_THREADS) k_sim_noise(const SimTrialDev *ts, long long n_pairs, double *normals) {
  __shared__ unsigned mt[2][MT_N];
  __shared__ int s_cnt[8];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    unsigned v = ts[t].seed;
    mt[0][0] = v;
    for (int i = 1; i < MT_N; i++) {
      v = 1812433253u * (v ^ (v >> 30)) + (unsigned)i;
      mt[0][i] = v;
    }
  }
  __syncthreads();
  double *out = normals + (long long)t * 2 * n_pairs;
  int cur = 0;
  long long base = 0;
  while (base < n_pairs) {
    const unsigned *o = mt[cur];
    unsigned *n = mt[cur ^ 1];
    // the in-place recurrence mt[i] = mt[i+397] ^ f(mt[i], mt[i+1]) in three thirds: each third only needs words of
    // the old block and of the thirds before it
    if (tid < 227)
      n[tid] = mt_twist(o[tid], o[tid + 1], o[tid + 397]);
    __syncthreads();
    if (tid < 227) {
      const int i = 227 + tid;
      n[i] = mt_twist(o[i], o[i + 1], n[i - 227]);
    }
    __syncthreads();
    if (tid < 170) {
      const int i = 454 + tid;
      n[i] = mt_twist(o[i], (i == MT_N - 1) ? n[0] : o[i + 1], n[i - 227]);
    }
    __syncthreads();
    bool acc = false;
    double x = 0, y = 0, r2 = 1;
    if (tid < SIM_ATTEMPTS) {
      const unsigned w0 = mt_temper(n[4 * tid]), w1 = mt_temper(n[4 * tid + 1]);
      const unsigned w2 = mt_temper(n[4 * tid + 2]), w3 = mt_temper(n[4 * tid + 3]);
      x = __dadd_rn(__dmul_rn(2.0, canonical53(w0, w1)), -1.0);
      y = __dadd_rn(__dmul_rn(2.0, canonical53(w2, w3)), -1.0);
      r2 = __dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y));
      acc = !(r2 > 1.0 || r2 == 0.0);
    }
    const unsigned bal = __ballot_sync(0xffffffffu, acc);
    if (lane == 0)
      s_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = 0, total = 0;
#pragma unroll
    for (int w = 0; w < (SIM_ATTEMPTS + 31) / 32; w++) {
      const int c = s_cnt[w];
      off += (w < warp) ? c : 0;
      total += c;
    }
    if (acc) {
      const long long p = base + off + __popc(bal & ((1u << lane) - 1u));
      if (p < n_pairs) {
        const double mult = __dsqrt_rn(__ddiv_rn(__dmul_rn(-2.0, log(r2)), r2));
        *reinterpret_cast<double2 *>(out + 2 * p) = make_double2(__dmul_rn(y, mult), __dmul_rn(x, mult));
      }
    }
    base += total;
    cur ^= 1;
  }
}

// IMU clock stamps: out[i][8] = w_IinI (3), R_GtoI (a_IinG + g e_z) (3), ok, 0
__global__ void __launch_bounds__(128) k_sim_spline_imu(const double *cp_t, const v2sp::M4 *cp, long long K, const double *clk, long long n,
                                                        double g, double *out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const double ts = clk[i];
  long long i1;
  double *o = out + 8 * i;
  if (!v2sp::spline_bracket(cp_t, K, ts, i1)) {
    for (int k = 0; k < 8; k++)
      o[k] = 0.0;
    return;
  }
  double R[9], p[3], w[3], v[3], al[3], a[3];
  v2sp::spline_eval_core(cp[i1 - 1], cp[i1], cp[i1 + 1], cp[i1 + 2], cp_t[i1], cp_t[i1 + 1], ts, 2, R, p, w, v, al, a);
  const double ag[3] = {a[0], a[1], a[2] + g};
  double acc[3];
  mv3(R, ag, acc);
  o[0] = w[0]; o[1] = w[1]; o[2] = w[2];
  o[3] = acc[0]; o[4] = acc[1]; o[5] = acc[2];
  o[6] = 1.0;
  o[7] = 0.0;
}

// vicon clock stamps (order 0) or truth queries (order 1): out[i][16] = R_GtoI (9), p_IinG (3), v_IinG (3), ok
__global__ void __launch_bounds__(128) k_sim_spline_pose(const double *cp_t, const v2sp::M4 *cp, long long K, const double *clk, long long n,
                                                         int order, double *out) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const double ts = clk[i];
  long long i1;
  double *o = out + 16 * i;
  for (int k = 0; k < 16; k++)
    o[k] = 0.0;
  if (!v2sp::spline_bracket(cp_t, K, ts, i1))
    return;
  double R[9], p[3], w[3], v[3] = {0, 0, 0};
  v2sp::spline_eval_core(cp[i1 - 1], cp[i1], cp[i1 + 1], cp[i1 + 2], cp_t[i1], cp_t[i1 + 1], ts, order, R, p, w, v, nullptr, nullptr);
  for (int k = 0; k < 9; k++)
    o[k] = R[k];
  for (int k = 0; k < 3; k++) {
    o[9 + k] = p[k];
    o[12 + k] = v[k];
  }
  o[15] = 1.0;
}

// CTA (two warps) per trial: hist[t][i + 2][c] = bias component c of [bg3 ba3] after sample i (Simulator.cpp:182-185), summed
// in sample order by six lanes; rows 0 and 1 are the two zero entries the constructor pushes (Simulator.cpp:72-77).  The
// increments of a tile of 128 samples are fetched coalesced one tile ahead (registers -> shared memory), the running sums
// leave through shared memory as one contiguous block: the chain of dependent additions is all that is left in sequence.
constexpr int BIAS_TILE = 128, BIAS_THREADS = 64, BIAS_PER_THREAD = BIAS_TILE * 6 / BIAS_THREADS;
__global__ void __launch_bounds__(BIAS_THREADS) k_sim_bias(long long M, const SimTrialDev *ts, const double *normals, long long np2,
                                                           double *hist) {
  __shared__ double inc[2][BIAS_TILE * 6];
  __shared__ double res[BIAS_TILE * 6];
  const int t = blockIdx.x, tid = threadIdx.x;
  const double *n = normals + (long long)t * np2;
  double *h = hist + (long long)t * (M + 2) * 6;
  if (tid < 12)
    h[tid] = 0.0;
  const double k_wb = ts[t].k_wb, k_ab = ts[t].k_ab;
  const long long ntile = (M + BIAS_TILE - 1) / BIAS_TILE;
  double r[BIAS_PER_THREAD];
  auto load = [&](const long long tile) {
#pragma unroll
    for (int k = 0; k < BIAS_PER_THREAD; k++) {
      const int idx = tid + BIAS_THREADS * k, smp = idx / 6, c = idx - 6 * smp;
      const long long i = tile * BIAS_TILE + smp;
      r[k] = (i < M) ? __ldg(n + 12 * i + 6 + c) : 0.0; // padding adds exact zeros
    }
  };
  auto stash = [&](const int buf) {
#pragma unroll
    for (int k = 0; k < BIAS_PER_THREAD; k++) {
      const int idx = tid + BIAS_THREADS * k;
      inc[buf][idx] = __dmul_rn((idx % 6 < 3) ? k_wb : k_ab, r[k]); // the step sigma sqrt(dt) n, rounded as on the host
    }
  };
  if (ntile > 0) {
    load(0);
    stash(0);
  }
  __syncthreads();
  double b = 0.0;
  for (long long tile = 0; tile < ntile; tile++) {
    const int cur = (int)(tile & 1);
    if (tile + 1 < ntile)
      load(tile + 1);
    if (tid < 6) {
#pragma unroll 8
      for (int u = 0; u < BIAS_TILE; u++) {
        b = __dadd_rn(b, inc[cur][6 * u + tid]);
        res[6 * u + tid] = b;
      }
    }
    if (tile + 1 < ntile)
      stash(cur ^ 1);
    __syncthreads();
    const long long i0 = tile * BIAS_TILE;
    const long long cnt = ((M - i0 < BIAS_TILE) ? (M - i0) : (long long)BIAS_TILE) * 6;
#pragma unroll
    for (int k = 0; k < BIAS_PER_THREAD; k++) {
      const int idx = tid + BIAS_THREADS * k;
      if (idx < cnt)
        h[(i0 + 2) * 6 + idx] = res[idx];
    }
    __syncthreads();
  }
}

// thread per (trial, IMU sample): rows (t, wm, am, 0) + the dense time array, the batch solver's device layout
__global__ void __launch_bounds__(256) k_sim_imu(int T, long long M, const double *clk, const double *base, const SimTrialDev *ts,
                                                 const double *normals, long long np2, const double *hist, double *imu_t, double *imu_s) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= T * M)
    return;
  const int t = (int)(idx / M);
  const long long i = idx - (long long)t * M;
  const double *n = normals + (long long)t * np2 + 12 * i;
  const double *b = hist + ((long long)t * (M + 2) + i + 1) * 6; // the bias BEFORE this sample's random-walk step
  const double *s = base + 8 * i;
  const double kw = ts[t].k_w, ka = ts[t].k_a;
  const double tm = clk[i];
  double *o = imu_s + 8 * idx;
  imu_t[idx] = tm;
  double4 r0, r1;
  r0.x = tm;
  r0.y = __dadd_rn(__dadd_rn(s[0], b[0]), __dmul_rn(kw, n[0]));
  r0.z = __dadd_rn(__dadd_rn(s[1], b[1]), __dmul_rn(kw, n[1]));
  r0.w = __dadd_rn(__dadd_rn(s[2], b[2]), __dmul_rn(kw, n[2]));
  r1.x = __dadd_rn(__dadd_rn(s[3], b[3]), __dmul_rn(ka, n[3]));
  r1.y = __dadd_rn(__dadd_rn(s[4], b[4]), __dmul_rn(ka, n[4]));
  r1.z = __dadd_rn(__dadd_rn(s[5], b[5]), __dmul_rn(ka, n[5]));
  r1.w = 0.0;
  *reinterpret_cast<double4 *>(o) = r0;
  *reinterpret_cast<double4 *>(o + 4) = r1;
}

// thread per (trial, vicon pose): rows (q_VtoB, p_BinV, 0) + stamps clock - viconimu_dt
__global__ void __launch_bounds__(128) k_sim_vic(int T, long long V, const double *clk, const double *base, const SimTrialDev *ts,
                                                 const double *normals, long long np2, double *vic_t, double *vic_s) {
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= T * V)
    return;
  const int t = (int)(idx / V);
  const long long j = idx - (long long)t * V;
  const SimTrialDev &q = ts[t];
  const double *n = normals + (long long)t * np2 + 6 * j;
  const double *R_GtoI = base + 16 * j, *p_IinG = base + 16 * j + 9;
  double R_GtoB[9], pb[3], p_BinG[3], R_VtoB[9], p_BinV[3];
  mm3_tn(q.R_BtoI, R_GtoI, R_GtoB);
  mv3_t(R_GtoI, q.p_BinI, pb);
  for (int d = 0; d < 3; d++)
    p_BinG[d] = p_IinG[d] + pb[d];
  mm3_nt(R_GtoB, q.R_GtoV, R_VtoB);
  mv3(q.R_GtoV, p_BinG, p_BinV);
  double w_vec[3], p_vec[3];
  for (int d = 0; d < 3; d++) {
    w_vec[d] = __dmul_rn(q.sigma_vicon[d], n[d]);
    p_vec[d] = __dmul_rn(q.sigma_vicon[3 + d], n[3 + d]);
  }
  double E[9], Rn[9], qq[4];
  exp_so3(w_vec, E);
  mm3(E, R_VtoB, Rn);
  rot_2_quat(Rn, qq);
  vic_t[idx] = __dadd_rn(clk[j], -q.viconimu_dt);
  double *o = vic_s + 8 * idx;
  *reinterpret_cast<double4 *>(o) = make_double4(qq[0], qq[1], qq[2], qq[3]);
  *reinterpret_cast<double4 *>(o + 4) = make_double4(__dadd_rn(p_BinV[0], p_vec[0]), __dadd_rn(p_BinV[1], p_vec[1]),
                                                     __dadd_rn(p_BinV[2], p_vec[2]), 0.0);
}

// thread per query: out17 = [t, q_VtoI(4), p_IinV(3), v_IinV(3), bg(3), ba(3)] (Simulator.cpp:91-133)
__global__ void __launch_bounds__(128) k_sim_state_in_vicon(long long n, const double *times, const double *pose /*[n][16]*/, const double *R_GtoV,
                                                            const double *bias_t, long long nb, const double *hist, double *out17, int *ok) {
  const long long k = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (k >= n)
    return;
  const double t = times[k];
  double *o = out17 + 17 * k;
  for (int d = 0; d < 17; d++)
    o[d] = 0.0;
  o[4] = 1.0;
  const double *ps = pose + 16 * k;
  const bool success_vel = ps[15] != 0.0;
  // first index with bias_t >= t (lower_bound); the bracket is (id, id + 1) = (lb - 1, lb)
  long long lo = 0, hi = nb;
  while (lo < hi) {
    const long long mid = (lo + hi) >> 1;
    if (bias_t[mid] < t)
      lo = mid + 1;
    else
      hi = mid;
  }
  const bool success_bias = (lo != 0 && lo != nb);
  ok[k] = (success_vel && success_bias) ? 1 : 0;
  if (!success_vel || !success_bias)
    return;
  const long long id = lo - 1;
  const double lambda = (t - bias_t[id]) / (bias_t[id + 1] - bias_t[id]);
  double Rq[9];
  mm3_nt(ps, R_GtoV, Rq);
  o[0] = t;
  rot_2_quat(Rq, o + 1);
  mv3(R_GtoV, ps + 9, o + 5);
  mv3(R_GtoV, ps + 12, o + 8);
  for (int d = 0; d < 6; d++)
    o[11 + d] = (1 - lambda) * hist[6 * id + d] + lambda * hist[6 * (id + 1) + d];
}

} // namespace v2

using namespace v2;

struct v2gt_simbatch {
  int device = 0, T = 0;
  std::vector<v2gt_sim_params> prm;
  std::vector<std::array<double, 8>> traj;
  v2sp::Spline spline;
  std::vector<double> cp_t;
  std::vector<v2sp::M4> cp;
  std::vector<double> imu_clk, vic_clk, bias_t;
  std::vector<SimTrialDev> ts;
  std::vector<std::vector<double>> state_t, vic_t;
  bool planned = false, ran = false;
  long long n_pairs = 0;
  long long alloc_key[4] = {0, 0, 0, 0}; // (T, M, V, K) the device buffers were sized for
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr; // around one generation, on `stream`
  double *d_cp_t = nullptr, *d_imu_clk = nullptr, *d_vic_clk = nullptr, *d_bias_t = nullptr;
  v2sp::M4 *d_cp = nullptr;
  double *d_base_imu = nullptr, *d_base_vic = nullptr, *d_normals = nullptr, *d_hist = nullptr;
  double *d_imu_t = nullptr, *d_imu_s = nullptr, *d_vic_t = nullptr, *d_vic_s = nullptr;
  SimTrialDev *d_ts = nullptr;
  std::vector<void *> allocs;
  double ms_generate = 0, ms_plan = 0;
  std::string error;
};

namespace {
template <typename Tp> int sim_alloc(v2gt_simbatch *s, Tp **p, size_t count) {
  void *q = nullptr;
  if (cudaMalloc(&q, std::max<size_t>(1, count) * sizeof(Tp)) != cudaSuccess) {
    fprintf(stderr, "[vicon2gt_b200] simulator: cudaMalloc of %zu bytes failed\n", count * sizeof(Tp));
    cudaGetLastError();
    return V2GT_ERR_CUDA;
  }
  s->allocs.push_back(q);
  *p = (Tp *)q;
  return V2GT_OK;
}
void sim_free(v2gt_simbatch *s) {
  for (void *p : s->allocs)
    cudaFree(p);
  s->allocs.clear();
  s->ran = false;
}
} // namespace

#define V2_TRY(expr)                                                                                                             \
  do {                                                                                                                           \
    int _s = (expr);                                                                                                             \
    if (_s != V2GT_OK)                                                                                                           \
      return _s;                                                                                                                 \
  } while (0)

namespace v2 {
// internal view for v2gt_batch_set_trial_sim (batch.cu): host times + device rows of one simulated trial
int simbatch_view(v2gt_simbatch *s, int trial, int *device, int64_t *n_imu, const double **imu_t, const double **d_imu_s, int64_t *n_vic,
                  const double **vic_t, const double **d_vic_s, int64_t *n_times, const double **state_t) {
  if (!s || trial < 0 || trial >= s->T || !s->ran)
    return V2GT_ERR_INVALID_ARG;
  const long long M = (long long)s->imu_clk.size(), V = (long long)s->vic_clk.size();
  *device = s->device;
  *n_imu = M;
  *imu_t = s->imu_clk.data();
  *d_imu_s = s->d_imu_s + 8 * M * trial;
  *n_vic = V;
  *vic_t = s->vic_t[trial].data();
  *d_vic_s = s->d_vic_s + 8 * V * trial;
  *n_times = (int64_t)s->state_t[trial].size();
  *state_t = s->state_t[trial].data();
  return V2GT_OK;
}
} // namespace v2

extern "C" {

int v2gt_simbatch_create(int32_t device, int32_t n_trials, int64_t n_rows, const double *rows8, v2gt_simbatch **out) {
  if (!out || device < 0 || n_trials <= 0 || n_rows < 4 || !rows8)
    return V2GT_ERR_INVALID_ARG;
  v2gt_simbatch *s = new v2gt_simbatch();
  s->device = device;
  s->T = n_trials;
  s->prm.resize((size_t)n_trials);
  for (int t = 0; t < n_trials; t++) {
    v2gt_sim_params &p = s->prm[(size_t)t];
    std::memset(&p, 0, sizeof(p));
    p.sigma_w = 1.6968e-04; // SimulatorParams.h:41-47
    p.sigma_wb = 1.9393e-05;
    p.sigma_a = 2.0000e-3;
    p.sigma_ab = 3.0000e-03;
    const double sv[6] = {1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2}; // SimulatorParams.h:53
    std::memcpy(p.sigma_vicon_pose, sv, sizeof(sv));
    p.gravity_magnitude = 9.81;
    p.sim_freq_imu = 400.0; // SimulatorParams.h:106-112
    p.sim_freq_cam = 10.0;
    p.sim_freq_vicon = 100.0;
    p.seed = t;
    p.state_freq = 100; // run_simulation.cpp:54
  }
  s->traj.resize((size_t)n_rows);
  for (int64_t i = 0; i < n_rows; i++)
    for (int d = 0; d < 8; d++)
      s->traj[(size_t)i][d] = rows8[8 * i + d];
  s->spline.feed_trajectory(s->traj);
  s->cp_t = s->spline.cp_t;
  s->cp = s->spline.cp;
  *out = s;
  return V2GT_OK;
}

int v2gt_simbatch_destroy(v2gt_simbatch *s) {
  if (!s)
    return V2GT_ERR_INVALID_ARG;
  if (!s->allocs.empty() || s->stream) {
    cudaSetDevice(s->device);
    sim_free(s);
    if (s->ev0)
      cudaEventDestroy(s->ev0);
    if (s->ev1)
      cudaEventDestroy(s->ev1);
    if (s->stream)
      cudaStreamDestroy(s->stream);
  }
  delete s;
  return V2GT_OK;
}

int v2gt_simbatch_set_trial(v2gt_simbatch *s, int32_t trial, const struct v2gt_sim_params *p) {
  if (!s || trial < 0 || trial >= s->T || !p)
    return V2GT_ERR_INVALID_ARG;
  s->prm[(size_t)trial] = *p;
  s->planned = false;
  s->ran = false;
  return V2GT_OK;
}

// Host part: everything that does not depend on the measurement noise.
int v2gt_simbatch_plan(v2gt_simbatch *s) {
  if (!s)
    return V2GT_ERR_INVALID_ARG;
  const v2gt_sim_params &P = s->prm[0];
  for (int t = 1; t < s->T; t++) {
    const v2gt_sim_params &Q = s->prm[(size_t)t];
    // one clock schedule and one spline per batch: the rates, the gravity magnitude and the sample cap are shared
    if (Q.sim_freq_imu != P.sim_freq_imu || Q.sim_freq_cam != P.sim_freq_cam || Q.sim_freq_vicon != P.sim_freq_vicon ||
        Q.gravity_magnitude != P.gravity_magnitude || Q.max_imu != P.max_imu || Q.state_freq != P.state_freq)
      return V2GT_ERR_INVALID_ARG;
  }
  // ---- clock schedule (Simulator.cpp:135-262; the spline lookups only decide when the loop stops)
  const v2sp::Spline &sp = s->spline;
  // a stamp is inside the spline where BsplineSE3::find_bounding_control_points succeeds (v2sp::spline_bracket states the
  // same condition on the sorted control-point times); every clock only moves forward, so each keeps a cursor = the
  // first control point later than its last stamp instead of searching
  const std::vector<double> &ct = s->cp_t;
  const long long K = (long long)ct.size();
  struct Cursor {
    long long ub = 0;
  } cur_imu, cur_vic, cur_start;
  auto in_spline_at = [&](Cursor &c, double ts) {
    while (c.ub < K && !(ct[(size_t)c.ub] > ts))
      c.ub++;
    return c.ub < K && c.ub >= 2 && c.ub + 1 < K;
  };
  auto in_spline = [&](double ts) { return in_spline_at(cur_start, ts); };
  s->imu_clk.clear();
  s->vic_clk.clear();
  s->bias_t.clear();
  if (!in_spline(sp.timestamp_start))
    return V2GT_ERR_INVALID_ARG;
  double last_imu = sp.timestamp_start, last_cam = sp.timestamp_start, last_vic = sp.timestamp_start;
  s->bias_t.push_back(last_imu - 1.0 / P.sim_freq_imu);
  s->bias_t.push_back(last_imu);
  const int64_t cap = P.max_imu > 0 ? P.max_imu : INT64_MAX;
  bool is_running = true;
  while (is_running) {
    if (!(last_cam + 1.0 / P.sim_freq_cam < last_imu + 1.0 / P.sim_freq_imu ||
          last_vic + 1.0 / P.sim_freq_vicon < last_imu + 1.0 / P.sim_freq_imu)) {
      last_imu += 1.0 / P.sim_freq_imu;
      if (!in_spline_at(cur_imu, last_imu) || (int64_t)s->imu_clk.size() >= cap) {
        is_running = false;
      } else {
        s->imu_clk.push_back(last_imu);
        s->bias_t.push_back(last_imu);
      }
    }
    if (!(last_imu + 1.0 / P.sim_freq_imu < last_cam + 1.0 / P.sim_freq_cam ||
          last_vic + 1.0 / P.sim_freq_vicon < last_cam + 1.0 / P.sim_freq_cam))
      last_cam += 1.0 / P.sim_freq_cam;
    if (!(last_imu + 1.0 / P.sim_freq_imu < last_vic + 1.0 / P.sim_freq_vicon ||
          last_cam + 1.0 / P.sim_freq_cam < last_vic + 1.0 / P.sim_freq_vicon)) {
      last_vic += 1.0 / P.sim_freq_vicon;
      if (!in_spline_at(cur_vic, last_vic))
        is_running = false;
      else
        s->vic_clk.push_back(last_vic);
    }
  }
  // ---- truth per seed (Simulator.cpp:29-38), noise scales, vicon stamps, state grids (run_simulation.cpp:149-158)
  const double dt = 1.0 / P.sim_freq_imu;
  s->ts.assign((size_t)s->T, SimTrialDev());
  s->state_t.resize((size_t)s->T); // capacities are kept from one plan to the next
  s->vic_t.resize((size_t)s->T);
  auto plan_trial = [&](int t) {
    const v2gt_sim_params &Q = s->prm[(size_t)t];
    SimTrialDev &d = s->ts[(size_t)t];
    std::memset(&d, 0, sizeof(d));
    {
      std::normal_distribution<double> w(0, 1);
      std::mt19937 r((unsigned)Q.seed);
      double wv[3];
      wv[0] = 0.1 * w(r);
      wv[1] = 0.1 * w(r);
      wv[2] = 0.1 * w(r);
      exp_so3(wv, d.R_BtoI);
      d.p_BinI[0] = 0.2 * w(r);
      d.p_BinI[1] = 0.2 * w(r);
      d.p_BinI[2] = 0.2 * w(r);
      d.viconimu_dt = 0.08 * w(r);
      const double ax = 0.1 * M_PI * w(r); // rot_x's angle is drawn first in a GCC build (see host/sim.cpp)
      const double ay = 0.1 * M_PI * w(r);
      rot_xy(ax, ay, d.R_GtoV);
    }
    if (Q.use_fixed_truth) {
      std::memcpy(d.R_BtoI, Q.fixed_R_BtoI, sizeof(d.R_BtoI));
      std::memcpy(d.p_BinI, Q.fixed_p_BinI, sizeof(d.p_BinI));
      std::memcpy(d.R_GtoV, Q.fixed_R_GtoV, sizeof(d.R_GtoV));
      d.viconimu_dt = Q.fixed_viconimu_dt;
    }
    std::memcpy(d.sigma_vicon, Q.sigma_vicon_pose, sizeof(d.sigma_vicon));
    d.k_w = Q.sigma_w / std::sqrt(dt);
    d.k_a = Q.sigma_a / std::sqrt(dt);
    d.k_wb = Q.sigma_wb * std::sqrt(dt);
    d.k_ab = Q.sigma_ab * std::sqrt(dt);
    d.seed = (unsigned)Q.seed;
    std::vector<double> &vt = s->vic_t[(size_t)t];
    vt.resize(s->vic_clk.size());
    for (size_t j = 0; j < vt.size(); j++)
      vt[j] = s->vic_clk[j] - d.viconimu_dt;
    std::vector<double> &st = s->state_t[(size_t)t];
    st.clear();
    if (Q.state_freq > 0 && !vt.empty() && vt.front() < vt.back()) {
      double temp_time = vt.front();
      const double end_time = vt.back();
      st.reserve((size_t)((end_time - temp_time) * (double)Q.state_freq) + 16);
      while (temp_time < end_time) {
        st.push_back(temp_time);
        temp_time += 1.0 / (double)Q.state_freq;
      }
    } else if (Q.state_freq <= 0) {
      for (const auto &row : s->traj)
        st.push_back(row[0]);
    }
  };
  {
    // independent per trial: a few host threads, static interleaved partition
    const int nthr = std::max(1, std::min({s->T, (int)std::thread::hardware_concurrency(), 16}));
    std::vector<std::thread> th;
    for (int k = 1; k < nthr; k++)
      th.emplace_back([&, k]() {
        for (int t = k; t < s->T; t += nthr)
          plan_trial(t);
      });
    for (int t = 0; t < s->T; t += nthr)
      plan_trial(t);
    for (auto &x : th)
      x.join();
  }
  s->planned = true;
  s->ran = false;
  return V2GT_OK;
}

int v2gt_simbatch_run(v2gt_simbatch *s) {
  if (!s)
    return V2GT_ERR_INVALID_ARG;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || s->device < 0 || s->device >= ndev) {
    cudaGetLastError();
    return V2GT_ERR_NO_DEVICE;
  }
  if (!s->planned)
    V2_TRY(v2gt_simbatch_plan(s));
  V2_CUDA_CHECK(cudaSetDevice(s->device));
  if (!s->stream)
    V2_CUDA_CHECK(cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking));
  if (!s->ev0)
    V2_CUDA_CHECK(cudaEventCreate(&s->ev0));
  if (!s->ev1)
    V2_CUDA_CHECK(cudaEventCreate(&s->ev1));
  const int T = s->T;
  const long long M = (long long)s->imu_clk.size(), V = (long long)s->vic_clk.size(), K = (long long)s->cp_t.size();
  s->n_pairs = std::max(std::max(6 * M, 3 * V), 5LL);
  const long long np2 = 2 * s->n_pairs;
  const bool reuse = !s->allocs.empty() && s->alloc_key[0] == T && s->alloc_key[1] == M && s->alloc_key[2] == V && s->alloc_key[3] == K;
  if (!reuse) {
    sim_free(s);
    s->alloc_key[0] = 0;
    V2_TRY(sim_alloc(s, &s->d_cp_t, (size_t)K));
    V2_TRY(sim_alloc(s, &s->d_cp, (size_t)K));
    V2_TRY(sim_alloc(s, &s->d_imu_clk, (size_t)M));
    V2_TRY(sim_alloc(s, &s->d_vic_clk, (size_t)V));
    V2_TRY(sim_alloc(s, &s->d_bias_t, (size_t)M + 2));
    V2_TRY(sim_alloc(s, &s->d_base_imu, (size_t)M * 8));
    V2_TRY(sim_alloc(s, &s->d_base_vic, (size_t)V * 16));
    V2_TRY(sim_alloc(s, &s->d_normals, (size_t)T * np2));
    V2_TRY(sim_alloc(s, &s->d_hist, (size_t)T * (M + 2) * 6));
    V2_TRY(sim_alloc(s, &s->d_imu_t, (size_t)T * M));
    V2_TRY(sim_alloc(s, &s->d_imu_s, (size_t)T * M * 8));
    V2_TRY(sim_alloc(s, &s->d_vic_t, (size_t)T * V));
    V2_TRY(sim_alloc(s, &s->d_vic_s, (size_t)T * V * 8));
    V2_TRY(sim_alloc(s, &s->d_ts, (size_t)T));
    s->alloc_key[0] = T;
    s->alloc_key[1] = M;
    s->alloc_key[2] = V;
    s->alloc_key[3] = K;
  }
  s->ran = false;
  cudaStream_t st = s->stream;
  const cudaEvent_t e0 = s->ev0, e1 = s->ev1;
  V2_CUDA_CHECK(cudaEventRecord(e0, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_cp_t, s->cp_t.data(), sizeof(double) * K, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_cp, s->cp.data(), sizeof(v2sp::M4) * K, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_imu_clk, s->imu_clk.data(), sizeof(double) * M, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_vic_clk, s->vic_clk.data(), sizeof(double) * V, cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_bias_t, s->bias_t.data(), sizeof(double) * (M + 2), cudaMemcpyHostToDevice, st));
  V2_CUDA_CHECK(cudaMemcpyAsync(s->d_ts, s->ts.data(), sizeof(SimTrialDev) * T, cudaMemcpyHostToDevice, st));
  k_sim_noise<<<T, SIM_NOISE_THREADS, 0, st>>>(s->d_ts, s->n_pairs, s->d_normals);
  if (M > 0)
    k_sim_spline_imu<<<(unsigned)((M + 127) / 128), 128, 0, st>>>(s->d_cp_t, s->d_cp, K, s->d_imu_clk, M, s->prm[0].gravity_magnitude, s->d_base_imu);
  if (V > 0)
    k_sim_spline_pose<<<(unsigned)((V + 127) / 128), 128, 0, st>>>(s->d_cp_t, s->d_cp, K, s->d_vic_clk, V, 0, s->d_base_vic);
  k_sim_bias<<<T, BIAS_THREADS, 0, st>>>(M, s->d_ts, s->d_normals, np2, s->d_hist);
  if (M > 0)
    k_sim_imu<<<(unsigned)(((long long)T * M + 255) / 256), 256, 0, st>>>(T, M, s->d_imu_clk, s->d_base_imu, s->d_ts, s->d_normals, np2,
                                                                           s->d_hist, s->d_imu_t, s->d_imu_s);
  if (V > 0)
    k_sim_vic<<<(unsigned)(((long long)T * V + 127) / 128), 128, 0, st>>>(T, V, s->d_vic_clk, s->d_base_vic, s->d_ts, s->d_normals, np2,
                                                                          s->d_vic_t, s->d_vic_s);
  V2_CUDA_CHECK(cudaGetLastError());
  V2_CUDA_CHECK(cudaEventRecord(e1, st));
  V2_CUDA_CHECK(cudaStreamSynchronize(st));
  float ms = 0;
  cudaEventElapsedTime(&ms, e0, e1);
  s->ms_generate = ms;
  s->ran = true;
  return V2GT_OK;
}

int v2gt_simbatch_sizes(v2gt_simbatch *s, int32_t trial, int64_t *n_imu, int64_t *n_vicon, int64_t *n_states) {
  if (!s || trial < 0 || trial >= s->T || !s->planned)
    return V2GT_ERR_INVALID_ARG;
  if (n_imu)
    *n_imu = (int64_t)s->imu_clk.size();
  if (n_vicon)
    *n_vicon = (int64_t)s->vic_clk.size();
  if (n_states)
    *n_states = (int64_t)s->state_t[(size_t)trial].size();
  return V2GT_OK;
}

// Caller layout of the host simulator (v2gt_sim_get).  The stamps come from the host plan; the measurements need run().
int v2gt_simbatch_get(v2gt_simbatch *s, int32_t trial, double *imu_t, double *imu_w, double *imu_a, double *vic_t, double *vic_q,
                      double *vic_p, double *state_t) {
  if (!s || trial < 0 || trial >= s->T || !s->planned)
    return V2GT_ERR_INVALID_ARG;
  const size_t M = s->imu_clk.size(), V = s->vic_clk.size();
  if (imu_t && M)
    std::memcpy(imu_t, s->imu_clk.data(), sizeof(double) * M);
  if (vic_t && V)
    std::memcpy(vic_t, s->vic_t[(size_t)trial].data(), sizeof(double) * V);
  if (state_t && !s->state_t[(size_t)trial].empty())
    std::memcpy(state_t, s->state_t[(size_t)trial].data(), sizeof(double) * s->state_t[(size_t)trial].size());
  if (!(imu_w || imu_a || vic_q || vic_p))
    return V2GT_OK;
  if (!s->ran)
    return V2GT_ERR_NOT_BUILT;
  V2_CUDA_CHECK(cudaSetDevice(s->device));
  std::vector<double> rows(8 * std::max(M, V));
  if ((imu_w || imu_a) && M) {
    V2_CUDA_CHECK(cudaMemcpy(rows.data(), s->d_imu_s + 8 * M * (size_t)trial, sizeof(double) * 8 * M, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < M; i++)
      for (int d = 0; d < 3; d++) {
        if (imu_w)
          imu_w[3 * i + d] = rows[8 * i + 1 + d];
        if (imu_a)
          imu_a[3 * i + d] = rows[8 * i + 4 + d];
      }
  }
  if ((vic_q || vic_p) && V) {
    V2_CUDA_CHECK(cudaMemcpy(rows.data(), s->d_vic_s + 8 * V * (size_t)trial, sizeof(double) * 8 * V, cudaMemcpyDeviceToHost));
    for (size_t j = 0; j < V; j++) {
      if (vic_q)
        for (int d = 0; d < 4; d++)
          vic_q[4 * j + d] = rows[8 * j + d];
      if (vic_p)
        for (int d = 0; d < 3; d++)
          vic_p[3 * j + d] = rows[8 * j + 4 + d];
    }
  }
  return V2GT_OK;
}

int v2gt_simbatch_get_truth(v2gt_simbatch *s, int32_t trial, double *R_BtoI, double *p_BinI, double *viconimu_dt, double *R_GtoV) {
  if (!s || trial < 0 || trial >= s->T || !s->planned)
    return V2GT_ERR_INVALID_ARG;
  const SimTrialDev &d = s->ts[(size_t)trial];
  if (R_BtoI)
    std::memcpy(R_BtoI, d.R_BtoI, sizeof(d.R_BtoI));
  if (p_BinI)
    std::memcpy(p_BinI, d.p_BinI, sizeof(d.p_BinI));
  if (viconimu_dt)
    *viconimu_dt = d.viconimu_dt;
  if (R_GtoV)
    std::memcpy(R_GtoV, d.R_GtoV, sizeof(d.R_GtoV));
  return V2GT_OK;
}

int v2gt_simbatch_get_state_in_vicon(v2gt_simbatch *s, int32_t trial, int64_t n, const double *times, double *out17, int32_t *ok) {
  if (!s || trial < 0 || trial >= s->T || n < 0 || (n && (!times || !out17)))
    return V2GT_ERR_INVALID_ARG;
  if (!s->ran)
    return V2GT_ERR_NOT_BUILT;
  if (n == 0)
    return V2GT_OK;
  V2_CUDA_CHECK(cudaSetDevice(s->device));
  const long long M = (long long)s->imu_clk.size(), K = (long long)s->cp_t.size();
  double *d_t = nullptr, *d_pose = nullptr, *d_out = nullptr;
  int *d_ok = nullptr;
  int rc = V2GT_ERR_CUDA;
  cudaStream_t st = s->stream;
  std::vector<int> h_ok((size_t)n);
  if (cudaMalloc(&d_t, sizeof(double) * n) == cudaSuccess && cudaMalloc(&d_pose, sizeof(double) * 16 * n) == cudaSuccess &&
      cudaMalloc(&d_out, sizeof(double) * 17 * n) == cudaSuccess && cudaMalloc(&d_ok, sizeof(int) * n) == cudaSuccess) {
    cudaMemcpyAsync(d_t, times, sizeof(double) * n, cudaMemcpyHostToDevice, st);
    k_sim_spline_pose<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(s->d_cp_t, s->d_cp, K, d_t, n, 1, d_pose);
    // R_GtoV of the trial lives at a fixed offset of its SimTrialDev
    const double *d_RGtoV = reinterpret_cast<const double *>(s->d_ts + trial) + offsetof(SimTrialDev, R_GtoV) / sizeof(double);
    k_sim_state_in_vicon<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, d_t, d_pose, d_RGtoV, s->d_bias_t, M + 2,
                                                                      s->d_hist + (long long)trial * (M + 2) * 6, d_out, d_ok);
    cudaMemcpyAsync(out17, d_out, sizeof(double) * 17 * n, cudaMemcpyDeviceToHost, st);
    cudaMemcpyAsync(h_ok.data(), d_ok, sizeof(int) * n, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) == cudaSuccess && cudaGetLastError() == cudaSuccess)
      rc = V2GT_OK;
  }
  cudaFree(d_t);
  cudaFree(d_pose);
  cudaFree(d_out);
  cudaFree(d_ok);
  if (rc == V2GT_OK && ok)
    for (int64_t k = 0; k < n; k++)
      ok[k] = h_ok[(size_t)k];
  return rc;
}

int v2gt_simbatch_get_timing(v2gt_simbatch *s, double *ms_generate) {
  if (!s || !ms_generate)
    return V2GT_ERR_INVALID_ARG;
  *ms_generate = s->ms_generate;
  return V2GT_OK;
}

} // extern "C"
