// vicon2gt_b200: evaluation-only entry points of the C ABI (parity tests).  They drive the
// production kernels on a built batch and copy intermediate results back to the host.
#include "v2gt_interp.cuh"
#include <cstdio>
#include <cstring>
#include <vector>

namespace v2 {
int launch_linearize(const DevPtrs &d, int n_max, int force_all, cudaStream_t st);
int launch_cost(const DevPtrs &d, int n_max, int which, int force_all, cudaStream_t st);
int launch_solve(const DevPtrs &d, int force_all, cudaStream_t st);

__global__ void k_eval_interp(DevPtrs d, int t, long long n, const double *__restrict__ tq, double thresh, int *idx0, int *idx1,
                              int *ok, double *q, double *p, double *R, double *Ht) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const TrialDev &tr = d.trials[t];
  const bool cc = tr.vic_cov_const != 0;
  const double *vc = cc ? tr.vic_Rconst : d.vic_c + d.vic_c_off[t];
  InterpResult r;
  vicon_interp(d.vic_t + tr.vic_off, d.vic_s + 8 * tr.vic_off, vc, cc, tr.n_vic, tq[i], thresh, true, r);
  idx0[i] = r.idx0;
  idx1[i] = r.idx1;
  ok[i] = r.ok;
  for (int k = 0; k < 4; k++)
    q[4 * i + k] = r.q[k];
  for (int k = 0; k < 3; k++)
    p[3 * i + k] = r.p[k];
  double R36[36];
  interp_dense_cov(r, R36);
  for (int k = 0; k < 36; k++)
    R[36 * i + k] = R36[k];
  for (int k = 0; k < 6; k++)
    Ht[6 * i + k] = r.Ht[k];
}
} // namespace v2

using namespace v2;

// mirror of the opaque handle's leading members is NOT used; the accessors below are provided by batch.cu
struct v2gt_batch;
extern "C" {
int v2gt_batch_sync(v2gt_batch *b);
}
namespace v2 {
// accessors implemented in batch_access.cu-style section of batch.cu
const DevPtrs *batch_devptrs(v2gt_batch *b);
cudaStream_t batch_stream(v2gt_batch *b);
int batch_device(v2gt_batch *b);
int batch_nmax(v2gt_batch *b);
bool batch_built(v2gt_batch *b);
bool batch_uploaded(v2gt_batch *b);
void batch_invalidate(v2gt_batch *b);
} // namespace v2

namespace {
struct TmpDev {
  std::vector<void *> p;
  ~TmpDev() {
    for (void *q : p)
      cudaFree(q);
  }
  template <typename Tp> Tp *get(size_t n) {
    void *q = nullptr;
    if (cudaMalloc(&q, (n ? n : 1) * sizeof(Tp)) != cudaSuccess)
      return nullptr;
    p.push_back(q);
    return (Tp *)q;
  }
};

int trial_geom(v2gt_batch *b, int trial, TrialDev &td, int &n, LmState &lm) {
  const DevPtrs &d = *batch_devptrs(b);
  V2_CUDA_CHECK(cudaStreamSynchronize(batch_stream(b)));
  V2_CUDA_CHECK(cudaMemcpy(&td, d.trials + trial, sizeof(TrialDev), cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(&n, d.n_states + trial, sizeof(int), cudaMemcpyDeviceToHost));
  V2_CUDA_CHECK(cudaMemcpy(&lm, d.lm + trial, sizeof(LmState), cudaMemcpyDeviceToHost));
  return V2GT_OK;
}
} // namespace

extern "C" {

int v2gt_eval_interp(v2gt_batch *b, int32_t trial, int64_t n, const double *t_query, double gap_thresh, int32_t *idx0,
                     int32_t *idx1, int32_t *ok, double *q, double *p, double *R, double *H_toff) {
  if (!b || !batch_uploaded(b) || n < 0 || (n && !t_query))
    return V2GT_ERR_INVALID_ARG;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T)
    return V2GT_ERR_INVALID_ARG;
  if (n == 0)
    return V2GT_OK;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  TmpDev tmp;
  double *d_t = tmp.get<double>(n), *d_q = tmp.get<double>(4 * n), *d_p = tmp.get<double>(3 * n), *d_R = tmp.get<double>(36 * n),
         *d_H = tmp.get<double>(6 * n);
  int *d_i0 = tmp.get<int>(n), *d_i1 = tmp.get<int>(n), *d_ok = tmp.get<int>(n);
  if (!d_t || !d_q || !d_p || !d_R || !d_H || !d_i0 || !d_i1 || !d_ok)
    return V2GT_ERR_CUDA;
  cudaStream_t st = batch_stream(b);
  V2_CUDA_CHECK(cudaMemcpyAsync(d_t, t_query, sizeof(double) * n, cudaMemcpyHostToDevice, st));
  k_eval_interp<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d, trial, n, d_t, gap_thresh, d_i0, d_i1, d_ok, d_q, d_p, d_R, d_H);
  V2_CUDA_CHECK(cudaGetLastError());
  V2_CUDA_CHECK(cudaStreamSynchronize(st));
  if (idx0)
    V2_CUDA_CHECK(cudaMemcpy(idx0, d_i0, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (idx1)
    V2_CUDA_CHECK(cudaMemcpy(idx1, d_i1, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (ok)
    V2_CUDA_CHECK(cudaMemcpy(ok, d_ok, sizeof(int) * n, cudaMemcpyDeviceToHost));
  if (q)
    V2_CUDA_CHECK(cudaMemcpy(q, d_q, sizeof(double) * 4 * n, cudaMemcpyDeviceToHost));
  if (p)
    V2_CUDA_CHECK(cudaMemcpy(p, d_p, sizeof(double) * 3 * n, cudaMemcpyDeviceToHost));
  if (R)
    V2_CUDA_CHECK(cudaMemcpy(R, d_R, sizeof(double) * 36 * n, cudaMemcpyDeviceToHost));
  if (H_toff)
    V2_CUDA_CHECK(cudaMemcpy(H_toff, d_H, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost));
  return V2GT_OK;
}

int v2gt_batch_get_preint(v2gt_batch *b, int32_t trial, int64_t capacity, double *out, int64_t *n_out) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  TrialDev td;
  int n;
  LmState lm;
  int s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  const int m = n > 0 ? n - 1 : 0;
  if (n_out)
    *n_out = m;
  if (!out)
    return V2GT_OK;
  if (capacity < m)
    return V2GT_ERR_INVALID_ARG;
  // device layout is the tiled struct-of-arrays (tiles of SOA_W slots, [element][slot] inside): copy the tiles that
  // hold the trial's slots, then gather
  if (m) {
    const long long t0 = td.st_off / SOA_W, t1 = (td.st_off + m - 1) / SOA_W + 1; // tile range
    std::vector<double> pre((size_t)(t1 - t0) * PRE_STRIDE * SOA_W), inf((size_t)(t1 - t0) * INFO_STRIDE * SOA_W);
    V2_CUDA_CHECK(cudaMemcpy(pre.data(), d.pre + t0 * PRE_STRIDE * SOA_W, sizeof(double) * pre.size(), cudaMemcpyDeviceToHost));
    V2_CUDA_CHECK(cudaMemcpy(inf.data(), d.info + t0 * INFO_STRIDE * SOA_W, sizeof(double) * inf.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < m; i++) {
      double *o = out + (size_t)i * V2GT_PREINT_DIM;
      const long long s = td.st_off + i;
      const double *pp = pre.data() + (soa_base(s, PRE_STRIDE) - t0 * PRE_STRIDE * SOA_W);
      const double *ip = inf.data() + (soa_base(s, INFO_STRIDE) - t0 * INFO_STRIDE * SOA_W);
      for (int e = 0; e < 63; e++)
        o[e] = pp[(size_t)e * SOA_W]; // identical order for [0, 63)
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < 15; c++)
          o[63 + 15 * r + c] = ip[(size_t)((r >= c) ? tri(r, c) : tri(c, r)) * SOA_W];
    }
  }
  return V2GT_OK;
}

int v2gt_batch_get_preint_cov(v2gt_batch *b, int32_t trial, int64_t capacity, double *out, int64_t *n_out) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T || !d.cov)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  TrialDev td;
  int n;
  LmState lm;
  int s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  const int m = n > 0 ? n - 1 : 0;
  if (n_out)
    *n_out = m;
  if (!out)
    return V2GT_OK;
  if (capacity < m)
    return V2GT_ERR_INVALID_ARG;
  if (m)
    V2_CUDA_CHECK(cudaMemcpy(out, d.cov + td.st_off * 225, sizeof(double) * 225 * m, cudaMemcpyDeviceToHost));
  return V2GT_OK;
}

int v2gt_batch_set_estimate(v2gt_batch *b, int32_t trial, int64_t n_in, const double *states, const double *globals) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  TrialDev td;
  int n;
  LmState lm;
  int s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  if (states) {
    if (n_in != n)
      return V2GT_ERR_INVALID_ARG;
    if (n)
      V2_CUDA_CHECK(cudaMemcpy(d.x + ((long long)lm.cur * d.S + td.st_off) * X_STRIDE, states, sizeof(double) * n * X_STRIDE,
                               cudaMemcpyHostToDevice));
  }
  if (globals) {
    double g[GLOB_STRIDE] = {0};
    std::memcpy(g, globals, sizeof(double) * 10);
    V2_CUDA_CHECK(cudaMemcpy(d.glob + ((long long)lm.cur * d.T + trial) * GLOB_STRIDE, g, sizeof(g), cudaMemcpyHostToDevice));
  }
  batch_invalidate(b);
  return V2GT_OK;
}

static int host_cost(v2gt_batch *b, int trial, int n, double *cost) {
  const DevPtrs &d = *batch_devptrs(b);
  const int used = (n + V2GT_CHUNK - 1) / V2GT_CHUNK;
  std::vector<double> pc((size_t)used * 4);
  if (used)
    V2_CUDA_CHECK(cudaMemcpy(pc.data(), d.part_cost + (long long)trial * d.nchunk * 4, sizeof(double) * 4 * used,
                             cudaMemcpyDeviceToHost));
  double c = 0;
  for (int k = 0; k < used; k++)
    c += pc[4 * k];
  *cost = c;
  return V2GT_OK;
}

int v2gt_eval_cost(v2gt_batch *b, int32_t trial, double *cost) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T || !cost)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  int s = launch_cost(d, batch_nmax(b), 0, 1, batch_stream(b));
  if (s != V2GT_OK)
    return s;
  TrialDev td;
  int n;
  LmState lm;
  s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  return host_cost(b, trial, n, cost);
}

int v2gt_eval_linearize(v2gt_batch *b, int32_t trial, int64_t capacity, double *D, double *O, double *B, double *g, double *Gamma,
                        double *g_gamma, double *cost, double *lin_error0) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  cudaStream_t st = batch_stream(b);
  TrialDev td;
  int n;
  LmState lm;
  int s;
  if (cost) {
    s = launch_cost(d, batch_nmax(b), 0, 1, st);
    if (s != V2GT_OK)
      return s;
    s = trial_geom(b, trial, td, n, lm);
    if (s != V2GT_OK)
      return s;
    s = host_cost(b, trial, n, cost);
    if (s != V2GT_OK)
      return s;
  }
  s = launch_linearize(d, batch_nmax(b), 1, st);
  if (s != V2GT_OK)
    return s;
  s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  if ((D || O || B || g) && capacity < n)
    return V2GT_ERR_INVALID_ARG;
  if (n && (D || O || B || g)) {
    // unpack the per-state records [D packed lower triangle | O 15x15 | BG 15x10 = B 9, g 1]
    std::vector<double> rec((size_t)n * NE_STRIDE);
    V2_CUDA_CHECK(cudaMemcpy(rec.data(), d.ne + td.st_off * NE_STRIDE, sizeof(double) * rec.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) {
      const double *r = &rec[(size_t)i * NE_STRIDE];
      if (D)
        for (int a = 0; a < 15; a++)
          for (int c = 0; c <= a; c++)
            D[225 * (size_t)i + 15 * a + c] = D[225 * (size_t)i + 15 * c + a] = r[NE_D + tri(a, c)];
      for (int a = 0; a < 15; a++) {
        if (O)
          std::memcpy(O + 225 * (size_t)i + 15 * a, r + NE_O + 15 * a, sizeof(double) * 15);
        if (B)
          std::memcpy(B + 135 * (size_t)i + 9 * a, r + NE_BG + NE_BGW * a, sizeof(double) * 9);
        if (g)
          g[15 * (size_t)i + a] = r[NE_BG + NE_BGW * a + 9];
      }
    }
  }
  double gm[96];
  V2_CUDA_CHECK(cudaMemcpy(gm, d.gamma + (long long)trial * 96, sizeof(gm), cudaMemcpyDeviceToHost));
  if (Gamma)
    std::memcpy(Gamma, gm, sizeof(double) * 81);
  if (g_gamma)
    std::memcpy(g_gamma, gm + 81, sizeof(double) * 9);
  if (lin_error0)
    *lin_error0 = lm.lin_error0;
  batch_invalidate(b);
  return V2GT_OK;
}

int v2gt_eval_solve(v2gt_batch *b, int32_t trial, double lambda, int64_t capacity, double *delta, int32_t *ok) {
  if (!b || !batch_built(b))
    return V2GT_ERR_NOT_BUILT;
  const DevPtrs &d = *batch_devptrs(b);
  if (trial < 0 || trial >= d.T)
    return V2GT_ERR_INVALID_ARG;
  V2_CUDA_CHECK(cudaSetDevice(batch_device(b)));
  cudaStream_t st = batch_stream(b);
  TrialDev td;
  int n;
  LmState lm;
  int s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  if (delta && capacity < 15LL * n + 9)
    return V2GT_ERR_INVALID_ARG;
  lm.lambda = lambda;
  lm.solve_failed = 0;
  V2_CUDA_CHECK(cudaMemcpy(d.lm + trial, &lm, sizeof(LmState), cudaMemcpyHostToDevice));
  s = launch_solve(d, 1, st);
  if (s != V2GT_OK)
    return s;
  s = trial_geom(b, trial, td, n, lm);
  if (s != V2GT_OK)
    return s;
  if (ok)
    *ok = lm.solve_failed ? 0 : 1;
  if (delta && !lm.solve_failed) {
    std::vector<double> tmp((size_t)n * 16);
    if (n)
      V2_CUDA_CHECK(cudaMemcpy(tmp.data(), d.delta + td.st_off * 16, sizeof(double) * tmp.size(), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++)
      std::memcpy(delta + 15 * (size_t)i, &tmp[(size_t)i * 16], sizeof(double) * 15);
    double xg[16];
    V2_CUDA_CHECK(cudaMemcpy(xg, d.xg + (long long)trial * 16, sizeof(xg), cudaMemcpyDeviceToHost));
    std::memcpy(delta + 15 * (size_t)n, xg, sizeof(double) * 9);
  }
  batch_invalidate(b);
  return V2GT_OK;
}

} // extern "C"

// ---- FP64 FMA peak microbenchmark (the roofline denominator MEASURED_PEAKS.json does not carry)
namespace v2 {
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-9, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int u = 0; u < 8; u++) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}
} // namespace v2

extern "C" int v2gt_measure_fp64_peak(int32_t device, double *tflops) {
  if (!tflops)
    return V2GT_ERR_INVALID_ARG;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return V2GT_ERR_NO_DEVICE;
  }
  V2_CUDA_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  V2_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
  const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 4096;
  double *d_out = nullptr;
  V2_CUDA_CHECK(cudaMalloc(&d_out, sizeof(double) * blocks * threads));
  cudaEvent_t e0, e1;
  V2_CUDA_CHECK(cudaEventCreate(&e0));
  V2_CUDA_CHECK(cudaEventCreate(&e1));
  double best = 0;
  for (int rep = 0; rep < 5; rep++) {
    V2_CUDA_CHECK(cudaEventRecord(e0));
    v2::k_dfma_peak<<<blocks, threads>>>(d_out, iters, 0.999999, 1e-7);
    V2_CUDA_CHECK(cudaEventRecord(e1));
    V2_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
    const double tf = flops / (ms * 1e-3) * 1e-12;
    if (rep > 0 && tf > best)
      best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d_out);
  *tflops = best;
  return V2GT_OK;
}
