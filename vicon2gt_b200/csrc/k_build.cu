// vicon2gt_b200 kernels of the build stage (ViconGraphSolver::build_problem, .cpp:390-536):
//   k_init_states   initial guess + validity of every state (3 vicon lookups per state)
//   k_compact       erase invalid states (stream compaction per trial, order preserving)
// (kernel (1), the continuous preintegration, lives in k_preint.cu)
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
  if (i >= n || d.lm[t].status != V2GT_OK)
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

} // namespace v2
