// vicon2gt_b200: batched trajectory-error statistics of a Monte-Carlo batch ("next" row 2).
//
// Replaces, for every trial of a batch at once, the error report of the reference's simulation driver:
//   per state   ori = 2 |vec(q_gt (x) q_est^-1)| 180/pi,  pos = |p_est - p_gt|,  vel = |v_est - v_gt|
//               (src/run_simulation.cpp:216-218; no alignment, both in the vicon frame)
//   per trial   Stats::calculate (src/utils/stats.h:64-112): values sorted ascending, min, max, median, mean and rmse
//               summed in sorted order, std with the n-1 denominator, "99th percentile" = mean + 2.326 std
// The truth states come from the host simulator (Simulator::get_state_in_vicon) at the surviving state times.
//
//   k_err_values   one thread per state slot: the three error values (slots that hold no state get +inf)
//   cub::DeviceSegmentedSort   one segment per (metric, trial)
//   k_err_stats    one warp per (metric, trial): fixed-order sums over the sorted values
#include "v2gt_common.cuh"
#include "v2gt_math.cuh"
#include <cub/device/device_segmented_sort.cuh>

namespace v2 {

// truth: [S][10] = q_VtoI (JPL, 4), p_IinV (3), v_IinV (3) per state slot; err: [3][S]
__global__ void __launch_bounds__(128) k_err_values(DevPtrs d, const double *__restrict__ truth, double *__restrict__ err) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= d.S)
    return;
  const int t = d.trial_of[s];
  const TrialDev &tr = d.trials[t];
  const long long i = s - tr.st_off;
  double e_ori = INFINITY, e_pos = INFINITY, e_vel = INFINITY;
  if (i >= 0 && i < d.n_states[t] && d.lm[t].status == V2GT_OK) {
    const double *x = d.x + ((long long)d.lm[t].cur * d.S + s) * X_STRIDE;
    const double *g = truth + s * 10;
    double qi[4], qe[4];
    const double q_est[4] = {x[0], x[1], x[2], x[3]};
    quat_inv(q_est, qi);
    quat_mul(g, qi, qe);
    e_ori = 180.0 / M_PI * (2.0 * sqrt(qe[0] * qe[0] + qe[1] * qe[1] + qe[2] * qe[2]));
    const double dp[3] = {x[13] - g[4], x[14] - g[5], x[15] - g[6]};
    const double dv[3] = {x[7] - g[7], x[8] - g[8], x[9] - g[9]};
    e_pos = norm3(dp);
    e_vel = norm3(dv);
  }
  err[s] = e_ori;
  err[d.S + s] = e_pos;
  err[2 * d.S + s] = e_vel;
}

// segment offsets of the (metric, trial) segments inside err[3][S]
__global__ void k_err_segments(DevPtrs d, long long *seg_begin, long long *seg_end) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 3 * d.T)
    return;
  const int m = k / d.T, t = k - m * d.T;
  const long long b = (long long)m * d.S + d.trials[t].st_off;
  seg_begin[k] = b;
  seg_end[k] = b + ((d.lm[t].status == V2GT_OK) ? d.n_states[t] : 0);
}

__device__ __forceinline__ double warp_sum_fixed(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
    v += __shfl_down_sync(0xffffffffu, v, o);
  return __shfl_sync(0xffffffffu, v, 0);
}

// out: [T][3][8] = rmse, mean, median, min, max, std, ninetynine, count
__global__ void __launch_bounds__(32) k_err_stats(DevPtrs d, const double *__restrict__ sorted, const long long *seg_begin,
                                                  const long long *seg_end, double *__restrict__ out) {
  const int k = blockIdx.x; // = m * T + t
  const int m = k / d.T, t = k - m * d.T;
  const int lane = threadIdx.x;
  const double *v = sorted + seg_begin[k];
  const long long n = seg_end[k] - seg_begin[k];
  double *o = out + ((long long)t * 3 + m) * 8;
  if (n <= 0) {
    if (lane < 8)
      o[lane] = (lane == 7) ? 0.0 : NAN;
    return;
  }
  double s1 = 0.0, s2 = 0.0;
  for (long long i = lane; i < n; i += 32) {
    const double x = v[i];
    s1 += x;
    s2 += x * x;
  }
  s1 = warp_sum_fixed(s1);
  s2 = warp_sum_fixed(s2);
  const double mean = s1 / (double)n;
  const double rmse = sqrt(s2 / (double)n);
  double s3 = 0.0;
  for (long long i = lane; i < n; i += 32) {
    const double x = v[i] - mean;
    s3 += x * x;
  }
  s3 = warp_sum_fixed(s3);
  if (lane == 0) {
    const double sd = sqrt(s3 / (double)(n - 1)); // n == 1: 0 / 0 = NaN, as the reference computes it
    double median;
    if (n == 1)
      median = v[0];
    else if (n % 2 == 1)
      median = v[n / 2];
    else
      median = 0.5 * (v[n / 2 - 1] + v[n / 2]);
    o[0] = rmse;
    o[1] = mean;
    o[2] = median;
    o[3] = v[0];
    o[4] = v[n - 1];
    o[5] = sd;
    o[6] = mean + 2.326 * sd;
    o[7] = (double)n;
  }
}

// d_truth: [S][10] device; d_out: [T][3][8] device
int launch_error_stats(const DevPtrs &d, const double *d_truth, double *d_out, cudaStream_t st) {
  if (d.T > 0 && cudaMemsetAsync(d_out, 0, sizeof(double) * 24 * (size_t)d.T, st) != cudaSuccess) // "no samples" rows
    return V2GT_ERR_CUDA;
  if (d.S <= 0 || d.T <= 0)
    return V2GT_OK;
  double *err = nullptr, *sorted = nullptr;
  long long *segs = nullptr;
  void *tmp = nullptr;
  size_t tmp_bytes = 0;
  int rc = V2GT_OK;
  auto fail = [&](cudaError_t e) {
    if (e != cudaSuccess) {
      fprintf(stderr, "[vicon2gt_b200] CUDA error in error_stats: %s\n", cudaGetErrorString(e));
      rc = V2GT_ERR_CUDA;
    }
    return e != cudaSuccess;
  };
  do {
    if (fail(cudaMalloc(&err, sizeof(double) * 3 * d.S)) || fail(cudaMalloc(&sorted, sizeof(double) * 3 * d.S)) ||
        fail(cudaMalloc(&segs, sizeof(long long) * 6 * d.T)))
      break;
    k_err_values<<<(unsigned)((d.S + 127) / 128), 128, 0, st>>>(d, d_truth, err);
    k_err_segments<<<(3 * d.T + 127) / 128, 128, 0, st>>>(d, segs, segs + 3 * d.T);
    if (fail(cub::DeviceSegmentedSort::SortKeys(nullptr, tmp_bytes, err, sorted, (long long)3 * d.S, 3 * d.T, segs, segs + 3 * d.T, st)))
      break;
    if (fail(cudaMalloc(&tmp, tmp_bytes ? tmp_bytes : 1)))
      break;
    if (fail(cub::DeviceSegmentedSort::SortKeys(tmp, tmp_bytes, err, sorted, (long long)3 * d.S, 3 * d.T, segs, segs + 3 * d.T, st)))
      break;
    k_err_stats<<<3 * d.T, 32, 0, st>>>(d, sorted, segs, segs + 3 * d.T, d_out);
    if (fail(cudaGetLastError()) || fail(cudaStreamSynchronize(st)))
      break;
  } while (false);
  cudaFree(err);
  cudaFree(sorted);
  cudaFree(segs);
  cudaFree(tmp);
  return rc;
}

} // namespace v2
