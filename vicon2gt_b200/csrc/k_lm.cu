// vicon2gt_b200 kernel (5): device-resident Levenberg-Marquardt bookkeeping.
//
// Restates the GTSAM 4.2a5 policy the reference configures in ViconGraphSolver.cpp:547-567
// (LevenbergMarquardtOptimizer::iterate / tryLambda, NonlinearOptimizer::defaultOptimize,
// checkConvergence) as a per-trial state machine that lives in device memory.  Every trial of
// the batch advances by one damped solve ("try") per round; all decisions (accept / reject,
// lambda update, relinearise, stop) are taken on the device, the host only enqueues rounds.
#include "v2gt_common.cuh"
#include <float.h>
#include <stdio.h>

namespace v2 {

__device__ inline double sum_cost(const DevPtrs &d, int t, int which) {
  if (d.chain)
    return d.c_sums[which]; // all-reduced over the ranks of a chain-sharded solve
  const int n = d.n_states[t];
  const int used = (n + V2GT_CHUNK - 1) / V2GT_CHUNK;
  double s = 0;
  for (int k = 0; k < used; k++)
    s += d.part_cost[((long long)t * d.nchunk + k) * 4 + which];
  return s;
}

// one thread per trial; the cost partials of the CURRENT estimate must be in part_cost
__global__ void k_lm_begin(DevPtrs d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T)
    return;
  LmState &lm = d.lm[t];
  if (lm.frozen) // keeps the result and the counters of its last pass
    return;
  const v2gt_params &p = d.trials[t].prm;
  lm.lambda = p.lm_lambda_initial;
  lm.iterations = 0;
  lm.tries = 0;
  lm.linearizations = 0;
  lm.need_lin = 1;
  lm.solve_failed = 0;
  lm.n_trace = 0;
  lm.lm_state = 0;
  if (lm.status != V2GT_OK) {
    lm.lm_state = 2;
    return;
  }
  lm.error = sum_cost(d, t, 0);
  lm.initial_error = lm.error;
  lm.current_error = lm.error;
  if (lm.iterations >= p.lm_max_iterations) // defaultOptimize's early return
    lm.lm_state = 1;
}

// relinearisation loop of the batch (ViconGraphSolver.cpp:163-181 loops per solver): pass `pass` (0-based) only runs the
// trials whose num_loop_relin reaches it
__global__ void k_lm_freeze(DevPtrs d, int pass) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < d.T)
    d.lm[t].frozen = (d.trials[t].prm.num_loop_relin < pass) ? 1 : 0;
}

// before each try
__global__ void k_lm_pre_try(DevPtrs d) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T)
    return;
  LmState &lm = d.lm[t];
  if (lm.lm_state != 0)
    return;
  const v2gt_params &p = d.trials[t].prm;
  if (p.lm_max_tries > 0 && lm.tries >= p.lm_max_tries) {
    lm.lm_state = 4;
    return;
  }
  if (lm.need_lin) {
    lm.current_error = lm.error; // currentError = error() at the top of defaultOptimize's loop
    lm.linearizations++;
  }
  lm.tries++;
  lm.solve_failed = 0;
}

// after each try: tryLambda's decision + the loop condition of defaultOptimize
__global__ void k_lm_decide(DevPtrs d, int count_active) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= d.T)
    return;
  LmState &lm = d.lm[t];
  if (lm.lm_state != 0)
    return;
  const v2gt_params &p = d.trials[t].prm;
  bool step_ok = false, stop_search = false;
  double new_error = INFINITY, lin_change = 0.0, cost_change = 0.0;
  const bool failed = d.chain ? (d.c_sums[4] > 0.0) : (lm.solve_failed != 0);
  if (!failed) {
    // g.delta and |delta|^2: the states' parts come from the trial-point evaluation, the globals' from the reduced solve
    const double *xg = d.xg + (long long)t * 16;
    const double gd = xg[10] + sum_cost(d, t, 2), dd = xg[11] + sum_cost(d, t, 3);
    // linear.error(0) - linear.error(delta) = g.delta - 1/2 delta' H delta = 1/2 (g.delta + lambda |delta|^2)
    lin_change = 0.5 * (gd + lm.lambda * dd);
    if (lin_change >= 0.0) {
      new_error = sum_cost(d, t, 0); // partials of the trial point
      cost_change = lm.error - new_error;
      if (lin_change > DBL_EPSILON * lm.lin_error0) {
        const double fidelity = cost_change / lin_change;
        step_ok = fidelity > p.lm_min_model_fidelity;
      }
      const double min_abs_tol = p.lm_relative_error_tol * lm.error;
      if (fabs(cost_change) < min_abs_tol)
        stop_search = true;
    }
  }
  if (lm.n_trace < TRACE_MAX) {
    double *tr = d.trace + ((long long)t * TRACE_MAX + lm.n_trace) * 5;
    tr[0] = lm.lambda;
    tr[1] = lm.error;
    tr[2] = new_error;
    tr[3] = lin_change;
    tr[4] = step_ok ? 1.0 : 0.0;
    lm.n_trace++;
  }
  bool end_iterate = false, gave_up = false;
  if (step_ok) {
    lm.cur = 1 - lm.cur; // the trial point becomes the estimate
    lm.error = new_error;
    lm.lambda = fmax(p.lm_lambda_lower_bound, lm.lambda / p.lm_lambda_factor);
    lm.iterations++;
    end_iterate = true;
  } else if (!stop_search) {
    lm.lambda *= p.lm_lambda_factor;
    if (lm.lambda >= p.lm_lambda_upper_bound) {
      gave_up = true;
      end_iterate = true;
    }
  } else {
    end_iterate = true;
  }
  lm.need_lin = 0;
  if (end_iterate) {
    lm.need_lin = 1;
    const double cur_err = lm.current_error, new_err = lm.error;
    if (lm.iterations >= p.lm_max_iterations) {
      lm.lm_state = 1;
    } else if (new_err <= 0.0) {
      lm.lm_state = 2;
    } else {
      const double abs_dec = cur_err - new_err;
      const double rel_dec = abs_dec / cur_err;
      const bool converged = (p.lm_relative_error_tol != 0.0 && rel_dec <= p.lm_relative_error_tol) ||
                             (abs_dec <= p.lm_absolute_error_tol);
      if (converged)
        lm.lm_state = gave_up ? 3 : 2;
      else if (!isfinite(cur_err))
        lm.lm_state = 2;
    }
  }
  if (count_active && lm.lm_state == 0)
    atomicAdd(d.n_active, 1);
}

// ---- chain-sharded mode helpers (single trial)
// JPLNavState::retract of one state row (src/gtsam/JPLNavState.cpp:25-54)
__device__ inline void retract_row(const double *x0, const double *dx, double *x1) {
  double dq[4], q[4];
  small_quat(dx, dq);
  quat_mul(dq, x0, q);
  x1[0] = q[0]; x1[1] = q[1]; x1[2] = q[2]; x1[3] = q[3];
  for (int k = 0; k < 12; k++)
    x1[4 + k] = x0[4 + k] + dx[3 + k];
}
// This rank's contribution to the second exchange of a damped solve (CSUM_N doubles, all-gathered): the local sums of
// the last factor evaluation [cost, lin_error0, g.delta, |delta|^2, solve_failed] and, at CSUM_HALO, the step of its
// last own state -- the next rank's left halo (with_halo; the evaluation of the current estimate has no step)
__global__ void k_chain_sum(DevPtrs d, int with_halo) {
  const int k = threadIdx.x;
  if (blockIdx.x != 0 || k >= CSUM_N)
    return;
  double s = 0;
  if (k < 4) {
    const int n = d.n_states[0];
    const int used = (n + V2GT_CHUNK - 1) / V2GT_CHUNK;
    for (int c = 0; c < used; c++)
      s += d.part_cost[(long long)c * 4 + k];
  } else if (k == 4) {
    s = d.lm[0].solve_failed ? 1.0 : 0.0;
  } else if (k >= CSUM_HALO && with_halo && d.c_hi >= 1) {
    s = d.delta[(d.trials[0].st_off + d.c_hi - 1) * 16 + (k - CSUM_HALO)];
  }
  d.c_sums_send[k] = s;
}
// after the all-gather of the above: sums in rank order (bit-identical on every rank) -> c_sums; the left halo (the
// previous rank's last own state, an interior state of its last segment) takes its step and its trial point.  The right
// halo is a separator: the redundant reduced solve has already written its step.
__global__ void k_chain_reduce(DevPtrs d, int with_halo) {
  const int k = threadIdx.x;
  if (blockIdx.x != 0)
    return;
  if (k < 5) {
    double s = 0;
    for (int r = 0; r < d.c_world; r++)
      s += d.c_sums_all[r * CSUM_N + k];
    d.c_sums[k] = s;
  }
  if (k == 31 && with_halo && d.c_lo >= 1 && d.c_rank >= 1) {
    const long long s = d.trials[0].st_off + d.c_lo - 1;
    const double *dl = d.c_sums_all + (d.c_rank - 1) * CSUM_N + CSUM_HALO;
    for (int e = 0; e < 16; e++)
      d.delta[s * 16 + e] = dl[e];
    const int cur = d.lm[0].cur;
    retract_row(d.x + ((long long)cur * d.S + s) * X_STRIDE, dl, d.x + ((long long)(1 - cur) * d.S + s) * X_STRIDE);
  }
}
// pack what the reduced (separator) system needs from this rank: its segments' outputs, the records of its separators
// and its part of the global block.  grid: own segments + 1
__global__ void k_chain_pack(DevPtrs d, int stride) {
  const int ploc = d.c_j1 - d.c_j0;
  const int jl = blockIdx.x; // local segment index
  if (jl == ploc) {
    double *go = d.c_pack_send + (long long)ploc * (SEG_STRIDE + NE_STRIDE);
    for (int k = threadIdx.x; k < 96; k += blockDim.x)
      go[k] = d.c_gamma_loc[k];
    return;
  }
  if (jl > ploc)
    return;
  const int j = d.c_j0 + jl;
  const double *sg = d.seg + (long long)j * SEG_STRIDE;
  double *so = d.c_pack_send + (long long)jl * SEG_STRIDE;
  for (int k = threadIdx.x; k < SEG_STRIDE; k += blockDim.x)
    so[k] = sg[k];
  double *no = d.c_pack_send + (long long)ploc * SEG_STRIDE + (long long)jl * NE_STRIDE;
  const long long slot = (long long)j * stride - 1 - d.c_goff; // separator in front of segment j (none for j = 0)
  for (int k = threadIdx.x; k < NE_STRIDE; k += blockDim.x)
    no[k] = (j >= 1 && slot >= 0 && slot < d.n_states[0]) ? d.ne[(d.trials[0].st_off + slot) * NE_STRIDE + k] : 0.0;
}
// after the all-gather: every rank's contribution into the globally indexed buffers; the global block is summed in
// rank order (identical on every rank).  grid: all segments of the chain + 1
__global__ void k_chain_unpack(DevPtrs d) {
  const int ploc = d.c_j1 - d.c_j0; // the same on every rank
  const long long per_rank = (long long)ploc * (SEG_STRIDE + NE_STRIDE) + 96;
  const int j = blockIdx.x;
  if (j == d.P) {
    for (int k = threadIdx.x; k < 96; k += blockDim.x) {
      double v = 0.0;
      for (int r = 0; r < d.c_world; r++)
        v += d.c_pack_all[r * per_rank + (long long)ploc * (SEG_STRIDE + NE_STRIDE) + k];
      d.gamma[k] = v;
      if (k == 90)
        d.lm[0].lin_error0 = v; // the summed linear error
    }
    return;
  }
  if (j > d.P)
    return;
  const int r = j / ploc, jl = j - r * ploc;
  const double *base = d.c_pack_all + r * per_rank;
  const double *si = base + (long long)jl * SEG_STRIDE;
  const double *ni = base + (long long)ploc * SEG_STRIDE + (long long)jl * NE_STRIDE;
  double *so = d.seg + (long long)j * SEG_STRIDE;
  double *no = d.c_sepne + (long long)j * NE_STRIDE;
  for (int k = threadIdx.x; k < SEG_STRIDE; k += blockDim.x)
    so[k] = si[k];
  for (int k = threadIdx.x; k < NE_STRIDE; k += blockDim.x)
    no[k] = ni[k];
}
int launch_chain_sum(const DevPtrs &d, int with_halo, cudaStream_t st) {
  k_chain_sum<<<1, 32, 0, st>>>(d, with_halo);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_chain_reduce(const DevPtrs &d, int with_halo, cudaStream_t st) {
  k_chain_reduce<<<1, 32, 0, st>>>(d, with_halo);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_chain_unpack(const DevPtrs &d, cudaStream_t st) {
  k_chain_unpack<<<d.P + 1, 256, 0, st>>>(d);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_chain_pack(const DevPtrs &d, int stride, cudaStream_t st) {
  k_chain_pack<<<d.c_j1 - d.c_j0 + 1, 256, 0, st>>>(d, stride);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_lm_freeze(const DevPtrs &d, int pass, cudaStream_t st) {
  k_lm_freeze<<<(d.T + 127) / 128, 128, 0, st>>>(d, pass);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_lm_begin(const DevPtrs &d, cudaStream_t st) {
  k_lm_begin<<<(d.T + 127) / 128, 128, 0, st>>>(d);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_lm_pre_try(const DevPtrs &d, cudaStream_t st) {
  k_lm_pre_try<<<(d.T + 127) / 128, 128, 0, st>>>(d);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}
int launch_lm_decide(const DevPtrs &d, int count_active, cudaStream_t st) {
  k_lm_decide<<<(d.T + 127) / 128, 128, 0, st>>>(d, count_active);
  V2_CUDA_CHECK(cudaGetLastError());
  return V2GT_OK;
}

} // namespace v2
