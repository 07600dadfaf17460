/*
 * vicon2gt_b200 -- C ABI of the B200-native batch solver for vicon2gt's
 * mocap + IMU ground-truth optimisation (ViconGraphSolver::build_and_solve).
 *
 * The reference (rpng/vicon2gt) has no FFI: its hot path sits behind the C++
 * class ViconGraphSolver (src/solver/ViconGraphSolver.h:61-168), filled through
 * Propagator::feed_imu (src/meas/Propagator.h:49) and Interpolator::feed_pose
 * (src/meas/Interpolator.h:55).  This header is the boundary a maintainer binds
 * instead: plain pointers and sizes, no C++/torch types,
 * every function returns an int status (0 = ok) and never exits the process.
 *
 * Conventions (same as the reference): quaternions are JPL, vector first
 * [x y z w]; R = quat_2_Rot(q) maps global->local; 3x3 matrices are row-major;
 * a state row is 16 doubles [q(4) bg(3) v(3) ba(3) p(3)]; the error state order
 * is [theta bg v ba p]; all arithmetic is fp64.
 */
#ifndef VICON2GT_B200_H
#define VICON2GT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define V2GT_STATE_DIM 16   /* q4 bg3 v3 ba3 p3 */
#define V2GT_ERR_DIM 15     /* theta bg v ba p */
#define V2GT_GLOBAL_DIM 9   /* C0 theta(3) | C1 p(3) | T(1) | G thetax,thetay(2) */
#define V2GT_PREINT_DIM 288 /* see v2gt_batch_get_preint */

/* status codes */
enum {
  V2GT_OK = 0,
  V2GT_ERR_INVALID_ARG = 1,
  V2GT_ERR_CUDA = 2,
  V2GT_ERR_NO_TIMESTAMPS = 3,    /* ViconGraphSolver.cpp:99-104  (reference: exit)  */
  V2GT_ERR_NO_VICON = 4,         /* ViconGraphSolver.cpp:107-112 (reference: exit)  */
  V2GT_ERR_ALL_OUT_OF_RANGE = 5, /* ViconGraphSolver.cpp:145-150 (reference: exit)  */
  V2GT_ERR_INVALID_PREINT = 6,   /* ViconGraphSolver.cpp:508-512 (reference: exit)  */
  V2GT_ERR_NAN_COVARIANCE = 7,   /* ViconGraphSolver.cpp:520-524 (reference: exit)  */
  V2GT_ERR_NOT_BUILT = 8,
  V2GT_ERR_NO_DEVICE = 9,
  V2GT_ERR_NCCL = 10
};

/*
 * ViconGraphSolverParams: replaces the ros::NodeHandle the reference reads its
 * parameters from (names and defaults: ViconGraphSolver.cpp:34-75), plus the
 * IMU noise densities held by Propagator (Propagator.h:41-46), the GTSAM
 * Levenberg-Marquardt settings of ViconGraphSolver.cpp:547-555 (GTSAM 4.2a5
 * defaults for the rest) and the reference's hard-coded constants.
 */
typedef struct v2gt_params {
  double R_GtoV[9];                   /* init gravity->vicon rotation, row-major, default I  (.cpp:34-39) */
  double gravity_magnitude;           /* 9.81                                                 (.cpp:42)    */
  double R_BtoI[9];                   /* init marker-body->IMU rotation, default I            (.cpp:45-49) */
  double p_BinI[3];                   /* init marker-body position in IMU, default 0          (.cpp:51-54) */
  double toff_imu_to_vicon;           /* init time offset, default 0                          (.cpp:57)    */
  int32_t estimate_toff_vicon_to_imu; /* default 1                                            (.cpp:70)    */
  int32_t estimate_ori_vicon_to_imu;  /* default 1                                            (.cpp:71)    */
  int32_t estimate_pos_vicon_to_imu;  /* default 1                                            (.cpp:72)    */
  int32_t num_loop_relin;             /* default 0                                            (.cpp:75)    */
  double sigma_w;                     /* gyroscope_noise_density      1.6968e-4  (run_simulation.cpp:79)   */
  double sigma_wb;                    /* gyroscope_random_walk        1.9393e-5                            */
  double sigma_a;                     /* accelerometer_noise_density  2.0e-3                               */
  double sigma_ab;                    /* accelerometer_random_walk    3.0e-3                               */
  int32_t lm_max_iterations;          /* 30                                                   (.cpp:555)   */
  int32_t lm_max_tries;               /* safety cap on damped solves per optimise call; 0 = unlimited      */
  double lm_lambda_initial;           /* 1e-5  (GTSAM default)  */
  double lm_lambda_factor;            /* 10    (GTSAM default)  */
  double lm_lambda_upper_bound;       /* 1e20                                                 (.cpp:554)   */
  double lm_lambda_lower_bound;       /* 0     (GTSAM default)  */
  double lm_min_model_fidelity;       /* 1e-3  (GTSAM default)  */
  double lm_relative_error_tol;       /* 1e-30                                                (.cpp:553)   */
  double lm_absolute_error_tol;       /* 1e-30                                                (.cpp:552)   */
  double huber_k;                     /* 1.345 (MeasBased_ViconPoseTimeoffsetFactor.h:51)     */
  double time_offset_window;          /* TIME_OFFSET 0.25 s (ViconGraphSolver.h:167)          */
  double interp_gap_init;             /* 5.0 s  (Interpolator.cpp:120)                        */
  double interp_gap_factor;           /* 0.1 s  (Interpolator.cpp:232)                        */
} v2gt_params;

/* Fill `p` with the reference defaults listed above. */
void v2gt_params_default(v2gt_params *p);

/* Per-trial outcome (replaces the reference's console log + std::exit). */
typedef struct v2gt_trial_info {
  int32_t status;             /* V2GT_OK or one of the V2GT_ERR_* codes for this trial              */
  int32_t n_states;           /* surviving states after cleaning / erasure                          */
  int32_t removed_no_imu;     /* ct_remove_imu      (ViconGraphSolver.cpp:127-130)                  */
  int32_t removed_before;     /* ct_remove_before   (:131-134)                                      */
  int32_t removed_after;      /* ct_remove_after    (:135-138)                                      */
  int32_t removed_no_vicon;   /* states erased in build_problem (:443-461)                          */
  int32_t iterations;         /* accepted LM iterations (optimizer.iterations())                    */
  int32_t tries;              /* damped solves (lambda tries)                                       */
  int32_t linearizations;     /* outer iterations started                                           */
  int32_t lm_state;           /* 0 running, 1 max iterations, 2 converged (no decrease), 3 lambda gave up, 4 try cap */
  double initial_error;       /* graph.error(values) before optimisation                            */
  double final_error;         /* total cost: sum 1/2 e'P^-1 e + sum huber(|We|)                     */
  double lambda;              /* current lambda                                                     */
} v2gt_trial_info;

typedef struct v2gt_batch v2gt_batch; /* opaque; owns all device memory; one host thread per handle */

/* Library / device queries. */
const char *v2gt_version(void);
int v2gt_device_count(int32_t *count);
/* sizeof(v2gt_params), sizeof(v2gt_trial_info) as compiled: lets a binding check its struct mirrors. */
int v2gt_abi_struct_sizes(int32_t *params_bytes, int32_t *trial_info_bytes);

/* Create a batch of `n_trials` independent problems on CUDA device `device`. */
int v2gt_batch_create(int32_t device, int32_t n_trials, v2gt_batch **out);
int v2gt_batch_destroy(v2gt_batch *b);

/*
 * Stage one trial's inputs (HOST pointers; copied).  Replaces
 *   Propagator::feed_imu(t, wm, am)            x M   (src/meas/Propagator.cpp:21-31)
 *   Interpolator::feed_pose(t, q, p, R_q, R_p) x V   (src/meas/Interpolator.cpp:21-38)
 *   the timestamp_cameras ctor argument              (src/solver/ViconGraphSolver.cpp:20-27)
 * imu_w/imu_a: [M][3]; vic_q: [V][4]; vic_p: [V][3]; vic_Rq/vic_Rp: [V][9] row-major,
 * or NULL with vic_R_const = 18 doubles (R_q, R_p) applied to every pose.
 * Vicon poses are sorted by time and de-duplicated (first one wins) exactly as the
 * reference's std::set does; IMU samples are kept in feed order.
 */
int v2gt_batch_set_trial(v2gt_batch *b, int32_t trial, const v2gt_params *params,
                         int64_t n_imu, const double *imu_t, const double *imu_w, const double *imu_a,
                         int64_t n_vicon, const double *vic_t, const double *vic_q, const double *vic_p,
                         const double *vic_Rq, const double *vic_Rp, const double *vic_R_const,
                         int64_t n_times, const double *state_t);

/*
 * The same, without the host copy: the arrays are BORROWED and must stay valid and unchanged until the next
 * v2gt_batch_upload() (or v2gt_batch_build_and_solve() on a batch that was not uploaded yet) has returned.  The
 * shared_ptr<Propagator> / shared_ptr<Interpolator> stores of the reference (ViconGraphSolver.h:117-120) have exactly
 * this lifetime.  The conversion to the device layout is done by v2gt_batch_upload, on several host threads, straight
 * into pinned staging memory.
 */
int v2gt_batch_set_trial_view(v2gt_batch *b, int32_t trial, const v2gt_params *params,
                              int64_t n_imu, const double *imu_t, const double *imu_w, const double *imu_a,
                              int64_t n_vicon, const double *vic_t, const double *vic_q, const double *vic_p,
                              const double *vic_Rq, const double *vic_Rp, const double *vic_R_const,
                              int64_t n_times, const double *state_t);

/* Clean timestamps (ViconGraphSolver.cpp:114-159) and copy every staged trial to the device. */
int v2gt_batch_upload(v2gt_batch *b);

/* ViconGraphSolver::build_problem (.cpp:390-536): initial guess (first call), state erasure,
 * continuous preintegration of every interval.  `init_states` as in the reference. */
int v2gt_batch_build(v2gt_batch *b, int32_t init_states);
/* ViconGraphSolver::optimize_problem (.cpp:538-570): device-resident Levenberg-Marquardt. */
int v2gt_batch_optimize(v2gt_batch *b);
/* ViconGraphSolver::build_and_solve (.cpp:96-192): upload (if needed), then
 * for i in 0..num_loop_relin: build(i==0); optimize.  Blocks until the batch is done. */
int v2gt_batch_build_and_solve(v2gt_batch *b);
/* Block until all queued device work of this handle has finished. */
int v2gt_batch_sync(v2gt_batch *b);

/* Results (HOST pointers).  get_states replaces get_imu_poses (.cpp:357-375): times[n], states[n][16]. */
int v2gt_batch_get_info(v2gt_batch *b, int32_t trial, v2gt_trial_info *info);
int v2gt_batch_get_states(v2gt_batch *b, int32_t trial, int64_t capacity, double *times, double *states, int64_t *n_out);
/* The same for every trial of the batch in one pass (asynchronous copies through pinned memory, moved to the caller's
 * arrays on several host threads): offsets[n_trials + 1] = prefix sums of the surviving-state counts, times[offsets[T]],
 * states[offsets[T]][16], globals10[n_trials][10] = q_BtoI(4) p_BinI(3) t_off theta_x theta_y.  `capacity` = states the
 * arrays can hold; with times = states = globals10 = NULL only the offsets are filled. */
int v2gt_batch_get_all_states(v2gt_batch *b, int64_t capacity, int64_t *offsets, double *times, double *states, double *globals10);
/* get_calibration (.cpp:377-388): q_BtoI[4], p_BinI[3], t_off, (theta_x, theta_y) of R_GtoV. */
int v2gt_batch_get_calibration(v2gt_batch *b, int32_t trial, double *q_BtoI, double *p_BinI, double *t_off, double *theta_xy);
/* ViconGraphSolver::write_to_file (.cpp:194-248): byte-compatible CSV + info text. */
int v2gt_batch_write_to_file(v2gt_batch *b, int32_t trial, const char *csv_path, const char *info_path);
/* The same for every solved trial of the batch (one device read-back, files written on several host threads):
 * <dir>/<trial, 5 digits>_states.csv and _info.txt; with_txt != 0 adds _states.txt, the pose text file
 * ("timestamp(s) tx ty tz qx qy qz qw") the reference's Monte-Carlo script hands to its evaluation tool after
 * converting the CSV (scripts/run_monte_carlo.sh:66-72). */
int v2gt_batch_write_all(v2gt_batch *b, const char *dir, int32_t with_txt, int32_t *n_written);

/* Batched trajectory-error report of a simulated batch (the per-run report of src/run_simulation.cpp:198-297 with
 * src/utils/stats.h:64-112, for all trials at once on the device).
 * set_truth: truth10 = [n][10] = q_VtoI (JPL, 4), p_IinV (3), v_IinV (3) from Simulator::get_state_in_vicon at the n
 * surviving state times of the trial (v2gt_batch_get_states).
 * error_stats: stats = [n_trials][3][8] with rows (orientation deg, position m, velocity m/s) and columns
 * rmse, mean, median, min, max, std, mean + 2.326 std, count; NaN rows for trials that did not solve. */
int v2gt_batch_set_truth(v2gt_batch *b, int32_t trial, int64_t n, const double *truth10);
int v2gt_batch_error_stats(v2gt_batch *b, double *stats);

/* ---- evaluation-only entry points (parity tests; they run the production kernels) ---- */

/* Interpolator::get_pose / get_pose_with_jacobian for n query times against trial's vicon store.
 * Any output pointer may be NULL.  idx0/idx1/ok: [n]; q [n][4]; p [n][3]; R [n][36]; H_toff [n][6]. */
int v2gt_eval_interp(v2gt_batch *b, int32_t trial, int64_t n, const double *t_query, double gap_thresh,
                     int32_t *idx0, int32_t *idx1, int32_t *ok, double *q, double *p, double *R, double *H_toff);
/* Preintegration of interval k -> k+1 for every k (after build).  Row layout (V2GT_PREINT_DIM = 288):
 * [0] DT, [1:4] alpha, [4:7] beta, [7:11] q_k2tau, [11:14] bg_lin, [14:17] ba_lin, [17:26] J_q,
 * [26:35] J_beta(J_b), [35:44] J_alpha(J_a), [44:53] H_beta(H_b), [53:62] H_alpha(H_a), [62] n_steps,
 * [63:288] information matrix P^-1 (15x15 row-major). */
int v2gt_batch_get_preint(v2gt_batch *b, int32_t trial, int64_t capacity, double *out, int64_t *n_out);
/* Preintegrated covariance P_meas (15x15 row-major) per interval. */
int v2gt_batch_get_preint_cov(v2gt_batch *b, int32_t trial, int64_t capacity, double *out, int64_t *n_out);
/* Overwrite the current estimate (states [n][16], globals q_BtoI[4] p_BinI[3] t_off thetax thetay = 10 doubles). */
int v2gt_batch_set_estimate(v2gt_batch *b, int32_t trial, int64_t n, const double *states, const double *globals);
/* Linearise at the current estimate: normal-equation blocks per state
 * D [n][15*15], O [n][15*15] (coupling k,k+1; last row zero), B [n][15*9], g [n][15],
 * Gamma [81], g_gamma [9], plus cost (nonlinear) and linear error at zero. Any pointer may be NULL. */
int v2gt_eval_linearize(v2gt_batch *b, int32_t trial, int64_t capacity, double *D, double *O, double *B, double *g,
                        double *Gamma, double *g_gamma, double *cost, double *lin_error0);
/* Solve (H + lambda I) delta = g for the blocks of the last linearisation; delta: [15 n + 9]. */
int v2gt_eval_solve(v2gt_batch *b, int32_t trial, double lambda, int64_t capacity, double *delta, int32_t *ok);
/* Total cost at the current estimate. */
int v2gt_eval_cost(v2gt_batch *b, int32_t trial, double *cost);
/* LM trace of the last optimise call: rows [lambda, cost_before, cost_after, lin_cost_change, accepted]. */
int v2gt_batch_get_lm_trace(v2gt_batch *b, int32_t trial, int64_t capacity, double *rows, int64_t *n_out);

/* Evaluation-only (host logic, works without a device): the chain cut the optimiser picks for trials of n_states
 * states when n_running trials are still being optimised -- segments per trial at level 1 (close to target_warps /
 * n_running, at most 4 sqrt(n), at least 32 states long, filling whole waves of the sm_count x resident-CTA slots) and
 * at level 2 (segments of the separator chain; 0 = the separator chain is walked directly). */
int v2gt_plan_segments(int32_t n_states, int32_t n_running, int32_t target_warps, int32_t sm_count, int32_t *p_level1, int32_t *p_level2);

/* ---- chain-sharded solve of ONE long sequence across ranks (BASELINE config 5) ----
 * Replaces, for a sequence too long for one GPU's latency budget, the single graph that
 * ViconGraphSolver::build_problem assembles (src/solver/ViconGraphSolver.cpp:390-536): the cleaned state chain is cut
 * into world x segs_per_rank segments; every rank creates a 1-trial batch holding its contiguous piece of the chain
 * plus one halo state on each inner side (the IMU factor across a rank boundary reads the neighbour's state,
 * src/gtsam/ImuFactorCPIv1.cpp:29-39).  Per damped solve the ranks exchange twice, both all-gathers in rank order:
 *   A  PACK_SEND -> PACK_ALL   per rank: Schur outputs of its segments, normal-equation records of its separators,
 *                              its part of the 9x9 global block + gradient (summed in rank order after the gather)
 *   B  SUMS_SEND -> SUMS_ALL   per rank: V2GT_CHAIN_SUMS_N doubles = cost / model-fidelity scalars of its own
 *                              factors + the step of its last own state (the next rank's left halo)
 * B cannot ride on A: the accept / reject decision needs the summed cost of the trial point, which needs the step,
 * which needs the reduced system gathered by A.  The reduced (separator + 9 globals) system is solved redundantly on
 * every rank from bit-identical data, so all ranks take identical Levenberg-Marquardt decisions.
 * Geometry: the chain of n_total states is cut into p_total segments with the same formula as the single-GPU solve;
 * this rank owns segments [j0, j1) and the separators in front of them, i.e. the local slots [own_lo, own_hi) of its
 * n_local states starting at global state g_off (v2gt_chain_plan computes all of it).  Call configure before upload.
 *
 * Two ways to drive it:
 *   v2gt_chain_optimize   the whole Levenberg-Marquardt loop inside the library: both exchanges are ncclAllGather calls
 *                         on the handle's stream (libnccl is loaded at run time; v2gt_chain_comm_init), the kernel +
 *                         collective sequence of one damped solve is captured once as a CUDA graph and replayed; the
 *                         host reads one counter every `tries_per_sync` damped solves.  world == 1 needs no NCCL.
 *   v2gt_chain_phase      phase by phase, the caller moves the buffers itself between the phases (used to run all
 *                         ranks of a small world on ONE device in the tests: same kernels, same buffers). */
typedef struct v2gt_chain_plan_t {
  int32_t rank, world;
  int32_t p_total, stride;   /* segments of the whole chain; separators sit at m * stride - 1 */
  int32_t j0, j1;            /* own segments [j0, j1) */
  int32_t own_lo, own_hi;    /* own local slots */
  int64_t n_total, g_off, n_local;
} v2gt_chain_plan_t;
/* Host only. Ownership of a chain of n_total cleaned states for `rank` of `world` with segs_per_rank segments each;
 * V2GT_ERR_INVALID_ARG when the chain is too short for that cut.  v2gt_chain_fit_segments: the largest count
 * <= target that is valid (0 if none). */
int v2gt_chain_plan(int64_t n_total, int32_t world, int32_t segs_per_rank, int32_t rank, v2gt_chain_plan_t *plan);
int v2gt_chain_fit_segments(int64_t n_total, int32_t world, int32_t target, int32_t *segs_per_rank);
#define V2GT_CHAIN_SUMS_N 24
#define V2GT_NCCL_ID_BYTES 128
enum {
  V2GT_CHAIN_BUF_SUMS = 0, V2GT_CHAIN_BUF_GAMMA = 1, V2GT_CHAIN_BUF_PACK_SEND = 2, V2GT_CHAIN_BUF_SEG_ALL = 3,
  V2GT_CHAIN_BUF_PACK_ALL = 4, V2GT_CHAIN_BUF_SEPNE_ALL = 5, V2GT_CHAIN_BUF_SUMS_SEND = 6, V2GT_CHAIN_BUF_SUMS_ALL = 7
};
enum {
  V2GT_CHAIN_PHASE_COST0 = 0,   /* cost of the current estimate -> SUMS_SEND            [exchange B]      */
  V2GT_CHAIN_PHASE_BEGIN = 1,   /* sums in rank order, LM state initialisation                            */
  V2GT_CHAIN_PHASE_LINEARIZE = 2, /* start of a try; relinearise if the last step was accepted            */
  V2GT_CHAIN_PHASE_ELIMINATE = 3, /* eliminate the own segments -> PACK_SEND             [exchange A]      */
  V2GT_CHAIN_PHASE_SOLVE = 4,   /* unpack PACK_ALL, reduced system (redundant), back-substitution         */
  V2GT_CHAIN_PHASE_COST = 5,    /* trial point + its cost, boundary step -> SUMS_SEND   [exchange B]      */
  V2GT_CHAIN_PHASE_DECIDE = 6   /* sums in rank order, left halo, accept / reject, lambda update          */
};
int v2gt_chain_configure(v2gt_batch *b, int32_t rank, int32_t world, int64_t n_total, int64_t g_off, int32_t p_total, int32_t j0,
                         int32_t j1, int32_t own_lo, int32_t own_hi);
int v2gt_chain_phase(v2gt_batch *b, int32_t phase);
/* device pointer + length (doubles) of one of the exchange buffers (valid after upload) */
int v2gt_chain_buffer(v2gt_batch *b, int32_t which, void **dev_ptr, int64_t *n_doubles);
/* NCCL inside the library.  unique_id: rank 0 calls it and hands the V2GT_NCCL_ID_BYTES bytes to every rank by any
 * means (a file, MPI, torch.distributed); comm_init: every rank, after v2gt_chain_configure (collective: returns when
 * all `world` ranks have joined).  V2GT_ERR_NCCL when libnccl.so.2 cannot be loaded or a call fails. */
int v2gt_chain_nccl_unique_id(uint8_t *id_bytes);
int v2gt_chain_comm_init(v2gt_batch *b, const uint8_t *id_bytes);
/* The Levenberg-Marquardt loop of the sharded chain (after v2gt_batch_build): at most max_tries damped solves
 * (<= 0: the policy's own bound).  use_graph: 1 = replay one captured CUDA graph per damped solve, 0 = plain launches
 * (per-phase CUDA-event timing in v2gt_batch_get_timing).  Collective when world > 1. */
int v2gt_chain_optimize(v2gt_batch *b, int32_t max_tries, int32_t use_graph);
/* the handle's CUDA stream (cudaStream_t), so that the caller's collectives can be ordered with the kernels */
int v2gt_batch_stream(v2gt_batch *b, void **stream);

/* ---- batched simulator on the device ----
 * src/sim of the reference (Simulator.cpp:21-262, BsplineSE3.cpp:21-351) and the measurement loop + state grid of
 * src/run_simulation.cpp:108-158 for ALL trials of a Monte-Carlo batch at once.  One trajectory per batch; trial t takes
 * a v2gt_sim_params (vicon2gt_b200/host/v2gt_sim.h, the SimulatorParams mirror of the host simulator; default: seed = t);
 * the rates, gravity magnitude, sample cap and state rate must be the same for every trial of a batch.  Each seed's
 * streams are those of a GCC build of the reference: std::mt19937(seed) words through libstdc++'s polar
 * normal_distribution, regenerated and compacted in parallel on the device (csrc/k_sim.cu).
 *   plan  host only (works without a device): clock schedule, truth per seed, vicon stamps, state-time grids
 *   run   device: spline evaluation once per stamp, noise streams, bias random walks, measurements
 *   get   caller layout of v2gt_sim_get (any pointer may be NULL; stamps only need plan)
 *   get_state_in_vicon   Simulator::get_state_in_vicon (Simulator.cpp:91-133) for n query times of one trial
 * v2gt_batch_set_trial_sim stages trial `sim_trial` of a simulated batch as trial `trial` of a solver batch on the same
 * device: the stamps are cleaned on the host as usual, the measurement rows go device -> device (no host round trip). */
struct v2gt_sim_params;
typedef struct v2gt_simbatch v2gt_simbatch;
int v2gt_simbatch_create(int32_t device, int32_t n_trials, int64_t n_rows, const double *rows8, v2gt_simbatch **out);
int v2gt_simbatch_destroy(v2gt_simbatch *s);
int v2gt_simbatch_set_trial(v2gt_simbatch *s, int32_t trial, const struct v2gt_sim_params *p);
int v2gt_simbatch_plan(v2gt_simbatch *s);
int v2gt_simbatch_run(v2gt_simbatch *s);
int v2gt_simbatch_sizes(v2gt_simbatch *s, int32_t trial, int64_t *n_imu, int64_t *n_vicon, int64_t *n_states);
int v2gt_simbatch_get(v2gt_simbatch *s, int32_t trial, double *imu_t, double *imu_w, double *imu_a, double *vic_t, double *vic_q,
                      double *vic_p, double *state_t);
int v2gt_simbatch_get_truth(v2gt_simbatch *s, int32_t trial, double *R_BtoI, double *p_BinI, double *viconimu_dt, double *R_GtoV);
int v2gt_simbatch_get_state_in_vicon(v2gt_simbatch *s, int32_t trial, int64_t n, const double *times, double *out17, int32_t *ok);
/* device time (ms, CUDA events) of the last run: uploads of the plan + all generation kernels */
int v2gt_simbatch_get_timing(v2gt_simbatch *s, double *ms_generate);
int v2gt_batch_set_trial_sim(v2gt_batch *b, int32_t trial, const v2gt_params *params, v2gt_simbatch *s, int32_t sim_trial,
                             const double *vic_R_const);

/* ---- measurement helpers ---- */
/* Device time (ms, CUDA events on the handle's stream) spent in each phase of the last
 * build/optimise: [0] build(init+preint) [1] linearise+assemble [2] eliminate [3] reduced solve
 * [4] back-substitute+retract [5] cost evaluation [6] LM bookkeeping; launches[7] kernel launch counts. */
int v2gt_batch_get_timing(v2gt_batch *b, double *ms, int64_t *launches);
/* Options (before upload): "segments" (chain segments per trial, 0 = chosen per optimiser round), "target_warps" (segment
 * warps an elimination launch aims for), "keep_cov" (1: keep the preintegrated covariances for v2gt_batch_get_preint_cov),
 * "tries_per_sync" (damped solves between two host reads of the running-trial counter, default 8), "graph" (-1: replay a
 * captured CUDA graph per damped solve for small problems, 0: never, 1: always). */
int v2gt_batch_set_option(v2gt_batch *b, const char *name, double value);
/* Measured FP64 FMA peak of `device` in TFLOP/s (dependent-chain-free DFMA loop, best of 4): the FP64
 * roofline denominator (MEASURED_PEAKS.json only carries HBM and bf16 tensor numbers). */
int v2gt_measure_fp64_peak(int32_t device, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* VICON2GT_B200_H */
