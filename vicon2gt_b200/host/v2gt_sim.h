/*
 * vicon2gt_b200 host library: synthetic input generator (C ABI).
 * Mirrors the reference's SimulatorParams (src/sim/SimulatorParams.h:36-132) and the
 * measurement loop of src/run_simulation.cpp:108-158.  CPU only, never on the solve path.
 */
#ifndef V2GT_SIM_H
#define V2GT_SIM_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct v2gt_sim_params {
  double sigma_w, sigma_wb, sigma_a, sigma_ab; /* IMU noise (SimulatorParams.h:41-47)                  */
  double sigma_vicon_pose[6];                  /* vicon noise: ori(3) rad, pos(3) m (SimulatorParams.h:53) */
  double gravity_magnitude;                    /* 9.81                                                 */
  double sim_freq_imu, sim_freq_cam, sim_freq_vicon; /* 400 / 10 / 100 Hz (sim.launch: 400 / 20 / 100) */
  int32_t seed;                                /* std::mt19937 seed for truth, IMU and vicon streams   */
  int32_t state_freq;                          /* state grid Hz (run_simulation.cpp:54); <=0: states at the trajectory rows' timestamps */
  int64_t max_imu;                             /* stop after this many IMU samples (0 = whole trajectory) */
  int32_t use_fixed_truth;                     /* 1: use the fixed_* truth below instead of the per-seed random draw */
  int32_t _pad;
  double fixed_R_BtoI[9], fixed_p_BinI[3], fixed_R_GtoV[9], fixed_viconimu_dt;
} v2gt_sim_params;

typedef struct v2gt_sim v2gt_sim;

void v2gt_sim_params_default(v2gt_sim_params *p);
v2gt_sim *v2gt_sim_create(const v2gt_sim_params *p);
void v2gt_sim_destroy(v2gt_sim *s);
const char *v2gt_sim_error(v2gt_sim *s);
/* trajectory rows "t x y z qx qy qz qw" (JPL q_GtoI), from a text file, from memory, or synthesised */
int v2gt_sim_load_trajectory_file(v2gt_sim *s, const char *path);
int v2gt_sim_set_trajectory(v2gt_sim *s, int64_t n_rows, const double *rows8);
int v2gt_sim_make_trajectory(double t_start, int64_t n_rows, double dt, double extent_m, double *rows8);
/* run the whole simulation (IMU / vicon streams + state time grid) */
int v2gt_sim_run(v2gt_sim *s);
int64_t v2gt_sim_n_imu(v2gt_sim *s);
int64_t v2gt_sim_n_vicon(v2gt_sim *s);
int64_t v2gt_sim_n_states(v2gt_sim *s);
int v2gt_sim_get(v2gt_sim *s, double *imu_t, double *imu_w, double *imu_a, double *vic_t, double *vic_q, double *vic_p,
                 double *state_t);
int v2gt_sim_get_truth(v2gt_sim *s, double *R_BtoI, double *p_BinI, double *viconimu_dt, double *R_GtoV);
int v2gt_sim_get_state_in_vicon(v2gt_sim *s, int64_t n, const double *times, double *out17, int32_t *ok);

#ifdef __cplusplus
}
#endif
#endif
