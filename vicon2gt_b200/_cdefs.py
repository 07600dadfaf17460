"""ctypes mirrors of the C structs in include/vicon2gt_b200.h and host/v2gt_sim.h."""
import ctypes as C

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)


class Params(C.Structure):
    """v2gt_params (ViconGraphSolverParams): see include/vicon2gt_b200.h for the field docs."""

    _fields_ = [
        ("R_GtoV", C.c_double * 9),
        ("gravity_magnitude", C.c_double),
        ("R_BtoI", C.c_double * 9),
        ("p_BinI", C.c_double * 3),
        ("toff_imu_to_vicon", C.c_double),
        ("estimate_toff_vicon_to_imu", C.c_int32),
        ("estimate_ori_vicon_to_imu", C.c_int32),
        ("estimate_pos_vicon_to_imu", C.c_int32),
        ("num_loop_relin", C.c_int32),
        ("sigma_w", C.c_double),
        ("sigma_wb", C.c_double),
        ("sigma_a", C.c_double),
        ("sigma_ab", C.c_double),
        ("lm_max_iterations", C.c_int32),
        ("lm_max_tries", C.c_int32),
        ("lm_lambda_initial", C.c_double),
        ("lm_lambda_factor", C.c_double),
        ("lm_lambda_upper_bound", C.c_double),
        ("lm_lambda_lower_bound", C.c_double),
        ("lm_min_model_fidelity", C.c_double),
        ("lm_relative_error_tol", C.c_double),
        ("lm_absolute_error_tol", C.c_double),
        ("huber_k", C.c_double),
        ("time_offset_window", C.c_double),
        ("interp_gap_init", C.c_double),
        ("interp_gap_factor", C.c_double),
    ]


class TrialInfo(C.Structure):
    """v2gt_trial_info."""

    _fields_ = [
        ("status", C.c_int32),
        ("n_states", C.c_int32),
        ("removed_no_imu", C.c_int32),
        ("removed_before", C.c_int32),
        ("removed_after", C.c_int32),
        ("removed_no_vicon", C.c_int32),
        ("iterations", C.c_int32),
        ("tries", C.c_int32),
        ("linearizations", C.c_int32),
        ("lm_state", C.c_int32),
        ("initial_error", C.c_double),
        ("final_error", C.c_double),
        ("lambda_", C.c_double),
    ]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


class SimParams(C.Structure):
    """v2gt_sim_params (SimulatorParams of the reference)."""

    _fields_ = [
        ("sigma_w", C.c_double),
        ("sigma_wb", C.c_double),
        ("sigma_a", C.c_double),
        ("sigma_ab", C.c_double),
        ("sigma_vicon_pose", C.c_double * 6),
        ("gravity_magnitude", C.c_double),
        ("sim_freq_imu", C.c_double),
        ("sim_freq_cam", C.c_double),
        ("sim_freq_vicon", C.c_double),
        ("seed", C.c_int32),
        ("state_freq", C.c_int32),
        ("max_imu", C.c_int64),
        ("use_fixed_truth", C.c_int32),
        ("_pad", C.c_int32),
        ("fixed_R_BtoI", C.c_double * 9),
        ("fixed_p_BinI", C.c_double * 3),
        ("fixed_R_GtoV", C.c_double * 9),
        ("fixed_viconimu_dt", C.c_double),
    ]


STATUS_NAMES = {
    0: "OK",
    1: "INVALID_ARG",
    2: "CUDA",
    3: "NO_TIMESTAMPS",
    4: "NO_VICON",
    5: "ALL_OUT_OF_RANGE",
    6: "INVALID_PREINT",
    7: "NAN_COVARIANCE",
    8: "NOT_BUILT",
    9: "NO_DEVICE",
    10: "NCCL",
}

PREINT_DIM = 288


def dptr(a):
    """numpy float64 C-contiguous array -> double* (None -> NULL)."""
    if a is None:
        return None
    return a.ctypes.data_as(c_double_p)


def i32ptr(a):
    if a is None:
        return None
    return a.ctypes.data_as(c_int32_p)
