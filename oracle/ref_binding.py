"""ORACLE (test infrastructure, NOT product code).

ctypes binding of oracle/_ref/libv2gt_ref.so: the reference's OWN sources for the factor-level
part of the path (src/utils/quat_ops.h, rpy_ops.h, src/cpi/CpiV1.h, src/meas/Propagator.cpp,
Interpolator.cpp, src/gtsam/ImuFactorCPIv1.cpp, MeasBased_ViconPoseTimeoffsetFactor.cpp,
JPLNavState.cpp, JPLQuaternion.cpp, RotationXY.cpp) compiled where they lie under
/root/reference by oracle/Makefile.ref against the interface stand-ins of oracle/shim/.
Used by tests/test_ref_pin_cpu.py and tools/make_ref_golden.py only.

available() is False where neither the prebuilt library nor /root/reference exists (then the
committed tests/golden/ref_pin.npz carries the reference's outputs).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libv2gt_ref.so")
REFERENCE_ROOT = os.environ.get("V2GT_REFERENCE_ROOT", "/root/reference")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
PREINT_DIM = 288
_lib = None


def _d(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def can_build():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "cpi"))


def build():
    """Compile the reference sources (only where /root/reference exists)."""
    if not can_build():
        raise RuntimeError("reference sources not present at %s" % REFERENCE_ROOT)
    subprocess.check_call(["make", "-s", "-C", _HERE, "-f", "Makefile.ref", "REF=" + REFERENCE_ROOT])


def available():
    return os.path.exists(_LIB_PATH) or can_build()


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if can_build():
        build()  # make is a no-op when the library is current
    L = C.CDLL(_LIB_PATH)
    for n in ("ref_rot_2_quat", "ref_quat_2_Rot", "ref_exp_so3", "ref_log_so3", "ref_Jl_so3", "ref_rot2rpy"):
        getattr(L, n).argtypes = [c_double_p, c_double_p]
    L.ref_quat_multiply.argtypes = [c_double_p, c_double_p, c_double_p]
    L.ref_propagate.argtypes = [c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p,
                                c_double_p, c_double_p, c_double_p]
    L.ref_has_bounding_imu.argtypes = [C.c_int64, c_double_p, C.c_int64, c_double_p, c_int32_p]
    L.ref_interp_create.restype = C.c_void_p
    L.ref_interp_create.argtypes = [C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int]
    L.ref_interp_destroy.argtypes = [C.c_void_p]
    L.ref_interp_size.restype = C.c_int64
    L.ref_interp_size.argtypes = [C.c_void_p]
    L.ref_eval_interp.argtypes = [C.c_void_p, C.c_int64, c_double_p, C.c_int, c_int32_p, c_int32_p, c_int32_p, c_double_p,
                                  c_double_p, c_double_p, c_double_p]
    L.ref_eval_imu_factor.argtypes = [c_double_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      c_double_p, c_double_p]
    L.ref_eval_vicon_factor.argtypes = [C.c_void_p, c_int32_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_double_p]
    L.ref_retract_state.argtypes = [c_double_p, c_double_p, c_double_p]
    L.ref_local_state.argtypes = [c_double_p, c_double_p, c_double_p]
    L.ref_retract_quat.argtypes = [c_double_p, c_double_p, c_double_p]
    L.ref_retract_rotxy.argtypes = [c_double_p, c_double_p, c_double_p]
    L.ref_rotxy_rot.argtypes = [c_double_p, c_double_p]
    L.ref_solver_create.restype = C.c_void_p
    L.ref_solver_create.argtypes = [c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p,
                                    C.c_int64, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int, C.c_int64,
                                    c_double_p]
    L.ref_solver_destroy.argtypes = [C.c_void_p]
    L.ref_solver_build_and_solve.argtypes = [C.c_void_p]
    L.ref_solver_n_states.restype = C.c_int64
    L.ref_solver_n_states.argtypes = [C.c_void_p]
    L.ref_solver_get.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
    L.ref_stats_calculate.argtypes = [C.c_int64, c_double_p, c_double_p]
    L.ref_sim_run.restype = C.c_void_p
    L.ref_sim_run.argtypes = [C.c_char_p, c_double_p, c_double_p]
    L.ref_sim_destroy.argtypes = [C.c_void_p]
    for n in ("ref_sim_n_imu", "ref_sim_n_vicon"):
        getattr(L, n).restype = C.c_int64
        getattr(L, n).argtypes = [C.c_void_p]
    L.ref_sim_get.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p]
    L.ref_sim_state_in_vicon.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_int32_p]
    L.ref_solver_get_trace.restype = C.c_int64
    L.ref_solver_get_trace.argtypes = [C.c_int64, c_double_p]
    L.ref_solver_write_to_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    _lib = L
    return L


def _unary(name, x, shape):
    a, out = _c(x), np.zeros(shape)
    getattr(lib(), name)(_d(a), _d(out))
    return out


def rot_2_quat(R):
    return _unary("ref_rot_2_quat", R, 4)


def quat_2_Rot(q):
    return _unary("ref_quat_2_Rot", q, (3, 3))


def exp_so3(w):
    return _unary("ref_exp_so3", w, (3, 3))


def log_so3(R):
    return _unary("ref_log_so3", R, 3)


def Jl_so3(w):
    return _unary("ref_Jl_so3", w, (3, 3))


def rot2rpy(R):
    return _unary("ref_rot2rpy", R, 3)


def rotxy_rot(th):
    return _unary("ref_rotxy_rot", th, (3, 3))


def quat_multiply(q, p):
    a, b, out = _c(q), _c(p), np.zeros(4)
    lib().ref_quat_multiply(_d(a), _d(b), _d(out))
    return out


def propagate(sigmas, imu_t, imu_w, imu_a, time0, time1, bg_lin=(0, 0, 0), ba_lin=(0, 0, 0)):
    """Propagator::propagate + CpiV1 -> (preint row[288] without the information matrix, P[15,15]) or None."""
    s, t, w, a = _c(sigmas), _c(imu_t), _c(imu_w), _c(imu_a)
    bg, ba = _c(bg_lin), _c(ba_lin)
    out, cov = np.zeros(PREINT_DIM), np.zeros((15, 15))
    ok = lib().ref_propagate(_d(s), len(t), _d(t), _d(w), _d(a), float(time0), float(time1), _d(bg), _d(ba), _d(out), _d(cov))
    return (out, cov) if ok else None


def has_bounding_imu(imu_t, t_query):
    t, q = _c(imu_t), _c(t_query)
    out = np.zeros(len(q), dtype=np.int32)
    lib().ref_has_bounding_imu(len(t), _d(t), len(q), _d(q), _i(out))
    return out


class RefInterpolator:
    """The reference's Interpolator fed pose by pose (any order, duplicates allowed: std::set semantics)."""

    def __init__(self, vic_t, vic_q, vic_p, Rq, Rp):
        t, q, p, rq, rp = _c(vic_t), _c(vic_q), _c(vic_p), _c(Rq), _c(Rp)
        per_pose = 1 if rq.size == 9 * len(t) and len(t) != 1 else 0
        self._L = lib()
        self.h = self._L.ref_interp_create(len(t), _d(t), _d(q), _d(p), _d(rq), _d(rp), per_pose)

    def close(self):
        if self.h:
            self._L.ref_interp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return int(self._L.ref_interp_size(self.h))

    def eval(self, t_query, with_jacobian=True):
        t = _c(t_query)
        n = len(t)
        i0, i1, ok = (np.zeros(n, dtype=np.int32) for _ in range(3))
        q, p, R, H = np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 36)), np.zeros((n, 6))
        self._L.ref_eval_interp(self.h, n, _d(t), 1 if with_jacobian else 0, _i(i0), _i(i1), _i(ok), _d(q), _d(p), _d(R), _d(H))
        return dict(idx0=i0, idx1=i1, ok=ok, q=q, p=p, R=R, H_toff=H)

    def eval_vicon_factor(self, t, x16, globals10, estimate=(1, 1, 1)):
        """estimate = (toff, ori, pos) flags of GtsamConfig."""
        x, g = _c(x16), _c(globals10)
        fl = np.asarray(estimate, dtype=np.int32)
        e, H1, H2, H3, H4 = np.zeros(6), np.zeros((6, 15)), np.zeros((6, 3)), np.zeros((6, 3)), np.zeros(6)
        self._L.ref_eval_vicon_factor(self.h, _i(fl), float(t), _d(x), _d(g), _d(e), _d(H1), _d(H2), _d(H3), _d(H4))
        return dict(e=e, H1=H1, H2=H2, H3=H3, H4=H4)


def eval_imu_factor(preint_row, grav_mag, xi, xj, thetaxy):
    pr, a, b, th = _c(preint_row), _c(xi), _c(xj), _c(thetaxy)
    e, H1, H2, H3 = np.zeros(15), np.zeros((15, 15)), np.zeros((15, 15)), np.zeros((15, 2))
    lib().ref_eval_imu_factor(_d(pr), float(grav_mag), _d(a), _d(b), _d(th), _d(e), _d(H1), _d(H2), _d(H3))
    return dict(e=e, H1=H1, H2=H2, H3=H3)


def retract_state(x16, xi15):
    x, xi, out = _c(x16), _c(xi15), np.zeros(16)
    lib().ref_retract_state(_d(x), _d(xi), _d(out))
    return out


def local_state(x16, y16):
    x, y, out = _c(x16), _c(y16), np.zeros(15)
    lib().ref_local_state(_d(x), _d(y), _d(out))
    return out


def retract_quat(q4, d3):
    a, b, out = _c(q4), _c(d3), np.zeros(4)
    lib().ref_retract_quat(_d(a), _d(b), _d(out))
    return out


def retract_rotxy(th, d):
    a, b, out = _c(th), _c(d), np.zeros(2)
    lib().ref_retract_rotxy(_d(a), _d(b), _d(out))
    return out


class RefSolver:
    """The reference's ViconGraphSolver (src/solver/ViconGraphSolver.cpp, compiled where it lies) over one trial.

    Cleaning, state erasure, initial guess, graph construction, relinearisation loop and write_to_file are the
    reference's code; optimizer.optimize() is the dense GTSAM-LM restatement of oracle/shim/gtsam/mini_gtsam_lm.h.
    `params` is a v2gt_params (the ROS parameters of ViconGraphSolver.cpp:34-88); max_iterations=None keeps the
    reference's hard-coded 30, 0 hands back the initial values.
    """

    def __init__(self, params, ds, max_iterations=None):
        self._L = lib()
        p = params
        R_GtoV, R_BtoI, p_BinI = (_c(list(getattr(p, k))) for k in ("R_GtoV", "R_BtoI", "p_BinI"))
        par = _c([p.gravity_magnitude, p.toff_imu_to_vicon, p.estimate_toff_vicon_to_imu, p.estimate_ori_vicon_to_imu,
                  p.estimate_pos_vicon_to_imu, p.num_loop_relin, p.sigma_w, p.sigma_wb, p.sigma_a, p.sigma_ab,
                  -1 if max_iterations is None else max_iterations, 0, 0])
        a = [_c(x) for x in (ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t)]
        if getattr(ds, "vic_Rq", None) is not None:
            rq, rp, per = _c(ds.vic_Rq), _c(ds.vic_Rp), 1
        else:
            rc = _c(ds.vic_R_const)
            rq, rp, per = _c(rc[:9]), _c(rc[9:]), 0
        self.h = self._L.ref_solver_create(_d(R_GtoV), _d(R_BtoI), _d(p_BinI), _d(par), len(a[0]), _d(a[0]), _d(a[1]), _d(a[2]),
                                           len(a[3]), _d(a[3]), _d(a[4]), _d(a[5]), _d(rq), _d(rp), per, len(a[6]), _d(a[6]))

    def close(self):
        if self.h:
            self._L.ref_solver_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def build_and_solve(self):
        return self._L.ref_solver_build_and_solve(self.h)

    def result(self):
        n = int(self._L.ref_solver_n_states(self.h))
        t, x, g, lm = np.zeros(n), np.zeros((n, 16)), np.zeros(10), np.zeros(6)
        self._L.ref_solver_get(self.h, _d(t), _d(x), _d(g), _d(lm))
        return dict(times=t, states=x, globals10=g, iterations=int(lm[0]), tries=int(lm[1]), linearizations=int(lm[2]),
                    initial_error=lm[3], final_error=lm[4], lambda_=lm[5])

    def lm_trace(self):
        """[tries, 5]: lambda, error before, error after (inf if not evaluated), linear cost change, accepted."""
        n = int(self._L.ref_solver_get_trace(0, None))
        rows = np.zeros((n, 5))
        if n:
            self._L.ref_solver_get_trace(n, _d(rows))
        return rows

    def write_to_file(self, csv, info):
        return self._L.ref_solver_write_to_file(self.h, csv.encode(), info.encode())


def simulate(traj_path, seed=0, freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, sigmas=(1.6968e-04, 1.9393e-05, 2.0e-3, 3.0e-3),
             vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2), gravity=9.81, truth_times=None):
    """The reference's Simulator (src/sim/Simulator.cpp, BsplineSE3.cpp) driven like src/run_simulation.cpp:104-137."""
    L = lib()
    par = _c([seed, freq_imu, freq_cam, freq_vicon, *sigmas, gravity, 0, 0, 0, 0])
    vs = _c(vicon_sigmas)
    h = L.ref_sim_run(traj_path.encode(), _d(par), _d(vs))
    try:
        M, V = int(L.ref_sim_n_imu(h)), int(L.ref_sim_n_vicon(h))
        imu, vic, truth = np.zeros((M, 7)), np.zeros((V, 8)), np.zeros(22)
        L.ref_sim_get(h, _d(imu), _d(vic), _d(truth))
        out = dict(imu_t=imu[:, 0].copy(), imu_w=imu[:, 1:4].copy(), imu_a=imu[:, 4:7].copy(), vic_t=vic[:, 0].copy(),
                   vic_q=vic[:, 1:5].copy(), vic_p=vic[:, 5:8].copy(), R_BtoI=truth[:9].reshape(3, 3), p_BinI=truth[9:12].copy(),
                   viconimu_dt=float(truth[12]), R_GtoV=truth[13:].reshape(3, 3))
        if truth_times is not None:
            t = _c(truth_times)
            st, ok = np.zeros((len(t), 17)), np.zeros(len(t), dtype=np.int32)
            L.ref_sim_state_in_vicon(h, len(t), _d(t), _d(st), _i(ok))
            out["states"], out["states_ok"] = st, ok
    finally:
        L.ref_sim_destroy(h)
    return out


def stats_calculate(values):
    """Stats::calculate (src/utils/stats.h:64-112) -> rmse, mean, median, min, max, std, ninetynine."""
    v, out = _c(values), np.zeros(7)
    lib().ref_stats_calculate(len(v), _d(v), _d(out))
    return out
