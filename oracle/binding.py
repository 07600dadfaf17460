"""ORACLE (test infrastructure, NOT product code) -- factor level pinned (oracle/_ref), optimiser level unpinned; see DESIGN.md section 2.

ctypes binding of oracle/liboracle.so (the C++ CPU restatement of the reference's
build-and-solve path).  Import only from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_int64_p = C.POINTER(C.c_int64)
PREINT_DIM = 288


def _d(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _i(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_LIB_PATH):
        build()
    L = C.CDLL(_LIB_PATH)
    L.orc_create.restype = C.c_void_p
    L.orc_create.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p,
                             c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p]
    L.orc_destroy.argtypes = [C.c_void_p]
    L.orc_params_default.argtypes = [C.c_void_p]
    L.orc_set_option.argtypes = [C.c_void_p, C.c_char_p, C.c_double]
    for n in ("orc_build_and_solve", "orc_clean", "orc_optimize"):
        getattr(L, n).argtypes = [C.c_void_p]
    L.orc_build.argtypes = [C.c_void_p, C.c_int]
    L.orc_get_times.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_get_info.argtypes = [C.c_void_p, C.c_void_p]
    L.orc_n_states.restype = C.c_int64
    L.orc_n_states.argtypes = [C.c_void_p]
    L.orc_n_vicon.restype = C.c_int64
    L.orc_n_vicon.argtypes = [C.c_void_p]
    L.orc_get_states.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_get_times_only.argtypes = [C.c_void_p, c_double_p]
    L.orc_get_calibration.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
    L.orc_set_estimate.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p]
    L.orc_eval_interp.argtypes = [C.c_void_p, C.c_int64, c_double_p, C.c_double, c_int32_p, c_int32_p, c_int32_p, c_double_p,
                                  c_double_p, c_double_p, c_double_p]
    L.orc_get_preint.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_propagate.argtypes = [c_double_p, C.c_int64, c_double_p, c_double_p, c_double_p, C.c_double, C.c_double, c_double_p,
                                c_double_p, C.c_int, c_double_p, c_double_p, c_int32_p]
    L.orc_eval_linearize.argtypes = [C.c_void_p] + [c_double_p] * 8
    L.orc_eval_solve.argtypes = [C.c_void_p, C.c_double, c_double_p, c_int32_p]
    L.orc_eval_cost.argtypes = [C.c_void_p, c_double_p]
    L.orc_eval_cost_at.argtypes = [C.c_void_p, c_double_p, c_double_p]
    L.orc_apply_delta.argtypes = [C.c_void_p, c_double_p]
    L.orc_get_lm_trace.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_int64_p]
    L.orc_write_to_file.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
    L.orc_eval_imu_factor.argtypes = [c_double_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                      c_double_p, c_double_p]
    L.orc_eval_vicon_factor.argtypes = [C.c_void_p, C.c_double, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p,
                                        c_double_p, c_double_p, c_int32_p]
    L.orc_retract_state.argtypes = [c_double_p, c_double_p, c_double_p]
    L.orc_retract_rotxy.argtypes = [c_double_p, c_double_p, c_double_p]
    for n in ("orc_rot_2_quat", "orc_quat_2_Rot", "orc_exp_so3", "orc_log_so3", "orc_Jl_so3", "orc_rot2rpy"):
        getattr(L, n).argtypes = [c_double_p, c_double_p]
    L.orc_quat_multiply.argtypes = [c_double_p, c_double_p, c_double_p]
    _lib = L
    return L


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def default_params(**kw):
    """v2gt_params with the reference defaults, filled by the oracle's own copy of the defaults."""
    from vicon2gt_b200._cdefs import Params

    p = Params()
    lib().orc_params_default(C.byref(p))
    for k, v in kw.items():
        if k in ("R_GtoV", "R_BtoI", "p_BinI"):
            arr = np.asarray(v, dtype=np.float64).reshape(-1)
            for i, x in enumerate(arr):
                getattr(p, k)[i] = float(x)
        else:
            setattr(p, k, v)
    return p


class OracleSolver:
    """CPU restatement of ViconGraphSolver over one trial's inputs."""

    def __init__(self, params, imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, state_t, vic_Rq=None, vic_Rp=None, vic_R_const=None):
        L = lib()
        self._L = L
        a = [_c(x) for x in (imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, vic_Rq, vic_Rp, vic_R_const, state_t)]
        self.params = params
        self.h = L.orc_create(C.byref(params), len(a[0]), _d(a[0]), _d(a[1]), _d(a[2]), len(a[3]), _d(a[3]), _d(a[4]), _d(a[5]),
                              _d(a[6]), _d(a[7]), _d(a[8]), len(a[9]), _d(a[9]))

    @classmethod
    def from_dataset(cls, params, ds):
        if getattr(ds, "vic_Rq", None) is not None and getattr(ds, "vic_Rp", None) is not None:
            return cls(params, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t, vic_Rq=ds.vic_Rq, vic_Rp=ds.vic_Rp)
        return cls(params, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t, vic_R_const=ds.vic_R_const)

    def close(self):
        if self.h:
            self._L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_option(self, name, value):
        return self._L.orc_set_option(self.h, name.encode(), float(value))

    def build_and_solve(self):
        return self._L.orc_build_and_solve(self.h)

    def clean(self):
        return self._L.orc_clean(self.h)

    def build(self, init_states=True):
        return self._L.orc_build(self.h, 1 if init_states else 0)

    def optimize(self):
        return self._L.orc_optimize(self.h)

    def phase_times(self):
        b, o = C.c_double(), C.c_double()
        self._L.orc_get_times(self.h, C.byref(b), C.byref(o))
        return b.value, o.value

    def info(self, info_cls):
        inf = info_cls()
        self._L.orc_get_info(self.h, C.byref(inf))
        return inf

    @property
    def n_states(self):
        return int(self._L.orc_n_states(self.h))

    @property
    def n_vicon(self):
        return int(self._L.orc_n_vicon(self.h))

    def times(self):
        t = np.zeros(self.n_states)
        self._L.orc_get_times_only(self.h, _d(t))
        return t

    def states(self):
        n = self.n_states
        t, x = np.zeros(n), np.zeros((n, 16))
        self._L.orc_get_states(self.h, _d(t), _d(x))
        return t, x

    def calibration(self):
        q, p, th = np.zeros(4), np.zeros(3), np.zeros(2)
        toff = C.c_double()
        self._L.orc_get_calibration(self.h, _d(q), _d(p), C.byref(toff), _d(th))
        return q, p, toff.value, th

    def globals10(self):
        q, p, toff, th = self.calibration()
        return np.concatenate([q, p, [toff], th])

    def set_estimate(self, states=None, globals10=None):
        s, g = _c(states), _c(globals10)
        return self._L.orc_set_estimate(self.h, self.n_states, _d(s), _d(g))

    def eval_interp(self, t_query, gap_thresh):
        t = _c(t_query)
        n = len(t)
        i0, i1, ok = (np.zeros(n, dtype=np.int32) for _ in range(3))
        q, p, R, H = np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 36)), np.zeros((n, 6))
        self._L.orc_eval_interp(self.h, n, _d(t), float(gap_thresh), _i(i0), _i(i1), _i(ok), _d(q), _d(p), _d(R), _d(H))
        return dict(idx0=i0, idx1=i1, ok=ok, q=q, p=p, R=R, H_toff=H)

    def preint(self):
        n = max(self.n_states - 1, 0)
        out, cov = np.zeros((n, PREINT_DIM)), np.zeros((n, 225))
        self._L.orc_get_preint(self.h, _d(out), _d(cov))
        return out, cov

    def linearize(self):
        n = self.n_states
        D, O, B, g = np.zeros((n, 225)), np.zeros((n, 225)), np.zeros((n, 135)), np.zeros((n, 15))
        G, gG = np.zeros(81), np.zeros(9)
        cost, lin0 = C.c_double(), C.c_double()
        self._L.orc_eval_linearize(self.h, _d(D), _d(O), _d(B), _d(g), _d(G), _d(gG), C.byref(cost), C.byref(lin0))
        return dict(D=D, O=O, B=B, g=g, Gamma=G, g_gamma=gG, cost=cost.value, lin_error0=lin0.value)

    def solve(self, lam):
        n = self.n_states
        delta = np.zeros(15 * n + 9)
        ok = C.c_int32()
        self._L.orc_eval_solve(self.h, float(lam), _d(delta), C.byref(ok))
        return delta, bool(ok.value)

    def cost(self):
        c = C.c_double()
        self._L.orc_eval_cost(self.h, C.byref(c))
        return c.value

    def cost_at(self, delta):
        c = C.c_double()
        d = _c(delta)
        self._L.orc_eval_cost_at(self.h, _d(d), C.byref(c))
        return c.value

    def apply_delta(self, delta):
        d = _c(delta)
        return self._L.orc_apply_delta(self.h, _d(d))

    def lm_trace(self):
        n = C.c_int64()
        self._L.orc_get_lm_trace(self.h, 0, None, C.byref(n))
        rows = np.zeros((n.value, 5))
        if n.value:
            self._L.orc_get_lm_trace(self.h, n.value, _d(rows), C.byref(n))
        return rows

    def write_to_file(self, csv, info):
        return self._L.orc_write_to_file(self.h, csv.encode(), info.encode())

    def eval_vicon_factor(self, t, x16, globals10):
        x, g = _c(x16), _c(globals10)
        e, H1, H2, H3, H4 = np.zeros(6), np.zeros((6, 15)), np.zeros((6, 3)), np.zeros((6, 3)), np.zeros(6)
        hv = C.c_int32()
        self._L.orc_eval_vicon_factor(self.h, float(t), _d(x), _d(g), _d(e), _d(H1), _d(H2), _d(H3), _d(H4), C.byref(hv))
        return dict(e=e, H1=H1, H2=H2, H3=H3, H4=H4, has_vicon=bool(hv.value))


def propagate(sigmas, imu_t, imu_w, imu_a, time0, time1, bg_lin=(0, 0, 0), ba_lin=(0, 0, 0), faithful=False):
    """Propagator::propagate over a stand-alone IMU store -> (preint row[288], P[15,15], n_steps) or None."""
    L = lib()
    s, t, w, a = _c(sigmas), _c(imu_t), _c(imu_w), _c(imu_a)
    bg, ba = _c(bg_lin), _c(ba_lin)
    out, cov = np.zeros(PREINT_DIM), np.zeros((15, 15))
    ns = C.c_int32()
    st = L.orc_propagate(_d(s), len(t), _d(t), _d(w), _d(a), float(time0), float(time1), _d(bg), _d(ba), 1 if faithful else 0,
                         _d(out), _d(cov), C.byref(ns))
    if st != 0:
        return None
    return out, cov, ns.value


def eval_imu_factor(preint_row, grav_mag, xi, xj, thetaxy):
    L = lib()
    pr, a, b, th = _c(preint_row), _c(xi), _c(xj), _c(thetaxy)
    e, H1, H2, H3 = np.zeros(15), np.zeros((15, 15)), np.zeros((15, 15)), np.zeros((15, 2))
    L.orc_eval_imu_factor(_d(pr), float(grav_mag), _d(a), _d(b), _d(th), _d(e), _d(H1), _d(H2), _d(H3))
    return dict(e=e, H1=H1, H2=H2, H3=H3)


def retract_state(x16, xi15):
    L = lib()
    x, xi, out = _c(x16), _c(xi15), np.zeros(16)
    L.orc_retract_state(_d(x), _d(xi), _d(out))
    return out


def retract_rotxy(th, d):
    L = lib()
    a, b, out = _c(th), _c(d), np.zeros(2)
    L.orc_retract_rotxy(_d(a), _d(b), _d(out))
    return out


def _unary(name, x, shape):
    L = lib()
    a = _c(x)
    out = np.zeros(shape)
    getattr(L, name)(_d(a), _d(out))
    return out


def rot_2_quat(R):
    return _unary("orc_rot_2_quat", R, 4)


def quat_2_Rot(q):
    return _unary("orc_quat_2_Rot", q, (3, 3))


def exp_so3(w):
    return _unary("orc_exp_so3", w, (3, 3))


def log_so3(R):
    return _unary("orc_log_so3", R, 3)


def Jl_so3(w):
    return _unary("orc_Jl_so3", w, (3, 3))


def rot2rpy(R):
    return _unary("orc_rot2rpy", R, 3)


def quat_multiply(q, p):
    L = lib()
    a, b, out = _c(q), _c(p), np.zeros(4)
    L.orc_quat_multiply(_d(a), _d(b), _d(out))
    return out
