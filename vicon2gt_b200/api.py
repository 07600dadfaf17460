"""Python mirror of the reference's solver interface over the C ABI (ctypes, no torch types).

`ViconGraphSolver` follows src/solver/ViconGraphSolver.h:61-168 of the reference: it is built from a
parameter struct, a `Propagator` (feed_imu) and an `Interpolator` (feed_pose) store and the state
timestamps, and offers build_and_solve / write_to_file / get_imu_poses / get_calibration.
`BatchSolver` is the batch form (many independent trials per GPU) the Monte-Carlo configs use.
Everything numeric happens in vicon2gt_b200/libvicon2gt_b200.so (CUDA, sm_100a); there is no CPU path.
"""
import ctypes as C
import os

import numpy as np

from ._cdefs import PREINT_DIM, STATUS_NAMES, Params, SimParams, TrialInfo, c_double_p, c_int32_p, c_int64_p, dptr, i32ptr
from .build import lib_path

_lib = None


class V2gtError(RuntimeError):
    def __init__(self, code, where=""):
        self.code = code
        super().__init__(f"vicon2gt_b200: {where} failed with status {code} ({STATUS_NAMES.get(code, '?')})")


def load_library():
    """Load the CUDA library; raises if it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path("cuda")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} is missing: build it with `python -m vicon2gt_b200.build` (nvcc, sm_100a). "
                           "vicon2gt_b200 has no CPU implementation.")
    L = C.CDLL(path)
    vp = C.c_void_p
    L.v2gt_version.restype = C.c_char_p
    L.v2gt_params_default.argtypes = [C.POINTER(Params)]
    L.v2gt_device_count.argtypes = [c_int32_p]
    L.v2gt_batch_create.argtypes = [C.c_int32, C.c_int32, C.POINTER(vp)]
    L.v2gt_batch_destroy.argtypes = [vp]
    L.v2gt_batch_set_option.argtypes = [vp, C.c_char_p, C.c_double]
    L.v2gt_batch_set_trial.argtypes = [vp, C.c_int32, C.POINTER(Params), C.c_int64, c_double_p, c_double_p, c_double_p, C.c_int64,
                                       c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, c_double_p, C.c_int64, c_double_p]
    L.v2gt_batch_set_trial_view.argtypes = L.v2gt_batch_set_trial.argtypes
    for n in ("v2gt_batch_upload", "v2gt_batch_optimize", "v2gt_batch_build_and_solve", "v2gt_batch_sync"):
        getattr(L, n).argtypes = [vp]
    L.v2gt_batch_build.argtypes = [vp, C.c_int32]
    L.v2gt_batch_get_info.argtypes = [vp, C.c_int32, C.POINTER(TrialInfo)]
    L.v2gt_batch_get_states.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_double_p, c_int64_p]
    L.v2gt_batch_get_all_states.argtypes = [vp, C.c_int64, c_int64_p, c_double_p, c_double_p, c_double_p]
    L.v2gt_batch_get_calibration.argtypes = [vp, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]
    L.v2gt_batch_write_to_file.argtypes = [vp, C.c_int32, C.c_char_p, C.c_char_p]
    L.v2gt_batch_write_all.argtypes = [vp, C.c_char_p, C.c_int32, c_int32_p]
    L.v2gt_batch_set_truth.argtypes = [vp, C.c_int32, C.c_int64, c_double_p]
    L.v2gt_batch_error_stats.argtypes = [vp, c_double_p]
    L.v2gt_eval_interp.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, C.c_double, c_int32_p, c_int32_p, c_int32_p, c_double_p,
                                   c_double_p, c_double_p, c_double_p]
    L.v2gt_batch_get_preint.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_int64_p]
    L.v2gt_batch_get_preint_cov.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_int64_p]
    L.v2gt_batch_set_estimate.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_double_p]
    L.v2gt_eval_linearize.argtypes = [vp, C.c_int32, C.c_int64] + [c_double_p] * 8
    L.v2gt_eval_solve.argtypes = [vp, C.c_int32, C.c_double, C.c_int64, c_double_p, c_int32_p]
    L.v2gt_eval_cost.argtypes = [vp, C.c_int32, c_double_p]
    L.v2gt_batch_get_lm_trace.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_int64_p]
    L.v2gt_batch_get_timing.argtypes = [vp, c_double_p, c_int64_p]
    L.v2gt_measure_fp64_peak.argtypes = [C.c_int32, c_double_p]
    # batched simulator on the device (csrc/k_sim.cu); the params struct is the host simulator's (host/v2gt_sim.h)
    L.v2gt_simbatch_create.argtypes = [C.c_int32, C.c_int32, C.c_int64, c_double_p, C.POINTER(vp)]
    L.v2gt_simbatch_destroy.argtypes = [vp]
    L.v2gt_simbatch_set_trial.argtypes = [vp, C.c_int32, C.POINTER(SimParams)]
    L.v2gt_simbatch_plan.argtypes = [vp]
    L.v2gt_simbatch_run.argtypes = [vp]
    L.v2gt_simbatch_sizes.argtypes = [vp, C.c_int32, c_int64_p, c_int64_p, c_int64_p]
    L.v2gt_simbatch_get.argtypes = [vp, C.c_int32] + [c_double_p] * 7
    L.v2gt_simbatch_get_truth.argtypes = [vp, C.c_int32, c_double_p, c_double_p, c_double_p, c_double_p]
    L.v2gt_simbatch_get_state_in_vicon.argtypes = [vp, C.c_int32, C.c_int64, c_double_p, c_double_p, c_int32_p]
    L.v2gt_simbatch_get_timing.argtypes = [vp, c_double_p]
    L.v2gt_batch_set_trial_sim.argtypes = [vp, C.c_int32, C.POINTER(Params), vp, C.c_int32, c_double_p]
    _lib = L
    return L


def default_params(**kw):
    """ViconGraphSolverParams with the reference defaults (ViconGraphSolver.cpp:34-75, sim.launch)."""
    p = Params()
    load_library().v2gt_params_default(C.byref(p))
    for k, v in kw.items():
        if k in ("R_GtoV", "R_BtoI", "p_BinI"):
            arr = np.asarray(v, dtype=np.float64).reshape(-1)
            for i, x in enumerate(arr):
                getattr(p, k)[i] = float(x)
        else:
            setattr(p, k, v)
    return p


def device_count():
    n = C.c_int32(0)
    load_library().v2gt_device_count(C.byref(n))
    return n.value


def measure_fp64_peak(device=0):
    """Measured FP64 FMA peak (TFLOP/s) of a device."""
    v = C.c_double()
    _check(load_library().v2gt_measure_fp64_peak(int(device), C.byref(v)), "v2gt_measure_fp64_peak")
    return v.value


def _c(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def _check(code, where):
    if code != 0:
        raise V2gtError(code, where)


class BatchSolver:
    """A batch of independent build-and-solve problems resident on one GPU."""

    def __init__(self, n_trials, device=0, **options):
        self._L = load_library()
        self.n_trials = int(n_trials)
        self.h = C.c_void_p()
        _check(self._L.v2gt_batch_create(int(device), self.n_trials, C.byref(self.h)), "v2gt_batch_create")
        for k, v in options.items():
            self.set_option(k, v)

    def close(self):
        if getattr(self, "h", None):
            self._L.v2gt_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_option(self, name, value):
        _check(self._L.v2gt_batch_set_option(self.h, name.encode(), float(value)), f"set_option({name})")

    def set_trial(self, trial, params, imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, state_t, vic_Rq=None, vic_Rp=None, vic_R_const=None,
                  borrow=False):
        """Stage one trial.  borrow=True hands the library views of the arrays instead of copies (v2gt_batch_set_trial_view);
        this object keeps them alive until the next upload."""
        a = [_c(x) for x in (imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, vic_Rq, vic_Rp, vic_R_const, state_t)]
        fn = self._L.v2gt_batch_set_trial_view if borrow else self._L.v2gt_batch_set_trial
        _check(fn(self.h, int(trial), C.byref(params), len(a[0]), dptr(a[0]), dptr(a[1]), dptr(a[2]),
                  len(a[3]), dptr(a[3]), dptr(a[4]), dptr(a[5]), dptr(a[6]), dptr(a[7]), dptr(a[8]),
                  len(a[9]), dptr(a[9])), "v2gt_batch_set_trial")
        if borrow:
            if not hasattr(self, "_views"):
                self._views = {}
            self._views[int(trial)] = a

    def set_trial_dataset(self, trial, params, ds, borrow=False):
        per_pose = getattr(ds, "vic_Rq", None) is not None and getattr(ds, "vic_Rp", None) is not None
        self.set_trial(trial, params, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t,
                       vic_Rq=ds.vic_Rq if per_pose else None, vic_Rp=ds.vic_Rp if per_pose else None,
                       vic_R_const=None if per_pose else ds.vic_R_const, borrow=borrow)

    def set_trial_sim(self, trial, params, simbatch, sim_trial=None):
        """Stage trial `sim_trial` of a device-simulated batch (sim.SimBatch, same GPU): the measurement rows are copied
        device to device at upload; `simbatch` must stay alive until then."""
        k = int(trial if sim_trial is None else sim_trial)
        R = _c(simbatch.vic_R_const(k))
        _check(self._L.v2gt_batch_set_trial_sim(self.h, int(trial), C.byref(params), simbatch.h, k, dptr(R)), "v2gt_batch_set_trial_sim")
        if not hasattr(self, "_views"):
            self._views = {}
        self._views[int(trial)] = [simbatch, R]

    def upload(self):
        _check(self._L.v2gt_batch_upload(self.h), "v2gt_batch_upload")

    def build(self, init_states=True):
        _check(self._L.v2gt_batch_build(self.h, 1 if init_states else 0), "v2gt_batch_build")

    def optimize(self):
        _check(self._L.v2gt_batch_optimize(self.h), "v2gt_batch_optimize")

    def build_and_solve(self):
        _check(self._L.v2gt_batch_build_and_solve(self.h), "v2gt_batch_build_and_solve")

    def sync(self):
        _check(self._L.v2gt_batch_sync(self.h), "v2gt_batch_sync")

    def info(self, trial=0):
        inf = TrialInfo()
        _check(self._L.v2gt_batch_get_info(self.h, int(trial), C.byref(inf)), "v2gt_batch_get_info")
        return inf

    def states(self, trial=0):
        n = C.c_int64()
        _check(self._L.v2gt_batch_get_states(self.h, int(trial), 0, None, None, C.byref(n)), "v2gt_batch_get_states")
        t, x = np.empty(n.value), np.empty((n.value, 16))
        _check(self._L.v2gt_batch_get_states(self.h, int(trial), n.value, dptr(t), dptr(x), C.byref(n)), "v2gt_batch_get_states")
        return t, x

    def all_states(self):
        """Every trial's result in one read-back: (offsets [T+1], times [n], states [n, 16], globals [T, 10]); trial t owns
        rows offsets[t]:offsets[t+1]."""
        off = np.zeros(self.n_trials + 1, dtype=np.int64)
        optr = off.ctypes.data_as(c_int64_p)
        _check(self._L.v2gt_batch_get_all_states(self.h, 0, optr, None, None, None), "v2gt_batch_get_all_states")
        n = int(off[-1])
        t, x, g = np.empty(n), np.empty((n, 16)), np.empty((self.n_trials, 10))
        _check(self._L.v2gt_batch_get_all_states(self.h, n, optr, dptr(t), dptr(x), dptr(g)), "v2gt_batch_get_all_states")
        return off, t, x, g

    def calibration(self, trial=0):
        q, p, th = np.zeros(4), np.zeros(3), np.zeros(2)
        toff = C.c_double()
        _check(self._L.v2gt_batch_get_calibration(self.h, int(trial), dptr(q), dptr(p), C.byref(toff), dptr(th)),
               "v2gt_batch_get_calibration")
        return q, p, toff.value, th

    def globals10(self, trial=0):
        q, p, toff, th = self.calibration(trial)
        return np.concatenate([q, p, [toff], th])

    def write_to_file(self, trial, csv_path, info_path):
        _check(self._L.v2gt_batch_write_to_file(self.h, int(trial), csv_path.encode(), info_path.encode()),
               "v2gt_batch_write_to_file")

    def write_all(self, directory, with_txt=True):
        """CSV + info (+ pose text) files of every solved trial; returns the number of trials written."""
        n = C.c_int32(0)
        _check(self._L.v2gt_batch_write_all(self.h, str(directory).encode(), 1 if with_txt else 0, C.byref(n)), "v2gt_batch_write_all")
        return n.value

    def set_truth(self, trial, truth10):
        t = _c(truth10)
        _check(self._L.v2gt_batch_set_truth(self.h, int(trial), t.shape[0], dptr(t)), "v2gt_batch_set_truth")

    def set_truth_from_dataset(self, trial, ds):
        """Truth of a simulated Dataset at the trial's surviving state times (Simulator::get_state_in_vicon rows
        [t, q_VtoI 4, p_IinV 3, v_IinV 3, bg 3, ba 3], matched by the exact time stamp)."""
        times, _ = self.states(trial)
        idx = np.searchsorted(ds.state_t, times)
        if np.any(idx >= len(ds.state_t)) or not np.array_equal(ds.state_t[idx], times):
            raise ValueError("surviving state times are not a subset of the dataset's state times")
        self.set_truth(trial, ds.truth["states"][idx, 1:11])

    def set_truth_from_sim(self, trial, simbatch, sim_trial=None):
        """The same from a device-simulated batch: Simulator::get_state_in_vicon evaluated on the GPU at the surviving times."""
        times, _ = self.states(trial)
        st, _ = simbatch.state_in_vicon(int(trial if sim_trial is None else sim_trial), times)
        self.set_truth(trial, st[:, 1:11])

    def error_stats(self):
        """[n_trials, 3, 8]: (ori deg, pos m, vel m/s) x (rmse, mean, median, min, max, std, mean + 2.326 std, count)."""
        out = np.zeros((self.n_trials, 3, 8))
        _check(self._L.v2gt_batch_error_stats(self.h, dptr(out)), "v2gt_batch_error_stats")
        return out

    # ---- evaluation-only entry points (parity tests)
    def eval_interp(self, trial, t_query, gap_thresh):
        t = _c(t_query)
        n = len(t)
        i0, i1, ok = (np.zeros(n, dtype=np.int32) for _ in range(3))
        q, p, R, H = np.zeros((n, 4)), np.zeros((n, 3)), np.zeros((n, 36)), np.zeros((n, 6))
        _check(self._L.v2gt_eval_interp(self.h, int(trial), n, dptr(t), float(gap_thresh), i32ptr(i0), i32ptr(i1), i32ptr(ok),
                                        dptr(q), dptr(p), dptr(R), dptr(H)), "v2gt_eval_interp")
        return dict(idx0=i0, idx1=i1, ok=ok, q=q, p=p, R=R, H_toff=H)

    def preint(self, trial=0, cov=False):
        n = C.c_int64()
        _check(self._L.v2gt_batch_get_preint(self.h, int(trial), 0, None, C.byref(n)), "v2gt_batch_get_preint")
        out = np.zeros((n.value, PREINT_DIM))
        _check(self._L.v2gt_batch_get_preint(self.h, int(trial), n.value, dptr(out), C.byref(n)), "v2gt_batch_get_preint")
        if not cov:
            return out
        P = np.zeros((n.value, 225))
        _check(self._L.v2gt_batch_get_preint_cov(self.h, int(trial), n.value, dptr(P), C.byref(n)), "v2gt_batch_get_preint_cov")
        return out, P

    def set_estimate(self, trial, states=None, globals10=None):
        s, g = _c(states), _c(globals10)
        n = 0 if s is None else s.shape[0]
        _check(self._L.v2gt_batch_set_estimate(self.h, int(trial), n, dptr(s), dptr(g)), "v2gt_batch_set_estimate")

    def linearize(self, trial=0):
        n = self.info(trial).n_states
        D, O, B, g = np.zeros((n, 225)), np.zeros((n, 225)), np.zeros((n, 135)), np.zeros((n, 15))
        G, gG = np.zeros(81), np.zeros(9)
        cost, lin0 = C.c_double(), C.c_double()
        _check(self._L.v2gt_eval_linearize(self.h, int(trial), n, dptr(D), dptr(O), dptr(B), dptr(g), dptr(G), dptr(gG),
                                           C.byref(cost), C.byref(lin0)), "v2gt_eval_linearize")
        return dict(D=D, O=O, B=B, g=g, Gamma=G, g_gamma=gG, cost=cost.value, lin_error0=lin0.value)

    def solve(self, trial, lam):
        n = self.info(trial).n_states
        delta = np.zeros(15 * n + 9)
        ok = C.c_int32()
        _check(self._L.v2gt_eval_solve(self.h, int(trial), float(lam), len(delta), dptr(delta), C.byref(ok)), "v2gt_eval_solve")
        return delta, bool(ok.value)

    def cost(self, trial=0):
        c = C.c_double()
        _check(self._L.v2gt_eval_cost(self.h, int(trial), C.byref(c)), "v2gt_eval_cost")
        return c.value

    def lm_trace(self, trial=0):
        n = C.c_int64()
        _check(self._L.v2gt_batch_get_lm_trace(self.h, int(trial), 0, None, C.byref(n)), "v2gt_batch_get_lm_trace")
        rows = np.zeros((n.value, 5))
        if n.value:
            _check(self._L.v2gt_batch_get_lm_trace(self.h, int(trial), n.value, dptr(rows), C.byref(n)), "v2gt_batch_get_lm_trace")
        return rows

    def timing(self):
        ms = np.zeros(8)
        launches = np.zeros(8, dtype=np.int64)
        self._L.v2gt_batch_get_timing(self.h, dptr(ms), launches.ctypes.data_as(c_int64_p))
        return ms, launches


class WaveRunner:
    """A Monte-Carlo batch larger than one GPU's memory, solved in waves (scripts/run_monte_carlo.sh runs its trials one
    roslaunch after the other; here `wave` trials at a time).  Two resident batch handles: while wave k is being solved,
    a host thread stages and uploads wave k+1 into the other handle (pinned fill + H2D copies on that handle's stream run
    under wave k's kernels) and another reads wave k-1's results back, so the GPU never waits for the host.

        runner = WaveRunner(wave=128)
        for result in runner.run(params, datasets):      # datasets: any sequence of Dataset
            first_trial, offsets, times, states, globals10, infos = result
    """

    def __init__(self, wave=128, device=0, **options):
        self.wave = int(wave)
        self.device = int(device)
        self.options = options
        self.handles = []
        self.device_ms = 0.0     # CUDA-event time of build + optimise, summed over the waves of the last run
        self.wall_s = 0.0        # first staging call -> last result read back
        self.stall_s = 0.0       # time the solving thread spent waiting for an upload to finish (0 when fully hidden)

    def close(self):
        for h in self.handles:
            h.close()
        self.handles = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _handle(self, k, n):
        while len(self.handles) <= k:
            self.handles.append(None)
        h = self.handles[k]
        if h is None or h.n_trials != n:
            if h is not None:
                h.close()
            h = self.handles[k] = BatchSolver(n, device=self.device, **self.options)
        return h

    def run(self, params, datasets, read_back=True):
        import threading
        import time

        n_total = len(datasets)
        starts = list(range(0, n_total, self.wave))
        self.device_ms = self.stall_s = 0.0
        t_begin = time.perf_counter()
        errors = []

        def stage(w):
            try:
                a = starts[w]
                chunk = datasets[a:a + self.wave]
                h = self._handle(w % 2, len(chunk))
                for i, ds in enumerate(chunk):
                    h.set_trial_dataset(i, params, ds, borrow=True)
                h.upload()
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        def fetch(w, out):
            try:
                h = self.handles[w % 2]
                infos = [h.info(i) for i in range(h.n_trials)]
                out.append((starts[w],) + (h.all_states() if read_back else (None, None, None, None)) + (infos,))
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        results = [[] for _ in starts]
        up = threading.Thread(target=stage, args=(0,))
        up.start()
        rb = None
        for w in range(len(starts)):
            t0 = time.perf_counter()
            up.join()
            self.stall_s += time.perf_counter() - t0 if w else 0.0
            if errors:
                raise errors[0]
            if w + 1 < len(starts):
                if rb is not None:  # wave w-1 lives in the handle wave w+1 is about to overwrite
                    rb.join()
                    rb = None
                up = threading.Thread(target=stage, args=(w + 1,))
                up.start()
            h = self.handles[w % 2]
            h.build_and_solve()
            ms, _ = h.timing()
            self.device_ms += float(ms[0] + ms[7])
            if rb is not None:
                rb.join()
            rb = threading.Thread(target=fetch, args=(w, results[w]))
            rb.start()
            if w >= 1 and results[w - 1]:
                yield results[w - 1][0]
                results[w - 1] = None
        if rb is not None:
            rb.join()
        if errors:
            raise errors[0]
        self.wall_s = time.perf_counter() - t_begin
        for r in results:
            if r:
                yield r[0]


class Propagator:
    """IMU store of the reference (src/meas/Propagator.h:41-55): holds the noise densities, appends samples."""

    def __init__(self, sigma_w, sigma_wb, sigma_a, sigma_ab):
        self.sigmas = (float(sigma_w), float(sigma_wb), float(sigma_a), float(sigma_ab))
        self._t, self._w, self._a = [], [], []

    def feed_imu(self, timestamp, wm, am):
        self._t.append(float(timestamp))
        self._w.append(np.asarray(wm, dtype=np.float64).reshape(3))
        self._a.append(np.asarray(am, dtype=np.float64).reshape(3))

    def feed_imu_array(self, t, w, a):
        self._t.extend(np.asarray(t, dtype=np.float64).tolist())
        self._w.extend(np.asarray(w, dtype=np.float64).reshape(-1, 3))
        self._a.extend(np.asarray(a, dtype=np.float64).reshape(-1, 3))

    def arrays(self):
        n = len(self._t)
        return (np.asarray(self._t, dtype=np.float64), np.asarray(self._w, dtype=np.float64).reshape(n, 3),
                np.asarray(self._a, dtype=np.float64).reshape(n, 3))


class Interpolator:
    """Vicon pose store of the reference (src/meas/Interpolator.h:55-81); sorting / de-duplication with
    std::set semantics is done by the library when the trial is staged."""

    def __init__(self):
        self._t, self._q, self._p, self._Rq, self._Rp = [], [], [], [], []

    def feed_pose(self, timestamp, q, p, R_q, R_p):
        self._t.append(float(timestamp))
        self._q.append(np.asarray(q, dtype=np.float64).reshape(4))
        self._p.append(np.asarray(p, dtype=np.float64).reshape(3))
        self._Rq.append(np.asarray(R_q, dtype=np.float64).reshape(9))
        self._Rp.append(np.asarray(R_p, dtype=np.float64).reshape(9))

    def arrays(self):
        n = len(self._t)
        return (np.asarray(self._t), np.asarray(self._q).reshape(n, 4), np.asarray(self._p).reshape(n, 3),
                np.asarray(self._Rq).reshape(n, 9), np.asarray(self._Rp).reshape(n, 9))

    def get_time_min(self):
        return min(self._t) if self._t else float("inf")

    def get_time_max(self):
        return max(self._t) if self._t else float("-inf")


class ViconGraphSolver:
    """Single-trajectory solver with the reference's method names (minus the ROS-only `visualize`)."""

    def __init__(self, params, propagator, interpolator, timestamp_cameras, device=0, **options):
        self.params = params
        sw, swb, sa, sab = propagator.sigmas
        params.sigma_w, params.sigma_wb, params.sigma_a, params.sigma_ab = sw, swb, sa, sab
        self._batch = BatchSolver(1, device=device, **options)
        it, iw, ia = propagator.arrays()
        vt, vq, vp, vRq, vRp = interpolator.arrays()
        self._batch.set_trial(0, params, it, iw, ia, vt, vq, vp, np.asarray(timestamp_cameras, dtype=np.float64), vic_Rq=vRq,
                              vic_Rp=vRp)

    def build_and_solve(self):
        self._batch.build_and_solve()
        inf = self._batch.info(0)
        if inf.status != 0:
            # the reference calls std::exit(EXIT_FAILURE) here (ViconGraphSolver.cpp:99-150, 508-524)
            raise V2gtError(inf.status, "build_and_solve")
        return inf

    def write_to_file(self, csvfilepath, infofilepath):
        self._batch.write_to_file(0, csvfilepath, infofilepath)

    def get_imu_poses(self):
        """times, poses[n][10] = [q(4) p(3) v(3)] as the reference's get_imu_poses (.cpp:357-375)."""
        t, x = self._batch.states(0)
        poses = np.concatenate([x[:, 0:4], x[:, 13:16], x[:, 7:10]], axis=1)
        return t, poses

    def get_calibration(self):
        """toff, R_BtoI, p_BinI, R_GtoV as the reference's get_calibration (.cpp:377-388)."""
        q, p, toff, th = self._batch.calibration(0)
        x, y, z, w = q
        # quat_2_Rot (JPL): (2w^2-1) I - 2w [v]x + 2 v v^T
        v = np.array([x, y, z])
        vx = np.array([[0, -z, y], [z, 0, -x], [-y, x, 0]])
        R_BtoI = (2 * w * w - 1) * np.eye(3) - 2 * w * vx + 2 * np.outer(v, v)
        cx, sx, cy, sy = np.cos(th[0]), np.sin(th[0]), np.cos(th[1]), np.sin(th[1])
        R_GtoV = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
        return toff, R_BtoI, p, R_GtoV

    @property
    def batch(self):
        return self._batch
