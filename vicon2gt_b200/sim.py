"""Synthetic input generator (host C++ library libv2gt_sim.so, see host/sim.cpp).

Produces the IMU / vicon streams and state-time grid of the reference's `run_simulation`
(src/run_simulation.cpp:108-158) for the BASELINE.json configurations.  The reference's
trajectory text files are not shipped with this repo; `make_trajectory` synthesises a smooth
trajectory of the same shape (start time, row count, spacing).
"""
import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from ._cdefs import SimParams, c_double_p, c_int32_p, dptr
from .build import lib_path

_lib = None


def _load():
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path("sim")
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `python -c 'import __graft_entry__ as g; g.build()'` first")
    lib = C.CDLL(path)
    lib.v2gt_sim_create.restype = C.c_void_p
    lib.v2gt_sim_create.argtypes = [C.POINTER(SimParams)]
    lib.v2gt_sim_destroy.argtypes = [C.c_void_p]
    lib.v2gt_sim_error.restype = C.c_char_p
    lib.v2gt_sim_error.argtypes = [C.c_void_p]
    lib.v2gt_sim_params_default.argtypes = [C.POINTER(SimParams)]
    lib.v2gt_sim_load_trajectory_file.argtypes = [C.c_void_p, C.c_char_p]
    lib.v2gt_sim_set_trajectory.argtypes = [C.c_void_p, C.c_int64, c_double_p]
    lib.v2gt_sim_make_trajectory.argtypes = [C.c_double, C.c_int64, C.c_double, C.c_double, c_double_p]
    lib.v2gt_sim_run.argtypes = [C.c_void_p]
    for n in ("v2gt_sim_n_imu", "v2gt_sim_n_vicon", "v2gt_sim_n_states"):
        getattr(lib, n).restype = C.c_int64
        getattr(lib, n).argtypes = [C.c_void_p]
    lib.v2gt_sim_get.argtypes = [C.c_void_p] + [c_double_p] * 7
    lib.v2gt_sim_get_truth.argtypes = [C.c_void_p, c_double_p, c_double_p, c_double_p, c_double_p]
    lib.v2gt_sim_get_state_in_vicon.argtypes = [C.c_void_p, C.c_int64, c_double_p, c_double_p, c_int32_p]
    _lib = lib
    return lib


@dataclass
class Dataset:
    """One trial's inputs, as fed to Propagator::feed_imu / Interpolator::feed_pose / the solver ctor."""

    imu_t: np.ndarray
    imu_w: np.ndarray
    imu_a: np.ndarray
    vic_t: np.ndarray
    vic_q: np.ndarray
    vic_p: np.ndarray
    vic_R_const: np.ndarray  # 18 doubles: R_q (3x3), R_p (3x3) for every pose (run_simulation.cpp:84-96)
    state_t: np.ndarray
    truth: dict = field(default_factory=dict)
    seed: int = 0
    vic_Rq: np.ndarray = None  # optional per-pose covariances [V][9] (real data with Odometry covariances, ingest.py)
    vic_Rp: np.ndarray = None

    @property
    def n_imu(self):
        return len(self.imu_t)

    @property
    def n_vicon(self):
        return len(self.vic_t)

    @property
    def n_times(self):
        return len(self.state_t)

    def nbytes(self):
        return sum(a.nbytes for a in (self.imu_t, self.imu_w, self.imu_a, self.vic_t, self.vic_q, self.vic_p, self.state_t))


def default_sim_params(**kw):
    p = SimParams()
    _load().v2gt_sim_params_default(C.byref(p))
    for k, v in kw.items():
        if k == "sigma_vicon_pose":
            for i in range(6):
                p.sigma_vicon_pose[i] = float(v[i])
        else:
            setattr(p, k, v)
    return p


def make_trajectory(t_start, n_rows, dt=0.05, extent_m=20.0):
    """Synthesised trajectory rows `t x y z qx qy qz qw` (JPL q_GtoI)."""
    rows = np.zeros((n_rows, 8))
    _load().v2gt_sim_make_trajectory(float(t_start), int(n_rows), float(dt), float(extent_m), dptr(rows))
    return rows


# Shapes of the reference's trajectory files (start time, rows, mean spacing) used by BASELINE.json's
# configs; the files themselves are not shipped (data/tum_corridor1_512_16_okvis.txt, data/euroc_V1_01_easy.txt).
TRAJ_SHAPES = {
    "tum_corridor1": dict(t_start=1520531829.301144123, n_rows=5986, dt=0.05, extent_m=20.0),
    "euroc_V1_01": dict(t_start=1403715273.26214, n_rows=2895, dt=0.05, extent_m=2.5),
    # BASELINE.json configs[4]: no reference data file is an hour long -- synthesised
    "one_hour": dict(t_start=1520531829.301144123, n_rows=72041, dt=0.05, extent_m=20.0),
}


def simulate(traj_rows=None, traj_path=None, seed=0, vicon_sigmas=(0.0174, 0.0174, 0.0174, 0.05, 0.05, 0.05),
             freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, state_freq=100, max_imu=0, fixed_truth=None, **noise):
    """Run the simulator once and return a Dataset (defaults = launch/sim.launch of the reference)."""
    lib = _load()
    p = default_sim_params(seed=int(seed), sim_freq_imu=float(freq_imu), sim_freq_cam=float(freq_cam),
                           sim_freq_vicon=float(freq_vicon), state_freq=int(state_freq), max_imu=int(max_imu),
                           sigma_vicon_pose=vicon_sigmas, **noise)
    if fixed_truth is not None:
        p.use_fixed_truth = 1
        for i in range(9):
            p.fixed_R_BtoI[i] = float(np.asarray(fixed_truth["R_BtoI"]).reshape(-1)[i])
            p.fixed_R_GtoV[i] = float(np.asarray(fixed_truth["R_GtoV"]).reshape(-1)[i])
        for i in range(3):
            p.fixed_p_BinI[i] = float(fixed_truth["p_BinI"][i])
        p.fixed_viconimu_dt = float(fixed_truth["viconimu_dt"])
    h = lib.v2gt_sim_create(C.byref(p))
    try:
        if traj_path is not None:
            if lib.v2gt_sim_load_trajectory_file(h, traj_path.encode()) != 0:
                raise RuntimeError(lib.v2gt_sim_error(h).decode())
        else:
            rows = np.ascontiguousarray(traj_rows, dtype=np.float64)
            if lib.v2gt_sim_set_trajectory(h, rows.shape[0], dptr(rows)) != 0:
                raise RuntimeError("trajectory too short")
        if lib.v2gt_sim_run(h) != 0:
            raise RuntimeError(lib.v2gt_sim_error(h).decode())
        M, V, N = lib.v2gt_sim_n_imu(h), lib.v2gt_sim_n_vicon(h), lib.v2gt_sim_n_states(h)
        imu_t, imu_w, imu_a = np.zeros(M), np.zeros((M, 3)), np.zeros((M, 3))
        vic_t, vic_q, vic_p = np.zeros(V), np.zeros((V, 4)), np.zeros((V, 3))
        state_t = np.zeros(N)
        lib.v2gt_sim_get(h, dptr(imu_t), dptr(imu_w), dptr(imu_a), dptr(vic_t), dptr(vic_q), dptr(vic_p), dptr(state_t))
        R_BtoI, p_BinI, R_GtoV = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        dt = C.c_double()
        lib.v2gt_sim_get_truth(h, dptr(R_BtoI), dptr(p_BinI), C.byref(dt), dptr(R_GtoV))
        truth = dict(R_BtoI=R_BtoI, p_BinI=p_BinI, viconimu_dt=dt.value, R_GtoV=R_GtoV)
        # truth states at the state times (Simulator::get_state_in_vicon), for accuracy checks
        st = np.zeros((N, 17))
        ok = np.zeros(N, dtype=np.int32)
        if N:
            lib.v2gt_sim_get_state_in_vicon(h, N, dptr(state_t), dptr(st), ok.ctypes.data_as(c_int32_p))
        truth["states"] = st
        truth["states_ok"] = ok
    finally:
        lib.v2gt_sim_destroy(h)
    R_const = _vic_R_const(vicon_sigmas)
    return Dataset(imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, R_const, state_t, truth, int(seed))


def _vic_R_const(sigmas):
    s = np.asarray(sigmas, dtype=np.float64)
    R = np.zeros(18)
    R[[0, 4, 8]] = s[:3] ** 2  # R_q diag (run_simulation.cpp:89-91)
    R[[9, 13, 17]] = s[3:] ** 2  # R_p diag
    return R


class SimBatch:
    """The simulator for a whole Monte-Carlo batch on the GPU (csrc/k_sim.cu, `v2gt_simbatch_*` of the C ABI).

    One trajectory, `n_trials` seeds (default: trial t has seed t) and per-trial vicon sigmas; the rates are shared.
    `plan()` (host only) fixes the stamps, the per-seed truth and the state grids; `run()` generates the measurements of
    every trial on the device, where `BatchSolver.set_trial_sim` picks them up without a host round trip.  `dataset(t)`
    downloads one trial in the host simulator's layout (parity checks, examples)."""

    def __init__(self, traj_rows, n_trials, device=0, seeds=None, vicon_sigmas=(0.0174, 0.0174, 0.0174, 0.05, 0.05, 0.05),
                 freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, state_freq=100, max_imu=0, **noise):
        from . import api

        self._L = api.load_library()
        self._api = api
        rows = np.ascontiguousarray(traj_rows, dtype=np.float64)
        self.n_trials = int(n_trials)
        self.h = C.c_void_p()
        api._check(self._L.v2gt_simbatch_create(int(device), self.n_trials, rows.shape[0], dptr(rows), C.byref(self.h)),
                   "v2gt_simbatch_create")
        sig = np.asarray(vicon_sigmas, dtype=np.float64)
        self._sigmas = np.broadcast_to(sig, (self.n_trials, 6)) if sig.ndim == 1 else sig.reshape(self.n_trials, 6)
        self.seeds = [int(x) for x in (range(self.n_trials) if seeds is None else seeds)]
        for t in range(self.n_trials):
            p = default_sim_params(seed=self.seeds[t], sim_freq_imu=float(freq_imu), sim_freq_cam=float(freq_cam),
                                   sim_freq_vicon=float(freq_vicon), state_freq=int(state_freq), max_imu=int(max_imu),
                                   sigma_vicon_pose=self._sigmas[t], **noise)
            api._check(self._L.v2gt_simbatch_set_trial(self.h, t, C.byref(p)), "v2gt_simbatch_set_trial")

    def close(self):
        if getattr(self, "h", None):
            self._L.v2gt_simbatch_destroy(self.h)
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

    def plan(self):
        self._api._check(self._L.v2gt_simbatch_plan(self.h), "v2gt_simbatch_plan")

    def run(self):
        self._api._check(self._L.v2gt_simbatch_run(self.h), "v2gt_simbatch_run")

    def generate_ms(self):
        v = C.c_double()
        self._api._check(self._L.v2gt_simbatch_get_timing(self.h, C.byref(v)), "v2gt_simbatch_get_timing")
        return v.value

    def sizes(self, trial=0):
        a, b, c = C.c_int64(), C.c_int64(), C.c_int64()
        self._api._check(self._L.v2gt_simbatch_sizes(self.h, int(trial), C.byref(a), C.byref(b), C.byref(c)), "v2gt_simbatch_sizes")
        return a.value, b.value, c.value

    def vic_R_const(self, trial):
        return _vic_R_const(self._sigmas[int(trial)])

    def truth(self, trial):
        R_BtoI, p_BinI, R_GtoV = np.zeros((3, 3)), np.zeros(3), np.zeros((3, 3))
        dt = C.c_double()
        self._api._check(self._L.v2gt_simbatch_get_truth(self.h, int(trial), dptr(R_BtoI), dptr(p_BinI), C.byref(dt), dptr(R_GtoV)),
                         "v2gt_simbatch_get_truth")
        return dict(R_BtoI=R_BtoI, p_BinI=p_BinI, viconimu_dt=dt.value, R_GtoV=R_GtoV)

    def stamps(self, trial):
        """(imu_t, vic_t, state_t) of a trial: host plan only, no device needed."""
        M, V, N = self.sizes(trial)
        imu_t, vic_t, state_t = np.zeros(M), np.zeros(V), np.zeros(N)
        self._api._check(self._L.v2gt_simbatch_get(self.h, int(trial), dptr(imu_t), None, None, dptr(vic_t), None, None, dptr(state_t)),
                         "v2gt_simbatch_get")
        return imu_t, vic_t, state_t

    def state_in_vicon(self, trial, times):
        times = np.ascontiguousarray(times, dtype=np.float64)
        out = np.zeros((len(times), 17))
        ok = np.zeros(len(times), dtype=np.int32)
        self._api._check(self._L.v2gt_simbatch_get_state_in_vicon(self.h, int(trial), len(times), dptr(times), dptr(out),
                                                                  ok.ctypes.data_as(c_int32_p)), "v2gt_simbatch_get_state_in_vicon")
        return out, ok

    def dataset(self, trial, with_truth_states=True):
        M, V, N = self.sizes(trial)
        imu_t, imu_w, imu_a = np.zeros(M), np.zeros((M, 3)), np.zeros((M, 3))
        vic_t, vic_q, vic_p = np.zeros(V), np.zeros((V, 4)), np.zeros((V, 3))
        state_t = np.zeros(N)
        self._api._check(self._L.v2gt_simbatch_get(self.h, int(trial), dptr(imu_t), dptr(imu_w), dptr(imu_a), dptr(vic_t), dptr(vic_q),
                                                   dptr(vic_p), dptr(state_t)), "v2gt_simbatch_get")
        truth = self.truth(trial)
        if with_truth_states and N:
            truth["states"], truth["states_ok"] = self.state_in_vicon(trial, state_t)
        return Dataset(imu_t, imu_w, imu_a, vic_t, vic_q, vic_p, self.vic_R_const(trial), state_t, truth, self.seeds[int(trial)])


# The reference's own trajectory files for BASELINE.json's configs (data/euroc_V1_01_easy.txt, data/tum_corridor1_512_16_okvis.txt:
# rows `t x y z qx qy qz qw`), kept as fixtures under tests/golden/ (V2GT_TRAJ_DIR overrides the directory).  The TUM file is
# jittered (row spacing 0.0431 ... 0.0550 s), which the synthesised stand-ins of the same shape are not.
TRAJ_FILES = {"tum_corridor1": "traj_tum_corridor1_512_16_okvis.txt", "euroc_V1_01": "traj_euroc_V1_01_easy.txt"}
_traj_cache = {}


def trajectory_file(shape_name):
    """Path of the reference's trajectory file for a shape, or None where it is not shipped (one_hour) / not found."""
    import os

    fn = TRAJ_FILES.get(shape_name)
    if fn is None:
        return None
    for d in (os.environ.get("V2GT_TRAJ_DIR"), os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")):
        if d and os.path.exists(os.path.join(d, fn)):
            return os.path.join(d, fn)
    return None


def trajectory_rows(shape_name, duration_s=None, source=None):
    """(rows [n][8], source) of a trajectory shape.  source: "file" = the reference's data file (error if absent), "synth" =
    the synthesised stand-in of the same start time / row count / spacing, None = the file where there is one."""
    path = trajectory_file(shape_name) if source in (None, "file") else None
    if source == "file" and path is None:
        raise FileNotFoundError(f"no trajectory file for {shape_name} (tests/golden/{TRAJ_FILES.get(shape_name)})")
    if path is not None:
        if path not in _traj_cache:
            # the reference's loader (Simulator.cpp:281-308): space-separated, lines starting with '#' skipped, atof per field
            _traj_cache[path] = np.loadtxt(path, comments="#", dtype=np.float64, ndmin=2)[:, :8].copy()
        rows = _traj_cache[path]
        if duration_s is not None:
            rows = rows[: max(8, int(np.searchsorted(rows[:, 0], rows[0, 0] + duration_s, "right")))]
        return np.ascontiguousarray(rows), "file:" + path.rsplit("/", 1)[-1]
    shape = dict(TRAJ_SHAPES[shape_name])
    if duration_s is not None:
        shape["n_rows"] = max(8, min(shape["n_rows"], int(round(duration_s / shape["dt"])) + 1))
    return make_trajectory(**shape), "synthesised, %s shape" % shape_name


CONFIG_SHAPES = {
    # name: (trajectory shape, simulator rates).  state_freq 0: states at the trajectory timestamps (run_simulation.cpp:149-158)
    "C1": ("tum_corridor1", dict(freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, state_freq=100)),
    "C2": ("euroc_V1_01", dict(freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0)),
    "C3": ("tum_corridor1", dict(freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0)),
    "C5": ("one_hour", dict(freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0)),
}


def make_config(name, seed=0, duration_s=None, trajectory=None, **kw):
    """Datasets of BASELINE.json's configs.

    name: "C1" sim.launch defaults (data/tum_corridor1_512_16_okvis.txt, 400 Hz IMU, 100 Hz vicon, 100 Hz states)
          "C2" EuRoC V1_01: states at the timestamps of data/euroc_V1_01_easy.txt (20 Hz), 200 Hz IMU, 100 Hz vicon
          "C3" TUM corridor1: states at the (jittered) timestamps of data/tum_corridor1_512_16_okvis.txt, 200 Hz IMU, 100 Hz vicon
          "C5" one-hour sequence (no reference file is that long: synthesised), 20 Hz states (~72 k), 200 Hz IMU (~720 k)
    trajectory: "file" (the reference's data file), "synth" (synthesised stand-in of the same shape) or None (file where
    there is one).  duration_s truncates the trajectory (smaller problems for tests).
    """
    if name not in CONFIG_SHAPES:
        raise ValueError(name)
    shape_name, sim_kw = CONFIG_SHAPES[name]
    rows, src = trajectory_rows(shape_name, duration_s, trajectory)
    sim_kw = dict(sim_kw)
    sim_kw.update(kw)
    ds = simulate(traj_rows=rows, seed=seed, **sim_kw)
    ds.truth["trajectory"] = src
    return ds
