"""Chain-sharded build-and-solve of ONE long sequence across GPUs (BASELINE.json config 5).

The cleaned state chain is cut into `world x segs_per_rank` segments with the single-GPU solver's geometry
(`v2gt_chain_plan` of the library is the one statement of it); rank r holds the states of its segments (+ one halo
state on each inner side), linearises and eliminates them locally, and per damped solve the ranks exchange twice:
    A  the Schur outputs of their segments, the normal-equation records of their separators and their part of the 9x9
       global block / gradient -- ONE all-gather (the global block is then summed in rank order on every rank),
    B  the cost scalars of their own factors + the step of their last own state (the neighbour's halo) -- one
       24-double all-gather (B needs the step, the step needs the reduced system of A: the two cannot be merged),
after which every rank solves the small reduced system (separators + 9 globals) redundantly and takes the identical
Levenberg-Marquardt decision.

Two drivers over the same kernels and buffers:
  * `NativeComm` (one process per GPU): the whole LM loop runs INSIDE the library (`v2gt_chain_optimize`): both
    exchanges are `ncclAllGather` calls on the solver's stream, one damped solve is a captured CUDA graph.  Python only
    hands the NCCL unique id from rank 0 to the other ranks (through `torch.distributed`, any backend).
  * `LocalComm`: all ranks of a small world in THIS process on one GPU, phase by phase (`v2gt_chain_phase`), the
    exchanges as tensor copies in rank order -- tests the sharded algorithm without a multi-GPU box.
Everything numeric happens in libvicon2gt_b200.so.
"""
import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import api

(BUF_SUMS, BUF_GAMMA, BUF_PACK_SEND, BUF_SEG_ALL, BUF_PACK_ALL, BUF_SEPNE_ALL, BUF_SUMS_SEND, BUF_SUMS_ALL) = range(8)
(PH_COST0, PH_BEGIN, PH_LINEARIZE, PH_ELIMINATE, PH_SOLVE, PH_COST, PH_DECIDE) = range(7)
NCCL_ID_BYTES = 128


class _CPlan(C.Structure):
    _fields_ = [("rank", C.c_int32), ("world", C.c_int32), ("p_total", C.c_int32), ("stride", C.c_int32), ("j0", C.c_int32),
                ("j1", C.c_int32), ("own_lo", C.c_int32), ("own_hi", C.c_int32), ("n_total", C.c_int64), ("g_off", C.c_int64),
                ("n_local", C.c_int64)]


@dataclass
class RankPlan:
    rank: int
    world: int
    n_total: int      # cleaned states of the whole chain
    p_total: int      # segments of the whole chain
    stride: int
    j0: int           # own segments [j0, j1)
    j1: int
    g_off: int        # global index of local slot 0
    n_local: int      # local slots (own + halos)
    own_lo: int       # own local slots [own_lo, own_hi)
    own_hi: int

    @property
    def own_global(self):
        return self.g_off + self.own_lo, self.g_off + self.own_hi


def _lib():
    L = api.load_library()
    if not getattr(L, "_chain_bound", False):
        vp = C.c_void_p
        L.v2gt_chain_plan.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_CPlan)]
        L.v2gt_chain_fit_segments.argtypes = [C.c_int64, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
        L.v2gt_chain_configure.argtypes = [vp, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32]
        L.v2gt_chain_phase.argtypes = [vp, C.c_int32]
        L.v2gt_chain_buffer.argtypes = [vp, C.c_int32, C.POINTER(vp), C.POINTER(C.c_int64)]
        L.v2gt_batch_stream.argtypes = [vp, C.POINTER(vp)]
        L.v2gt_chain_nccl_unique_id.argtypes = [C.c_char_p]
        L.v2gt_chain_comm_init.argtypes = [vp, C.c_char_p]
        L.v2gt_chain_optimize.argtypes = [vp, C.c_int32, C.c_int32]
        L._chain_bound = True
    return L


def seg_stride(n, p):
    """Segment stride of the solver's geometry (csrc/k_solve.cu seg_geom): separators at m * stride - 1."""
    return max(2, (n + 1 + p - 1) // p)


def chain_plan(n_total, world, segs_per_rank):
    """Ownership of a chain of n_total states (v2gt_chain_plan, host only): rank r owns segments [r*spr, (r+1)*spr) and the
    separators in front of them, i.e. the contiguous states [j0*stride - 1, j1*stride - 1) (rank 0 starts at 0, the last rank
    ends at n)."""
    L = _lib()
    plans = []
    for r in range(world):
        c = _CPlan()
        if L.v2gt_chain_plan(int(n_total), int(world), int(segs_per_rank), r, C.byref(c)) != 0:
            raise ValueError(f"chain of {n_total} states is too short for {world * segs_per_rank} segments")
        plans.append(RankPlan(c.rank, c.world, c.n_total, c.p_total, c.stride, c.j0, c.j1, c.g_off, c.n_local, c.own_lo, c.own_hi))
    return plans


def fit_segments_per_rank(n_total, world, target):
    """Largest segments-per-rank <= target whose geometry uses every segment (ceil rounding of the stride can leave the
    last segments of a finer cut empty)."""
    out = C.c_int32()
    if _lib().v2gt_chain_fit_segments(int(n_total), int(world), int(max(1, target)), C.byref(out)) != 0 or out.value < 1:
        raise ValueError(f"chain of {n_total} states cannot be cut for {world} ranks")
    return out.value


def clean_timestamps(params, imu_t, vic_t, state_t):
    """ViconGraphSolver::build_and_solve timestamp cleaning (ViconGraphSolver.cpp:114-142) on the whole sequence."""
    ts = np.asarray(state_t, dtype=np.float64)
    first = vic_t.min() + params.toff_imu_to_vicon + 2 * params.time_offset_window
    last = vic_t.max() + params.toff_imu_to_vicon - 2 * params.time_offset_window
    keep = (ts >= imu_t[0]) & (ts <= imu_t[-1]) & (ts >= first) & (ts <= last)
    return ts[keep]


def vicon_margin(params, toff_drift=1.0):
    """Seconds of vicon poses a rank keeps beyond its first / last local state: the initial guess looks
    +-time_offset_window around t - t_off and accepts brackets up to interp_gap_init away (ViconGraphSolver.cpp:424-483,
    Interpolator.cpp:110-123); the local re-run of the timestamp cleaning needs 2 windows + |t_off| inside the local pose
    range; `toff_drift` is what the time offset may move during the optimisation before a boundary state would lose its
    bracket (checked after the solve, ChainSolver.check_brackets)."""
    return abs(params.toff_imu_to_vicon) + 2 * params.time_offset_window + params.interp_gap_init + toff_drift


def slice_inputs(ds, ts_local, vicon_margin_s=1.0, imu_margin=0.05):
    """Inputs a rank needs for its local states: vicon poses within the margin, IMU samples bracketing the range."""
    t0, t1 = ts_local[0], ts_local[-1]
    vi0 = max(0, np.searchsorted(ds.vic_t, t0 - vicon_margin_s, "left") - 1)
    vi1 = min(len(ds.vic_t), np.searchsorted(ds.vic_t, t1 + vicon_margin_s, "right") + 1)
    ii0 = max(0, np.searchsorted(ds.imu_t, t0 - imu_margin, "left") - 2)
    ii1 = min(len(ds.imu_t), np.searchsorted(ds.imu_t, t1 + imu_margin, "right") + 2)
    return slice(ii0, ii1), slice(vi0, vi1)


class _DevBuf:
    """A device buffer of the library as a __cuda_array_interface__ object (wrapped by torch.as_tensor, no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class _Part:
    """One rank's share: a 1-trial batch in chain mode (+ torch views of its exchange buffers for the phase driver)."""

    def __init__(self, params, ds, ts_clean, plan, device, torch, margin):
        self.plan = plan
        L = _lib()
        self._L = L
        self.bs = api.BatchSolver(1, device=device)
        ts_local = ts_clean[plan.g_off: plan.g_off + plan.n_local]
        si, sv = slice_inputs(ds, ts_local, margin)
        self.vic_range = (float(ds.vic_t[sv][0]), float(ds.vic_t[sv][-1]), sv.start == 0, sv.stop == len(ds.vic_t))
        self.t_range = (float(ts_local[0]), float(ts_local[-1]))
        self.bs.set_trial(0, params, ds.imu_t[si], ds.imu_w[si], ds.imu_a[si], ds.vic_t[sv], ds.vic_q[sv], ds.vic_p[sv], ts_local,
                          vic_R_const=ds.vic_R_const)
        api._check(L.v2gt_chain_configure(self.bs.h, plan.rank, plan.world, plan.n_total, plan.g_off, plan.p_total, plan.j0, plan.j1,
                                          plan.own_lo, plan.own_hi), "v2gt_chain_configure")
        self.bs.upload()
        self.bs.build(True)
        inf = self.bs.info(0)
        if inf.status != 0 or inf.n_states != plan.n_local:
            raise api.V2gtError(inf.status or 1, f"chain build on rank {plan.rank}: {inf.n_states} of {plan.n_local} local states survive")
        self.buf = {}
        self.stream = None
        if torch is not None:
            for which in range(8):
                p, n = C.c_void_p(), C.c_int64()
                api._check(L.v2gt_chain_buffer(self.bs.h, which, C.byref(p), C.byref(n)), "v2gt_chain_buffer")
                self.buf[which] = torch.as_tensor(_DevBuf(p.value, n.value), device=f"cuda:{device}")
            sp = C.c_void_p()
            api._check(L.v2gt_batch_stream(self.bs.h, C.byref(sp)), "v2gt_batch_stream")
            self.stream = torch.cuda.ExternalStream(sp.value, device=f"cuda:{device}")

    def phase(self, ph):
        api._check(self._L.v2gt_chain_phase(self.bs.h, ph), f"v2gt_chain_phase({ph})")


class LocalComm:
    """All ranks of the world in this process (one GPU), driven phase by phase: an exchange is a concatenation in rank
    order copied into every rank's receive buffer."""

    native = False

    def __init__(self, world):
        self.world = world
        self.rank = 0
        self.ranks = list(range(world))

    def all_gather(self, parts, send, recv, torch):
        for p in parts:
            p.bs.sync()
        cat = torch.cat([p.buf[send] for p in parts])
        for p in parts:
            p.buf[recv][: cat.numel()].copy_(cat)
        torch.cuda.synchronize()


class NativeComm:
    """One rank per process; the exchanges are ncclAllGather calls made by the library itself on the solver's stream.
    `dist`: an initialised torch.distributed (any backend) used ONCE, to hand rank 0's NCCL unique id to the other
    ranks; None for a world of one (no NCCL at all)."""

    native = True

    def __init__(self, dist=None):
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.ranks = [self.rank]

    def init_comm(self, part):
        if self.world == 1:
            return
        L = _lib()
        idb = C.create_string_buffer(NCCL_ID_BYTES)
        if self.rank == 0:
            api._check(L.v2gt_chain_nccl_unique_id(idb), "v2gt_chain_nccl_unique_id")
        box = [bytes(idb.raw)]
        self.dist.broadcast_object_list(box, src=0)
        api._check(L.v2gt_chain_comm_init(part.bs.h, box[0]), "v2gt_chain_comm_init")


DistComm = NativeComm  # name used by the first round's scripts


class ChainSolver:
    """build_and_solve of one long sequence, sharded over `comm.world` ranks."""

    def __init__(self, params, ds, comm, segs_per_rank=16, device=0, use_graph=True):
        torch = None
        if not comm.native:
            import torch
        self.torch = torch
        self.comm = comm
        self.params = params
        self.use_graph = use_graph
        self.ts = clean_timestamps(params, ds.imu_t, ds.vic_t, ds.state_t)
        self.segs_per_rank = fit_segments_per_rank(len(self.ts), comm.world, segs_per_rank)
        self.plans = chain_plan(len(self.ts), comm.world, self.segs_per_rank)
        margin = vicon_margin(params)
        self.parts = [_Part(params, ds, self.ts, self.plans[r], device, torch, margin) for r in comm.ranks]
        if comm.native:
            comm.init_comm(self.parts[0])
        self.tries = 0
        self.check_every = 4  # phase driver: the host reads the replicated LM state every this many damped solves

    def rebuild(self):
        """Re-run the build (state initialisation + preintegration) from the uploaded inputs: a fresh solve."""
        for p in self.parts:
            p.bs.build(True)
        self.tries = 0

    def _phase(self, ph):
        for p in self.parts:
            p.phase(ph)

    def optimize(self, max_tries=None):
        if self.comm.native:
            p = self.parts[0]
            api._check(p._L.v2gt_chain_optimize(p.bs.h, int(max_tries or 0), 1 if self.use_graph else 0), "v2gt_chain_optimize")
            inf = p.bs.info(0)
            self.tries = inf.tries
            return inf
        c, t, P = self.comm, self.torch, self.parts
        self._phase(PH_COST0)
        c.all_gather(P, BUF_SUMS_SEND, BUF_SUMS_ALL, t)
        self._phase(PH_BEGIN)
        cap = max_tries or 100000
        while self.tries < cap:
            self._phase(PH_LINEARIZE)
            self._phase(PH_ELIMINATE)
            c.all_gather(P, BUF_PACK_SEND, BUF_PACK_ALL, t)  # exchange A: segment outputs + separator records + global block parts
            self._phase(PH_SOLVE)
            self._phase(PH_COST)
            c.all_gather(P, BUF_SUMS_SEND, BUF_SUMS_ALL, t)  # exchange B: cost scalars + boundary step
            self._phase(PH_DECIDE)
            self.tries += 1
            if self.tries % self.check_every == 0 or self.tries >= cap:
                if P[0].bs.info(0).lm_state != 0:  # replicated LM state: identical on every rank
                    break
        return P[0].bs.info(0)

    def check_brackets(self):
        """True when the estimated time offset left every local state of this process inside its rank's slice of vicon
        poses (a state whose query time t - t_off leaves the slice would silently lose its vicon factor)."""
        toff = self.calibration()[2]
        gap = self.params.interp_gap_factor
        for p in self.parts:
            v0, v1, at_start, at_end = p.vic_range
            t0, t1 = p.t_range
            if (not at_start and t0 - toff - gap < v0) or (not at_end and t1 - toff + gap > v1):
                return False
        return True

    def info(self):
        return self.parts[0].bs.info(0)

    def own_states(self):
        """(global index range, times, states[n][16]) of every rank held by this process."""
        out = []
        for p in self.parts:
            tl, x = p.bs.states(0)
            lo, hi = p.plan.own_lo, p.plan.own_hi
            out.append((p.plan.own_global, tl[lo:hi], x[lo:hi]))
        return out

    def states(self):
        """All states of the chain on every process (LocalComm: concatenation; NativeComm: all_gather_object)."""
        mine = self.own_states()
        if self.comm.native and self.comm.world > 1:
            gathered = [None] * self.comm.world
            self.comm.dist.all_gather_object(gathered, mine)
            mine = [x for g in gathered for x in g]
        mine.sort(key=lambda e: e[0][0])
        return np.concatenate([e[1] for e in mine]), np.concatenate([e[2] for e in mine])

    def calibration(self):
        return self.parts[0].bs.calibration(0)

    def timing(self):
        return self.parts[0].bs.timing()

    def close(self):
        for p in self.parts:
            p.bs.close()
