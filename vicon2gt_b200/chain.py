"""Chain-sharded build-and-solve of ONE long sequence across GPUs (BASELINE.json config 5).

The cleaned state chain is cut into `world x segs_per_rank` segments with the single-GPU solver's geometry; rank r
holds the states of its segments (+ one halo state on each inner side), linearises and eliminates them locally, and
per damped solve exchanges only
    * the Schur outputs of its segments, the normal-equation records of its separators and its part of the 9x9 global
      block / gradient, packed into ONE all-gather (the global block is then summed in rank order on every rank),
    * one boundary step (all-gather of 16 doubles) and a handful of cost scalars (all-reduce),
after which every rank solves the small reduced system (separators + 9 globals) redundantly and takes the identical
Levenberg-Marquardt decision.  Collectives go through `torch.distributed` (NCCL over NVLink on the GPUs; gloo in the
CPU tests of the host logic) on the solver's own CUDA stream; `LocalComm` runs all ranks of a small world in one
process on one GPU (same kernels, same exchange, collectives as tensor copies) -- used to test the sharded algorithm
without a multi-GPU box.  Everything numeric happens in libvicon2gt_b200.so.
"""
import ctypes as C
import math
from dataclasses import dataclass

import numpy as np

from . import api

(BUF_SUMS, BUF_GAMMA, BUF_PACK_SEND, BUF_SEG_ALL, BUF_PACK_ALL, BUF_SEPNE_ALL, BUF_HALO_SEND, BUF_HALO_RECV) = range(8)
(PH_COST0, PH_BEGIN, PH_LINEARIZE, PH_ELIMINATE, PH_SOLVE, PH_COST, PH_DECIDE) = range(7)


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


def seg_stride(n, p):
    """Segment stride of the solver's geometry (csrc/k_solve.cu seg_geom): separators at m * stride - 1."""
    return max(2, (n + 1 + p - 1) // p)


def chain_plan(n_total, world, segs_per_rank):
    """Ownership of a chain of n_total states: rank r owns segments [r*spr, (r+1)*spr) and the separators in front of
    them, i.e. the contiguous states [j0*stride - 1, j1*stride - 1) (rank 0 starts at 0, the last rank ends at n)."""
    p = world * segs_per_rank
    stride = seg_stride(n_total, p)
    if (p - 1) * stride >= n_total or stride < 3:
        raise ValueError(f"chain of {n_total} states is too short for {p} segments")
    plans = []
    for r in range(world):
        j0, j1 = r * segs_per_rank, (r + 1) * segs_per_rank
        start = j0 * stride - 1 if r > 0 else 0
        end = j1 * stride - 1 if r < world - 1 else n_total
        g_off = start - (1 if r > 0 else 0)                 # left halo: the previous rank's last own state
        n_local = end + (1 if r < world - 1 else 0) - g_off  # right halo: the next rank's leading separator
        plans.append(RankPlan(r, world, n_total, p, stride, j0, j1, g_off, n_local, start - g_off, end - g_off))
    return plans


def fit_segments_per_rank(n_total, world, target):
    """Largest segments-per-rank <= target whose geometry uses every segment (ceil rounding of the stride can leave the
    last segments of a finer cut empty)."""
    for spr in range(max(1, int(target)), 0, -1):
        try:
            chain_plan(n_total, world, spr)
            return spr
        except ValueError:
            continue
    raise ValueError(f"chain of {n_total} states cannot be cut for {world} ranks")


def clean_timestamps(params, imu_t, vic_t, state_t):
    """ViconGraphSolver::build_and_solve timestamp cleaning (ViconGraphSolver.cpp:114-142) on the whole sequence."""
    ts = np.asarray(state_t, dtype=np.float64)
    first = vic_t.min() + params.toff_imu_to_vicon + 2 * params.time_offset_window
    last = vic_t.max() + params.toff_imu_to_vicon - 2 * params.time_offset_window
    keep = (ts >= imu_t[0]) & (ts <= imu_t[-1]) & (ts >= first) & (ts <= last)
    return ts[keep]


def slice_inputs(ds, ts_local, vicon_margin=1.0, imu_margin=0.05):
    """Inputs a rank needs for its local states: vicon poses +-(0.25 s window + margin), IMU samples bracketing the range."""
    t0, t1 = ts_local[0], ts_local[-1]
    vi0 = max(0, np.searchsorted(ds.vic_t, t0 - vicon_margin, "left") - 1)
    vi1 = min(len(ds.vic_t), np.searchsorted(ds.vic_t, t1 + vicon_margin, "right") + 1)
    ii0 = max(0, np.searchsorted(ds.imu_t, t0 - imu_margin, "left") - 2)
    ii1 = min(len(ds.imu_t), np.searchsorted(ds.imu_t, t1 + imu_margin, "right") + 2)
    return slice(ii0, ii1), slice(vi0, vi1)


class _DevBuf:
    """A device buffer of the library as a __cuda_array_interface__ object (wrapped by torch.as_tensor, no copy)."""

    def __init__(self, ptr, n):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": "<f8", "data": (int(ptr), False), "version": 2}


class _Part:
    """One rank's share: a 1-trial batch in chain mode + torch views of its exchange buffers."""

    def __init__(self, params, ds, ts_clean, plan, device, torch):
        self.plan = plan
        L = api.load_library()
        L.v2gt_chain_configure.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, C.c_int64, C.c_int32, C.c_int32, C.c_int32,
                                           C.c_int32, C.c_int32]
        L.v2gt_chain_phase.argtypes = [C.c_void_p, C.c_int32]
        L.v2gt_chain_buffer.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.POINTER(C.c_int64)]
        L.v2gt_batch_stream.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
        self._L = L
        self.bs = api.BatchSolver(1, device=device)
        ts_local = ts_clean[plan.g_off: plan.g_off + plan.n_local]
        si, sv = slice_inputs(ds, ts_local)
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
    """All ranks of the world in this process (one GPU): collectives are tensor copies in rank order."""

    def __init__(self, world):
        self.world = world
        self.ranks = list(range(world))

    def barrier_streams(self, parts, torch):
        for p in parts:
            p.bs.sync()

    def all_reduce(self, parts, which, torch):
        self.barrier_streams(parts, torch)
        total = parts[0].buf[which].clone()
        for p in parts[1:]:
            total += p.buf[which]
        for p in parts:
            p.buf[which].copy_(total)
        torch.cuda.synchronize()

    def all_gather(self, parts, send, recv, torch):
        self.barrier_streams(parts, torch)
        cat = torch.cat([p.buf[send] for p in parts])
        for p in parts:
            p.buf[recv][: cat.numel()].copy_(cat)
        torch.cuda.synchronize()


class DistComm:
    """One rank per process: torch.distributed collectives (NCCL) enqueued on the solver's own stream."""

    def __init__(self, dist):
        self.dist = dist
        self.world = dist.get_world_size()
        self.ranks = [dist.get_rank()]

    def all_reduce(self, parts, which, torch):
        p = parts[0]
        with torch.cuda.stream(p.stream):
            self.dist.all_reduce(p.buf[which])

    def all_gather(self, parts, send, recv, torch):
        p = parts[0]
        n = p.buf[send].numel() * self.world
        with torch.cuda.stream(p.stream):
            self.dist.all_gather_into_tensor(p.buf[recv][:n], p.buf[send])


class ChainSolver:
    """build_and_solve of one long sequence, sharded over `comm.world` ranks."""

    def __init__(self, params, ds, comm, segs_per_rank=16, device=0):
        import torch

        self.torch = torch
        self.comm = comm
        self.params = params
        self.ts = clean_timestamps(params, ds.imu_t, ds.vic_t, ds.state_t)
        self.segs_per_rank = fit_segments_per_rank(len(self.ts), comm.world, segs_per_rank)
        self.plans = chain_plan(len(self.ts), comm.world, self.segs_per_rank)
        self.parts = [_Part(params, ds, self.ts, self.plans[r], device, torch) for r in comm.ranks]
        self.tries = 0
        self.check_every = 4  # host reads the replicated LM state every this many damped solves (phases of a finished
        #                       chain are no-ops on the device, exactly as in the batch path)

    def rebuild(self):
        """Re-run the build (state initialisation + preintegration) from the uploaded inputs: a fresh solve."""
        for p in self.parts:
            p.bs.build(True)
        self.tries = 0

    def _phase(self, ph):
        for p in self.parts:
            p.phase(ph)

    def optimize(self, max_tries=None):
        c, t, P = self.comm, self.torch, self.parts
        self._phase(PH_COST0)
        c.all_reduce(P, BUF_SUMS, t)
        self._phase(PH_BEGIN)
        cap = max_tries or 100000
        while self.tries < cap:
            self._phase(PH_LINEARIZE)
            self._phase(PH_ELIMINATE)
            c.all_gather(P, BUF_PACK_SEND, BUF_PACK_ALL, t)  # segment outputs + separator records + global block parts
            self._phase(PH_SOLVE)
            c.all_gather(P, BUF_HALO_SEND, BUF_HALO_RECV, t)
            self._phase(PH_COST)
            c.all_reduce(P, BUF_SUMS, t)
            self._phase(PH_DECIDE)
            self.tries += 1
            if self.tries % self.check_every == 0 or self.tries >= cap:
                if P[0].bs.info(0).lm_state != 0:  # replicated LM state: identical on every rank
                    break
        return P[0].bs.info(0)

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
        """All states of the chain on every process (LocalComm: concatenation; DistComm: all_gather_object)."""
        mine = self.own_states()
        if isinstance(self.comm, DistComm):
            gathered = [None] * self.comm.world
            self.comm.dist.all_gather_object(gathered, mine)
            mine = [x for g in gathered for x in g]
        mine.sort(key=lambda e: e[0][0])
        return np.concatenate([e[1] for e in mine]), np.concatenate([e[2] for e in mine])

    def calibration(self):
        return self.parts[0].bs.calibration(0)

    def close(self):
        for p in self.parts:
            p.bs.close()
