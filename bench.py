#!/usr/bin/env python
"""bench.py -- solved trajectories / second of the vicon2gt build-and-solve hot path on B200.

Workload (BASELINE.json configs[3], the one the metric is quoted on): a Monte-Carlo batch of independent
simulated trials of the `launch/sim.launch` shape (400 Hz IMU, 100 Hz vicon, 100 Hz states, ~29.8 k states,
~119.6 k IMU samples, ~29.9 k vicon poses per trial), sharded one batch per GPU with no communication
(weak scaling: --trials trials per GPU).  One "step" = the full build-and-solve of the rank's batch
(state initialisation, continuous preintegration, 30-iteration Levenberg-Marquardt).

  value  trajectories/s with the inputs already resident in HBM (CUDA events around build + optimise, max over ranks)
  e2e    the same through the public API with HOST buffers, pipelined (api.WaveRunner, two resident handles): per step
         the staging + pinned H2D copy of every input array, build + optimise and the D2H read of all states +
         calibrations, wall clock over all steps; e2e.serial = one blocking step at a time
  roofline  dominant kernel (k_seg_eliminate): algorithmic bytes / summed launch durations (CUDA events on the
         handle's stream, inside the timed steps) against MEASURED_PEAKS.json's HBM number; FP64 fraction
         against a DFMA peak measured in the same process
  per_rank  every rank's device time, damped solves and eliminated state blocks per step
  chain  BASELINE.json configs[4] on the same GPUs: one one-hour sequence chain-sharded over the ranks, the
         Levenberg-Marquardt loop inside the library (2 ncclAllGather per damped solve, CUDA graph); ms per damped solve
  waves_1024_on_one_gpu  (N = 1) the WHOLE 1024-trial batch on the one GPU in 8 double-buffered waves
  cpu_baseline  the oracle (CPU restatement of the reference; the reference itself cannot be built here:
         no Eigen / Boost / GTSAM / ROS) on a bounded sample of the same workload

`--impl reference` times that CPU restatement on all host threads, on this arm's workload and `config` (kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "solved_trajectories_per_second"
UNIT = "trajectories/s"
# algorithmic work per state and damped solve (stated in DESIGN.md)
ELIM_ALG_BYTES_PER_STATE = 2 * (120 + 225 + 135 + 15) * 8  # read D,O,B,g once + write L,W_O,W_B,w_g once = 7920 B
ELIM_ALG_FLOPS_PER_STATE = 19.4e3                           # block-tridiagonal-arrow Cholesky, Thomas-equivalent


def ncu_traffic_per_state_block():
    """DRAM bytes (read + write) per eliminated state block of k_seg_eliminate, from the committed ncu --set full capture
    (profiles/elim_traffic.json: dram__bytes_read.sum + dram__bytes_write.sum of one launch / its state blocks)."""
    path = os.path.join(ROOT, "profiles", "elim_traffic.json")
    if not os.path.exists(path):
        return None
    with open(path) as f:
        return float(json.load(f)["dram_bytes_per_state_block"])


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trials", type=int, default=int(os.environ.get("V2GT_BENCH_TRIALS", "128")),
                    help="trials per GPU (weak scaling; 128 per GPU = the 1024-trial batch on 8 GPUs)")
    ap.add_argument("--duration", type=float, default=0.0, help="truncate the trajectory to this many seconds (0 = full)")
    ap.add_argument("--segments", type=int, default=0)
    ap.add_argument("--target-warps", type=int, default=0, help="segment warps per elimination launch (0 = library default)")
    ap.add_argument("--cpu-sample-s", type=float, default=40.0, help="length of the CPU baseline sample trajectory")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-budget-s", type=float, default=300.0, help="--impl reference: wall-clock budget of all steps together")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--no-sim", action="store_true", help="skip the device-simulator leg (SURVEY 8f row 1)")
    ap.add_argument("--workload", default="mc", choices=["mc", "chain"],
                    help="mc: Monte-Carlo batch (BASELINE configs[3], the default the metric is quoted on); chain: ONE one-hour "
                         "sequence chain-sharded over the GPUs with NCCL (configs[4]; second half of the metric: LM latency)")
    ap.add_argument("--segs-per-rank", type=int, default=0, help="chain workload: segments per GPU (0 = ~sqrt(states per GPU))")
    ap.add_argument("--chain-duration", type=float, default=0.0, help="chain workload: truncate the hour to this many seconds")
    ap.add_argument("--waves-trials", type=int, default=1024,
                    help="N = 1 only: solve this many trials (the whole configs[3] batch) on the one GPU in waves of --trials, staging "
                         "of wave k+1 and read-back of wave k-1 overlapped with wave k's solve (0 = skip)")
    ap.add_argument("--no-pipeline", action="store_true", help="kernel experiments: report the serial end-to-end number only")
    ap.add_argument("--no-chain", action="store_true", help="skip the chain-sharded leg (configs[4]) of the default line")
    ap.add_argument("--chain-graph-only", action="store_true", help="chain leg: skip the plain-launch pass (per-phase times)")
    ap.add_argument("--chain-steps", type=int, default=3)
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def make_trials(n, seed0, duration):
    """n distinct simulated trials (seeds seed0..), generated on host threads (the C++ simulator releases the GIL)."""
    from vicon2gt_b200 import sim

    out = [None] * n
    kw = {} if duration <= 0 else dict(duration_s=duration)

    def work(i):
        out[i] = sim.make_config("C1", seed=seed0 + i, **kw)

    nthr = max(1, min(n, os.cpu_count() or 1))
    idx = list(range(n))
    threads = []
    for k in range(nthr):
        th = threading.Thread(target=lambda k=k: [work(i) for i in idx[k::nthr]])
        th.start()
        threads.append(th)
    for th in threads:
        th.join()
    return out


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def simulator_leg(api, sim, bs, p, T, seed0, duration, read_back):
    """the same trials generated by the batched device simulator (csrc/k_sim.cu) and handed to the solver
    device -> device, next to the host simulator timed on one trial.  One step of simulate -> stage -> solve -> read back."""
    rows, _ = sim.trajectory_rows("tum_corridor1", duration if duration > 0 else None)
    t0 = time.perf_counter()
    sim.simulate(traj_rows=rows, seed=seed0)
    host_s = time.perf_counter() - t0
    sb = sim.SimBatch(rows, T, seeds=range(seed0, seed0 + T))
    try:
        sb.plan()
        sb.run()  # first run allocates
        t0 = time.perf_counter()
        sb.plan()
        t1 = time.perf_counter()
        sb.run()
        t2 = time.perf_counter()
        for i in range(T):
            bs.set_trial_sim(i, p, sb)
        bs.upload()
        t3 = time.perf_counter()
        bs.build_and_solve()
        t4 = time.perf_counter()
        read_back()
        t5 = time.perf_counter()
        ok = sum(1 for i in range(T) if bs.info(i).status == 0)
        M, V, N = sb.sizes(0)
        return {"trials": T, "n_imu": M, "n_vicon": V, "n_states_raw": N, "plan_host_ms": 1e3 * (t1 - t0), "generate_wall_ms": 1e3 * (t2 - t1),
                "generate_device_ms": sb.generate_ms(), "simulated_trials_per_s": T / (t2 - t0),
                "host_simulator_s_per_trial_1_thread": host_s, "stage_device_to_device_ms": 1e3 * (t3 - t2),
                "build_and_solve_ms": 1e3 * (t4 - t3), "read_back_ms": 1e3 * (t5 - t4), "trials_ok": ok,
                "simulate_to_result_trajectories_per_s": T / (t5 - t0), "kernels_launched": 6}
    finally:
        sb.close()


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            d = json.load(f)
        return float(d.get("hbm_gbs", 6650.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


FULL_STATES = 29798.0  # states of a full sim.launch trial after cleaning


def workload_config(trials_per_gpu, world, trajectory):
    """`config` of the JSON line: the same dict in both arms (the reference arm runs this workload on the host cores)."""
    return {
        "workload": ("MC batch of independent sim.launch-shaped trials (BASELINE.json configs[3]): 400 Hz IMU, 100 Hz vicon, "
                     "100 Hz states, 30-iteration LM; %d trials per GPU, %d GPU(s), no communication" % (trials_per_gpu, world)),
        "trials_per_gpu": trials_per_gpu, "n_gpus": world, "trajectory": trajectory,
        "l2": "no flush: every trial's per-step working set (0.37 GB of blocks + factors) exceeds the 126 MB L2",
    }


def cpu_baseline(sample_s, threads, faithful=False, seed0=1000):
    """Oracle (CPU restatement of the reference) on a bounded sample: one sim.launch-shaped trial per host thread, the whole
    trajectory (sample_s <= 0) or truncated to `sample_s` seconds.  Returns equivalent full-size trajectories / s.
    faithful: the reference's own O(N M) linear scans in Propagator::propagate / has_bounding_imu (Propagator.cpp:46, :165-193)
    instead of the port's binary searches."""
    from oracle import binding as ob
    from vicon2gt_b200 import sim
    from vicon2gt_b200._cdefs import TrialInfo

    kw = {} if sample_s <= 0 else dict(duration_s=sample_s)
    dss = [None] * threads

    def gen(i):
        dss[i] = sim.make_config("C1", seed=seed0 + i, **kw)

    ths = [threading.Thread(target=gen, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    p = ob.default_params()
    solvers = [ob.OracleSolver.from_dataset(p, ds) for ds in dss]
    if faithful:
        for sv in solvers:
            sv.set_option("faithful_scan", 1)
    res = [None] * threads

    def work(i):
        st = solvers[i].build_and_solve()
        inf = solvers[i].info(TrialInfo)
        res[i] = (st, inf.n_states, inf.tries, inf.iterations)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    dt = time.perf_counter() - t0
    n_states = sum(r[1] for r in res)
    full = sample_s <= 0
    value = (threads if full else n_states / FULL_STATES) / dt
    what = "whole trajectory" if full else f"truncated to {sample_s:.0f} s"
    sample = (f"{threads} x sim.launch-shaped trial ({dss[0].truth.get('trajectory', '?')}), {what} ({res[0][1]} states, {res[0][2]} damped "
              f"solves, {res[0][3]} LM iterations each), full build_and_solve, {dt:.1f} s wall"
              + ("" if full else f"; scaled by states to the {int(FULL_STATES)}-state trial")
              + ("; reference-faithful O(N M) selection scans" if faithful else ""))
    for sv in solvers:
        sv.close()
    return value, sample, dt, n_states / threads


def run_reference(args):
    """The reference arm: the CPU restatement of the reference's path (oracle/, kind "port": GTSAM / Eigen / ROS cannot be
    installed here, and oracle/_ref's stand-in matrix class would flatter the ratio) on all host threads, on THIS arm's
    workload: every step solves one sim.launch-shaped trial per host thread.  Full-size trials (same config as the GPU arm)
    when --steps + --warmup of them fit the time budget (~45 s each), else the longest truncation that does, stated in
    `cpu_baseline.sample` and `reference_sample`."""
    rank, _, world = dist_env()
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n_steps = args.warmup + args.steps
    budget_s = args.ref_budget_s
    est_full_s = 48.0  # one full-size trial per thread, refined after every step
    vals, dts, fulls = [], [], []
    sample = ""
    t_start = time.perf_counter()
    traj = "?"
    for k in range(n_steps):
        left = max(1.0, budget_s - (time.perf_counter() - t_start))
        per_step = left / (n_steps - k)
        frac = per_step / est_full_s
        sample_s = 0.0 if frac >= 1.0 else max(8.0, 0.9 * frac * 299.0)
        v, sample, dt, n_st = cpu_baseline(sample_s, threads, seed0=1000 + k * threads)
        est_full_s = dt * FULL_STATES / max(1.0, n_st)
        if k >= args.warmup:
            vals.append(v)
            dts.append(dt)
            fulls.append(sample_s <= 0.0)
    value = float(np.mean(vals))
    from vicon2gt_b200 import sim

    traj = "file:" + sim.TRAJ_FILES["tum_corridor1"] if sim.trajectory_file("tum_corridor1") else "synthesised, tum_corridor1 shape"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(dts)), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args.trials, args.gpus, traj),
        "reference_sample": {"full_size_steps": int(sum(fulls)), "timed_steps": len(fulls), "time_budget_s": budget_s,
                             "note": "each step = one trial per host thread; full-size trials when the requested number of steps "
                                     "fits the time budget, else truncated and scaled by states (an approximation: the LM iteration "
                                     "and damped-solve counts of the sample are in cpu_baseline.sample)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def chain_leg(args, dist, device, rank, world, steps, warmup):
    """BASELINE.json configs[4]: ONE ~72 k-state sequence, contiguous chain segments per GPU.  The Levenberg-Marquardt loop
    runs inside the library (v2gt_chain_optimize): per damped solve two ncclAllGather calls on the solver's stream (segment
    Schur outputs + separator records + global-block parts; cost scalars + boundary step), the kernel + collective sequence
    of a damped solve replayed as one CUDA graph.  Returns the result dict on rank 0 (None elsewhere)."""
    import torch

    from vicon2gt_b200 import api, sim
    from vicon2gt_b200.chain import ChainSolver, NativeComm

    kw = {} if args.chain_duration <= 0 else dict(duration_s=args.chain_duration)
    ds = sim.make_config("C5", seed=5, **kw)  # identical on every rank (deterministic generator)
    p = api.default_params()
    comm = NativeComm(dist)
    # ~4 sqrt(n) segments over the whole chain (the separator chain is eliminated in segments of its own, nested level 2),
    # but never fewer than 16 per GPU: the per-rank sequential depth is what more GPUs are there to cut
    spr = args.segs_per_rank if args.segs_per_rank > 0 else max(16, int(round(4.0 * ds.n_times ** 0.5 / world)))
    out = None
    results = {}
    for mode in (["graph", "launches"] if not args.chain_graph_only else ["graph"]):
        cs = ChainSolver(p, ds, comm, segs_per_rank=spr, device=device, use_graph=(mode == "graph"))

        def barrier():
            if dist is not None:
                dist.barrier()

        def step():
            cs.rebuild()
            inf = cs.optimize()
            m, _ = cs.timing()
            return float(m[0]), float(m[7]), m, inf

        for _ in range(warmup):
            step()
        sampler = ClockSampler(device)
        barrier()
        torch.cuda.synchronize(device)
        sampler.start()
        b_ms = o_ms = 0.0
        m0, _ = cs.timing()
        for _ in range(steps):
            b, o, m1, inf = step()
            b_ms += b
            o_ms += o
        torch.cuda.synchronize(device)
        barrier()
        clocks = sampler.stop()
        phases = (m1 - m0)
        tot = torch.tensor([b_ms + o_ms, o_ms], dtype=torch.float64, device=f"cuda:{device}")
        if dist is not None:
            dist.all_reduce(tot, op=dist.ReduceOp.MAX)
        tot_ms, opt_ms = (float(v) for v in tot.tolist())
        results[mode] = {
            "ms_per_damped_solve": opt_ms / steps / max(1, inf.tries), "ms_per_lm_iteration": opt_ms / steps / max(1, inf.iterations),
            "build_ms": (tot_ms - opt_ms) / steps, "optimize_ms": opt_ms / steps, "iterations": inf.iterations,
            "damped_solves": inf.tries, "final_error": inf.final_error, "status": inf.status, "clocks": clocks,
        }
        if mode == "launches":  # per-phase CUDA events exist only with plain launches (rank 0's own times)
            names = ["linearize", "eliminate_pack", "exchangeA_reduced", "backsub", "cost", "exchangeB_decide"]
            results[mode]["phase_ms_per_damped_solve_rank0"] = {k: float(phases[1 + i]) / steps / max(1, inf.tries)
                                                                 for i, k in enumerate(names)}
        n_states, segs = len(cs.ts), cs.segs_per_rank
        brackets_ok = cs.check_brackets()
        cs.close()
    if rank == 0:
        g = results["graph"]
        out = {"workload": "ONE synthesised one-hour sequence (BASELINE.json configs[4]): 20 Hz states, 200 Hz IMU, 100 Hz vicon, "
                           "chain-sharded over %d GPU(s)" % world,
               "n_gpus": world, "n_states": n_states, "n_imu": ds.n_imu, "n_vicon": ds.n_vicon, "segments_per_gpu": segs,
               "collectives_per_damped_solve": 0 if world == 1 else 2,
               "collectives": "ncclAllGather x2 inside the library, on the solver stream, captured in the CUDA graph of a damped solve",
               "ms_per_damped_solve": g["ms_per_damped_solve"], "ms_per_lm_iteration": g["ms_per_lm_iteration"],
               "trajectories_per_second": 1e3 / (g["build_ms"] + g["optimize_ms"]), "brackets_inside_rank_slices": brackets_ok,
               "timing": "CUDA events on the solver stream around build + optimise, max over ranks; %d steps after %d warm-up" % (steps, warmup),
               "graph": g, "plain_launches": results.get("launches")}
    return out


def run_chain(args):
    """`--workload chain`: the chain leg alone, as its own JSON line (metric: latency of one damped solve)."""
    import torch

    rank, local_rank, world = dist_env()
    dist = None
    if world > 1:
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0
    res = chain_leg(args, dist, device, rank, world, args.steps, args.warmup)
    if rank == 0:
        line = {"metric": "lm_damped_solve_latency", "value": res["ms_per_damped_solve"], "unit": "ms", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": res["graph"]["build_ms"] + res["graph"]["optimize_ms"],
                "higher_is_better": False, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": res["workload"], "n_states": res["n_states"], "segments_per_gpu": res["segments_per_gpu"]},
                "chain": res, "clocks": res["graph"]["clocks"]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.workload == "chain":
        return run_chain(args)
    rank, local_rank, world = dist_env()
    from vicon2gt_b200 import api, build

    if not os.path.exists(build.lib_path("cuda")) or not os.path.exists(build.lib_path("sim")):
        build.build_sim()
        build.build_cuda()
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(local_rank)
        dist.init_process_group(backend="nccl", device_id=torch.device("cuda", local_rank))
    device = local_rank if world > 1 else 0
    T = args.trials
    p = api.default_params()
    trials = make_trials(T, seed0=rank * T, duration=args.duration)
    h2d = int(sum(ds.nbytes() for ds in trials))

    def barrier():
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        import torch

        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    opts = {}
    if args.segments > 0:
        opts["segments"] = args.segments
    if args.target_warps > 0:
        opts["target_warps"] = args.target_warps
    bs = api.BatchSolver(T, device=device, **opts)

    def stage_and_upload():
        for i, ds in enumerate(trials):
            bs.set_trial_dataset(i, p, ds)
        bs.upload()

    def read_back():
        # every trial's states + calibration in one call (v2gt_batch_get_all_states: pinned bounce buffer, chunked)
        _, t, x, g = bs.all_states()
        return t.nbytes + x.nbytes + g.nbytes

    # ---------------- device-resident steps: inputs already in HBM
    stage_and_upload()
    for _ in range(args.warmup):
        bs.build(True)
        bs.optimize()
    sampler = ClockSampler(device)
    barrier()
    bs.sync()
    sampler.start()
    ms0, l0 = bs.timing()
    dev_ms = build_ms = opt_ms = 0.0
    t_wall0 = time.perf_counter()
    for _ in range(args.steps):
        bs.build(True)
        bs.optimize()
        m_after, _ = bs.timing()
        build_ms += m_after[0]  # CUDA events on the handle's stream
        opt_ms += m_after[7]
    dev_ms = build_ms + opt_ms
    bs.sync()
    barrier()
    wall_s = time.perf_counter() - t_wall0
    clocks = sampler.stop()
    ms1, l1 = bs.timing()
    dev_s = max_over_ranks(dev_ms * 1e-3)
    value = world * T * args.steps / dev_s

    # per-trial work counters (for the roofline accounting)
    infos = [bs.info(i) for i in range(T)]
    ok = sum(1 for inf in infos if inf.status == 0)
    solve_units = float(sum(inf.tries * inf.n_states for inf in infos))  # state-blocks eliminated per step
    phase_ms = (ms1 - ms0)
    phase_ms[0], phase_ms[7] = build_ms, opt_ms  # these two slots hold the last call only
    launches = int((l1 - l0).sum())
    elim_ms = float(phase_ms[2])
    hbm_peak, peak_src = measured_peaks()
    fp64_peak = api.measure_fp64_peak(device)
    elim_bytes = solve_units * args.steps * ELIM_ALG_BYTES_PER_STATE
    elim_flops = solve_units * args.steps * ELIM_ALG_FLOPS_PER_STATE
    elim_launches = max(1, int((l1 - l0)[2]))
    tpsb = ncu_traffic_per_state_block()
    traffic = None if tpsb is None else tpsb * solve_units * args.steps / elim_launches
    achieved_gbs = elim_bytes / (elim_ms * 1e-3) / 1e9 if elim_ms > 0 else 0.0
    achieved_tf = elim_flops / (elim_ms * 1e-3) / 1e12 if elim_ms > 0 else 0.0

    # ---------------- end to end through the public API with host buffers
    e2e_s = 0.0
    d2h = 0
    barrier()
    e2e_parts = np.zeros(4)
    for _ in range(args.steps):
        t0 = time.perf_counter()
        for i, ds in enumerate(trials):
            bs.set_trial_dataset(i, p, ds, borrow=True)  # views of the host arrays; staged + copied by upload()
        t1 = time.perf_counter()
        bs.upload()
        t2 = time.perf_counter()
        bs.build_and_solve()
        t3 = time.perf_counter()
        d2h = read_back()
        t4 = time.perf_counter()
        e2e_s += t4 - t0
        e2e_parts += np.array([t1 - t0, t2 - t1, t3 - t2, t4 - t3])
    e2e_serial_own = e2e_s
    e2e_serial_value = world * T * args.steps / max_over_ranks(e2e_s)

    # ---------------- the same trials from the device simulator (inputs never on the host); rank 0 only
    simulator = None
    if not args.no_sim and rank == 0:
        try:
            from vicon2gt_b200 import sim as _sim

            simulator = simulator_leg(api, _sim, bs, p, T, rank * T, args.duration, read_back)
        except Exception as e:  # the leg is a widened row: it must never cost the headline line
            simulator = {"error": repr(e)}
    barrier()

    # ---------------- LM iteration latency on a single trajectory (second half of the metric)
    latency = None
    if not args.no_latency and rank == 0:
        one = api.BatchSolver(1, device=device)
        one.set_trial_dataset(0, p, trials[0])
        one.upload()
        one.build(True)
        one.optimize()  # warm
        one.build(True)
        one.optimize()
        m, _ = one.timing()
        inf = one.info(0)
        latency = {"config": "single sim.launch-shaped trajectory", "n_states": inf.n_states, "lm_iterations": inf.iterations,
                   "damped_solves": inf.tries, "build_ms": float(m[0]), "optimize_ms": float(m[7]),
                   "ms_per_lm_iteration": float(m[7]) / max(1, inf.iterations), "ms_per_damped_solve": float(m[7]) / max(1, inf.tries)}
        one.close()

    # ---------------- BASELINE.json configs[4]: one one-hour sequence chain-sharded over the same GPUs (all ranks take part)
    chain = None
    if not args.no_chain:
        bs.close()
        bs = None
        try:
            chain = chain_leg(args, dist, device, rank, world, args.chain_steps, 2)
        except Exception as e:  # a widened row must never cost the headline line
            chain = {"error": repr(e)}
        barrier()

    # ---------------- end to end, pipelined (the headline): the public WaveRunner API, HOST buffers, `steps` waves of the rank's
    # trials; every wave's staging + H2D copies and its D2H read-back are inside the timed region, overlapped with the solve
    # of the neighbouring waves (two resident handles).  Needs the memory of two batches: the resident handle goes first.
    if bs is not None:
        bs.close()
        bs = None
    barrier()
    e2e_own, e2e_pipe = e2e_serial_own, None
    with api.WaveRunner(wave=T, device=device, **opts) as runner:
        if args.no_pipeline:
            runner = None
        if runner is not None:
            for _ in runner.run(p, trials * 2, read_back=False):  # allocations of the two handles (untimed)
                pass
            barrier()
            rows_back = 0
            for first, off, t_, x_, g_, infs in runner.run(p, trials * args.steps):
                rows_back += int(off[-1])
            e2e_own = runner.wall_s
            e2e_pipe = {"device_s": runner.device_ms * 1e-3, "wall_s": runner.wall_s, "upload_stall_s": runner.stall_s,
                        "states_read_back": rows_back}
    e2e_s = max_over_ranks(e2e_own)
    e2e_value = world * T * args.steps / e2e_s

    # ---------------- the WHOLE 1024-trial batch on one GPU, in waves (strong-scaling row: 1 GPU here, 8 GPUs = the N = 8 line)
    waves = None
    if world == 1 and args.waves_trials > T and args.duration <= 0:
        try:
            if bs is not None:
                bs.close()
                bs = None
            extra = make_trials(args.waves_trials - T, seed0=T, duration=args.duration)
            every = trials + extra
            with api.WaveRunner(wave=T, device=device, **opts) as runner:
                for _ in runner.run(p, every[:2 * T], read_back=False):  # allocations of the two handles (untimed)
                    pass
                n_ok = n_rows = 0
                for first, off, t_, x_, g_, infs in runner.run(p, every):
                    n_ok += sum(1 for inf in infs if inf.status == 0)
                    n_rows += int(off[-1])
                waves = {"trials": len(every), "trials_per_wave": T, "waves": (len(every) + T - 1) // T, "trials_ok": n_ok,
                         "states_read_back": n_rows, "device_s": runner.device_ms * 1e-3, "wall_s": runner.wall_s,
                         "value": len(every) / (runner.device_ms * 1e-3), "e2e": len(every) / runner.wall_s, "unit": UNIT,
                         "e2e_over_value": (len(every) / runner.wall_s) / (len(every) / (runner.device_ms * 1e-3)),
                         "upload_stall_s": runner.stall_s,
                         "h2d_bytes": int(sum(ds.nbytes() for ds in every)),
                         "note": "two resident batch handles; host threads stage + upload wave k+1 and read back wave k-1 while "
                                 "wave k is solved (api.WaveRunner / WaveRunner of include/vicon2gt_b200.hpp)"}
            del every, extra
        except Exception as e:  # a widened row must never cost the headline line
            waves = {"error": repr(e)}

    cpu = None
    if not args.no_cpu_baseline and rank == 0:
        cs_s = args.cpu_sample_s if args.duration <= 0 else min(args.cpu_sample_s, args.duration)
        v, sample, _, _ = cpu_baseline(cs_s, 1)
        # second stated baseline: the same sample with the reference's own O(N M) sample-selection scans
        vf, sample_f, _, _ = cpu_baseline(cs_s, 1, faithful=True)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
               "reference_faithful_scans": {"value": vf, "unit": UNIT, "cores": 1, "sample": sample_f,
                                            "note": "the scans cost O(states x IMU samples): at the full 29.8 k-state trial they are "
                                                    "~60x this sample's share (not scaled here)"}}

    total_states = sum_over_ranks(float(sum(inf.n_states for inf in infos)))
    # what limits the scaling curve: every rank's own device time, damped solves and state blocks (different seeds per rank)
    mine = [dev_ms / args.steps, float(sum(inf.tries for inf in infos)), solve_units, 1e3 * e2e_own / args.steps,
            float(max(inf.tries for inf in infos)), float(phase_ms[2]) / args.steps]
    per_rank = None
    if dist is None:
        allr = [mine]
    else:
        import torch

        tt = torch.tensor(mine, dtype=torch.float64, device=f"cuda:{local_rank}")
        got = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(got, tt)
        allr = [g.tolist() for g in got]
    per_rank = {k: [r[i] for r in allr] for i, k in enumerate(["device_ms_per_step", "damped_solves_all_trials", "state_blocks_per_step",
                                                               "e2e_ms_per_step", "max_damped_solves_of_a_trial", "eliminate_ms_per_step"])}
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * dev_s / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": workload_config(T, world, trials[0].truth.get("trajectory", "?") +
                                      ("" if args.duration <= 0 else f", truncated to {args.duration:.0f} s")),
            "run": {"states_total": int(total_states), "trials_ok": ok, "wall_s_timed_region": wall_s,
                    "timing": "CUDA events on the solver stream (build + optimise), max over ranks",
                    "working_set_gb_per_gpu": total_states / world * 11.0e3 / 1e9},
            "per_rank": per_rank,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(d2h),
                    "how": "api.WaveRunner over %d steps of the rank's %d trials from HOST arrays: per step the staging + pinned H2D copy of "
                           "every input array and the D2H read of all states + calibrations are inside the timed wall clock (max over "
                           "ranks), overlapped with the solve of the neighbouring steps (two resident handles)" % (args.steps, T),
                    "pipeline": e2e_pipe,
                    "serial": {"value": e2e_serial_value, "unit": UNIT,
                               "how": "one blocking step at a time on one handle: stage, upload, build_and_solve, read back",
                               "ms_per_step": {k: 1e3 * float(v) / args.steps for k, v in
                                               zip(["stage_host", "upload_h2d", "build_and_solve", "read_back_d2h"], e2e_parts)}}},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {
                "kernel": "k_seg_eliminate", "bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                "frac": achieved_gbs / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                "launches": elim_launches, "avg_launch_ms": elim_ms / elim_launches,
                "algorithmic_bytes_per_launch": elim_bytes / elim_launches,
                "share_of_step": elim_ms / max(1e-9, dev_ms),
                "fp64": {"achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / max(fp64_peak, 1e-9),
                         "peak_source": "measured DFMA loop (v2gt_measure_fp64_peak)"},
                "algorithmic_bytes_per_state_block": ELIM_ALG_BYTES_PER_STATE, "algorithmic_flops_per_state_block": ELIM_ALG_FLOPS_PER_STATE,
                # what the kernel actually issues on the FP64 pipe per state block: 90 mma.m8n8k4.f64 (512 flop each; W = L^-1 M
                # and the Schur products incl. the fill-in column of the partitioning) + ~400 DFMA-equivalents of the factorisation
                "executed_fp64": {"dmma_per_state_block": 90, "tflops": 90 * 512 * solve_units * args.steps / (elim_ms * 1e-3) / 1e12
                                  if elim_ms > 0 else 0.0,
                                  "frac_of_measured_peak": (90 * 512 * solve_units * args.steps / (elim_ms * 1e-3) / 1e12) / max(fp64_peak, 1e-9)
                                  if elim_ms > 0 else 0.0},
            },
            "phases_ms_per_step": {k: float(phase_ms[i]) / args.steps for i, k in
                                   enumerate(["build", "linearize", "eliminate", "reduced", "backsub", "cost", "lm", "optimize_total"])},
            "lm_iteration_latency": latency,
            "simulator": simulator,
            "chain": chain,
            "waves_1024_on_one_gpu": waves,
            "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    if bs is not None:
        bs.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
