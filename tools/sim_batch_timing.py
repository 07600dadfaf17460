"""Timing of the batched device simulator at the benchmark's shape (128 sim.launch-shaped trials per GPU) next to the
host simulator, and of the device-to-device staging into the batch solver.  Prints one JSON line.
usage: python tools/sim_batch_timing.py [--trials 128] [--solve]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api, sim  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trials", type=int, default=128)
    ap.add_argument("--solve", action="store_true")
    a = ap.parse_args()
    rows = sim.make_trajectory(**sim.TRAJ_SHAPES["tum_corridor1"])
    T = a.trials
    out = {"trials": T}
    t0 = time.perf_counter()
    h = sim.simulate(traj_rows=rows, seed=3)
    out["host_simulator_s_per_trial_1_thread"] = time.perf_counter() - t0
    sb = sim.SimBatch(rows, T)
    t0 = time.perf_counter()
    sb.plan()
    out["plan_host_s"] = time.perf_counter() - t0
    for rep in range(2):
        t0 = time.perf_counter()
        sb.run()
        out["run_wall_s" if rep else "run_wall_s_first"] = time.perf_counter() - t0
    out["generate_device_ms"] = sb.generate_ms()
    out["n_imu"], out["n_vicon"], out["n_states"] = sb.sizes(3)
    out["trials_per_s_device"] = T / (out["plan_host_s"] + out["run_wall_s"])
    d = sb.dataset(3, with_truth_states=True)
    out["max_abs_diff_vs_host"] = {k: float(np.abs(getattr(d, k) - getattr(h, k)).max()) for k in ("imu_t", "imu_w", "imu_a", "vic_t", "vic_q", "vic_p", "state_t")}
    out["max_abs_diff_vs_host"]["truth_states"] = float(np.abs(d.truth["states"] - h.truth["states"]).max())
    if a.solve:
        p = api.default_params()
        bs = api.BatchSolver(T)
        t0 = time.perf_counter()
        for k in range(T):
            bs.set_trial_sim(k, p, sb)
        bs.upload()
        out["stage_upload_from_device_s"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        bs.build_and_solve()
        out["build_and_solve_s"] = time.perf_counter() - t0
        infos = [bs.info(k) for k in range(T)]
        out["trials_ok"] = sum(1 for i in infos if i.status == 0)
        err = [abs(bs.calibration(k)[2] - sb.truth(k)["viconimu_dt"]) for k in range(T)]
        out["toff_abs_error_max"] = float(max(err))
        bs.close()
    sb.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
