#!/usr/bin/env python
"""The Monte-Carlo sweep of the reference (scripts/run_monte_carlo.sh: a 5 x 5 grid of vicon orientation / position noise,
N seeds per cell, one `roslaunch vicon2gt sim.launch` per run) as ONE batch on the GPU: simulate every run on host threads,
solve them together, compute the per-run error report on the device and print the table of the tech report (Table 1:
mean ATE in degrees / metres per noise cell).  Optionally writes every run's CSV / info / pose-text files.

    python examples/run_monte_carlo.py --seeds 10 [--shape C3] [--out-dir /tmp/mc] [--device-sim]

--device-sim generates the runs with the batched device simulator (csrc/k_sim.cu) instead of the host one: same seeds, same
streams, the measurements never leave the GPU.
"""
import argparse
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api, sim  # noqa: E402

NOISES = [0.001, 0.005, 0.01, 0.05, 0.10]  # scripts/run_monte_carlo.sh:8-14 (rad / m)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seeds", type=int, default=10)
    ap.add_argument("--shape", default="C3", help="C3: TUM corridor1 shape, 200 Hz IMU / 20 Hz states / 100 Hz vicon (the tech report's setup); C1: sim.launch")
    ap.add_argument("--duration", type=float, default=None)
    ap.add_argument("--out-dir", default=None)
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--device-sim", action="store_true")
    a = ap.parse_args()
    if a.device_sim:
        return main_device_sim(a)
    cells = [(h, i) for h in range(5) for i in range(5)]
    runs = [(h, i, j) for (h, i) in cells for j in range(a.seeds)]
    dss = [None] * len(runs)
    kw = {} if a.duration is None else dict(duration_s=a.duration)

    def work(k0, step):
        for k in range(k0, len(runs), step):
            h, i, j = runs[k]
            s = (NOISES[h],) * 3 + (NOISES[i],) * 3
            dss[k] = sim.make_config(a.shape, seed=j, vicon_sigmas=s, **kw)

    t0 = time.perf_counter()
    nthr = min(len(runs), os.cpu_count() or 1)
    ths = [threading.Thread(target=work, args=(k, nthr)) for k in range(nthr)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    t1 = time.perf_counter()
    p = api.default_params()
    bs = api.BatchSolver(len(runs), device=a.device)
    for k, ds in enumerate(dss):
        bs.set_trial_dataset(k, p, ds, borrow=True)
    bs.build_and_solve()
    t2 = time.perf_counter()
    ok = np.array([bs.info(k).status == 0 for k in range(len(runs))])
    for k, ds in enumerate(dss):
        if ok[k]:
            bs.set_truth_from_dataset(k, ds)
        else:
            bs.set_truth(k, np.zeros((0, 10)))
    st = bs.error_stats()
    t3 = time.perf_counter()
    print(f"{len(runs)} runs ({a.shape}, {dss[0].n_times} state times each): simulate {t1 - t0:.1f} s, build_and_solve {t2 - t1:.2f} s, "
          f"error report {t3 - t2:.2f} s; {int(ok.sum())} solved")
    print("mean ATE per cell, orientation deg / position m (rows: orientation noise rad, columns: position noise m)")
    print("ori\\pos   " + "".join(f"{n:>17.3f}" for n in NOISES))
    for h in range(5):
        row = []
        for i in range(5):
            ks = [k for k, (hh, ii, _) in enumerate(runs) if hh == h and ii == i and ok[k]]
            row.append((np.mean(st[ks, 0, 1]), np.mean(st[ks, 1, 1])) if ks else (np.nan, np.nan))
        print(f"{NOISES[h]:<10.3f}" + "".join(f"   {o:6.3f} / {q:5.3f}" for o, q in row))
    toff_err = [abs(bs.calibration(k)[2] - dss[k].truth["viconimu_dt"]) for k in range(len(runs)) if ok[k]]
    print(f"time offset error: median {1e3 * np.median(toff_err):.3f} ms, max {1e3 * np.max(toff_err):.3f} ms")
    if a.out_dir:
        n = bs.write_all(a.out_dir, with_txt=True)
        print(f"wrote {n} runs under {a.out_dir}")
    bs.close()


def main_device_sim(a):
    cells = [(h, i) for h in range(5) for i in range(5)]
    runs = [(h, i, j) for (h, i) in cells for j in range(a.seeds)]
    shape = dict(sim.TRAJ_SHAPES["tum_corridor1"])
    if a.duration is not None:
        shape["n_rows"] = max(8, min(shape["n_rows"], int(round(a.duration / shape["dt"])) + 1))
    rates = dict(freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, state_freq=100) if a.shape == "C1" else \
        dict(freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0)
    t0 = time.perf_counter()
    sig = np.array([(NOISES[h],) * 3 + (NOISES[i],) * 3 for (h, i, _) in runs])
    sb = sim.SimBatch(sim.make_trajectory(**shape), len(runs), device=a.device, seeds=[j for (_, _, j) in runs], vicon_sigmas=sig, **rates)
    sb.run()
    t1 = time.perf_counter()
    p = api.default_params()
    bs = api.BatchSolver(len(runs), device=a.device)
    for k in range(len(runs)):
        bs.set_trial_sim(k, p, sb)
    bs.build_and_solve()
    t2 = time.perf_counter()
    ok = np.array([bs.info(k).status == 0 for k in range(len(runs))])
    for k in range(len(runs)):
        if ok[k]:
            bs.set_truth_from_sim(k, sb)
        else:
            bs.set_truth(k, np.zeros((0, 10)))
    st = bs.error_stats()
    t3 = time.perf_counter()
    print(f"{len(runs)} runs ({a.shape}, {sb.sizes(0)[2]} state times each, device simulator): simulate {t1 - t0:.2f} s "
          f"({sb.generate_ms():.1f} ms on the device), build_and_solve {t2 - t1:.2f} s, error report {t3 - t2:.2f} s; {int(ok.sum())} solved")
    print("mean ATE per cell, orientation deg / position m (rows: orientation noise rad, columns: position noise m)")
    print("ori\\pos   " + "".join(f"{n:>17.3f}" for n in NOISES))
    for h in range(5):
        row = []
        for i in range(5):
            ks = [k for k, (hh, ii, _) in enumerate(runs) if hh == h and ii == i and ok[k]]
            row.append((np.mean(st[ks, 0, 1]), np.mean(st[ks, 1, 1])) if ks else (np.nan, np.nan))
        print(f"{NOISES[h]:<10.3f}" + "".join(f"   {o:6.3f} / {q:5.3f}" for o, q in row))
    toff_err = [abs(bs.calibration(k)[2] - sb.truth(k)["viconimu_dt"]) for k in range(len(runs)) if ok[k]]
    print(f"time offset error: median {1e3 * np.median(toff_err):.3f} ms, max {1e3 * np.max(toff_err):.3f} ms")
    if a.out_dir:
        n = bs.write_all(a.out_dir, with_txt=True)
        print(f"wrote {n} runs under {a.out_dir}")
    bs.close()
    sb.close()


if __name__ == "__main__":
    main()
