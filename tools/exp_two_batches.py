"""Experiment: one 128-trial batch vs two 64-trial batches driven concurrently from two host threads (each batch has its own
non-blocking stream): do the HBM-bound phases of one half overlap the FP64-bound elimination of the other?"""
import os, sys, threading, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api
import bench

T = int(sys.argv[1]) if len(sys.argv) > 1 else 128
parts = int(sys.argv[2]) if len(sys.argv) > 2 else 2
trials = bench.make_trials(T, 100, 0)
p = api.default_params()

def run(groups, reps=2):
    solvers = []
    for g in groups:
        bs = api.BatchSolver(len(g))
        for i, ds in enumerate(g):
            bs.set_trial_dataset(i, p, ds)
        bs.upload()
        solvers.append(bs)
    def work(bs):
        bs.build(True); bs.optimize(); bs.sync()
    best = 1e9
    for r in range(reps + 1):
        th = [threading.Thread(target=work, args=(bs,)) for bs in solvers]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        dt = time.perf_counter() - t0
        if r > 0: best = min(best, dt)
    ok = sum(1 for bs in solvers for i in range(bs.n_trials) if bs.info(i).status == 0)
    for bs in solvers: bs.close()
    return best, ok

dt1, ok1 = run([trials])
print(f"1 x {T}: {dt1*1e3:.0f} ms  {T/dt1:.1f} traj/s ok={ok1}", flush=True)
k = T // parts
dt2, ok2 = run([trials[i*k:(i+1)*k] for i in range(parts)])
print(f"{parts} x {k}: {dt2*1e3:.0f} ms  {T/dt2:.1f} traj/s ok={ok2}", flush=True)
