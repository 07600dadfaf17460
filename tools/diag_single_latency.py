"""Latency of one damped solve on a single trajectory (second half of the metric): plain launches (per-phase CUDA events)
against the captured CUDA graph, for a few segment counts."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api, sim
p = api.default_params()
for name, kw in (("C1", {}), ("C5", {})):
    ds = sim.make_config(name, seed=5, **kw)
    for opts in ({"graph": 0}, {"graph": 1}, {"graph": 1, "tries_per_sync": 16}, {"graph": 1, "segments": 512}, {"graph": 1, "segments": 1024},
                 {"graph": 1, "segments": 2048}):
        bs = api.BatchSolver(1, **opts)
        bs.set_trial_dataset(0, p, ds)
        bs.upload(); bs.build(True); bs.optimize(); bs.build(True); bs.optimize()
        ms, l = bs.timing()
        inf = bs.info(0)
        names = ["build", "linearize", "eliminate", "reduced", "backsub", "cost", "lm", "optimize_total"]
        print(name, opts, "n", inf.n_states, "tries", inf.tries, "ms/try %.3f" % (ms[7] / inf.tries), {k: round(float(v) / inf.tries * 1e3) for k, v in zip(names[1:7], ms[1:7])}, "us/try", flush=True)
        bs.close()
