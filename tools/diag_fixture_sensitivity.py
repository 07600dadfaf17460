#!/usr/bin/env python
"""How much the converged states of a golden fixture move with the segment count of the elimination (pure round-off:
every count solves the same damped systems with a backward-stable factorisation).  Prints max |x - golden| per count."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api  # noqa: E402

for name in sys.argv[1:] or ["c1_6s", "c2_8s"]:
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", name + ".npz")))
    p = api.default_params(lm_max_iterations=int(g["lm_max_iterations"]))
    for seg in (1, 2, 4, 7, 10, 14, 22, 32, 48):
        bs = api.BatchSolver(1, segments=seg)
        bs.set_trial(0, p, g["imu_t"], g["imu_w"], g["imu_a"], g["vic_t"], g["vic_q"], g["vic_p"], g["state_t"], vic_R_const=g["vic_R_const"])
        bs.build_and_solve()
        inf = bs.info(0)
        _, x = bs.states(0)
        d = np.abs(x[:, 4:] - g["final_states"][:, 4:])
        print(f"{name} segments={seg:3d}: max|dx|={d.max():.2e} (bg {d[:, 0:3].max():.1e} v {d[:, 3:6].max():.1e} ba {d[:, 6:9].max():.1e} "
              f"p {d[:, 9:12].max():.1e})  chi2 rel {abs(inf.final_error - g['final_error']) / g['final_error']:.1e}  tries {inf.tries}")
        bs.close()
