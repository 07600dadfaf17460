import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api, sim
ds = sim.make_config("C3", seed=3, trajectory="synth", duration_s=60.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2))
p = api.default_params()
for seg in (0, 20, 36, 30, 48):
    bs = api.BatchSolver(1, **({"segments": seg} if seg else {}))
    bs.set_trial_dataset(0, p, ds)
    bs.build_and_solve()
    inf = bs.info(0)
    tr = bs.lm_trace(0)
    print("segments", seg, "n", inf.n_states, "tries", inf.tries, "iters", inf.iterations, "state", inf.lm_state, "err %.12e" % inf.final_error)
    if seg in (0, 36):
        print(tr[:12])
    bs.close()
