import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vicon2gt_b200 import api, sim
p = api.default_params()
T = 8
bs = api.BatchSolver(T)
for i in range(T):
    bs.set_trial_dataset(i, p, sim.make_config("C1", seed=i))
bs.build_and_solve()
for i in range(T):
    inf = bs.info(i)
    tr = bs.lm_trace(i)
    failed = np.isinf(tr[:, 2])
    print(f"trial {i}: n {inf.n_states} tries {inf.tries} iters {inf.iterations} state {inf.lm_state} failed-solve-or-negative-model {failed.sum()} accepted {int(tr[:,4].sum())}")
    if i == 0:
        for r in tr:
            print("  lam %.0e err %.9e new %.9e lin %.3e acc %d" % tuple(r))
bs.close()
