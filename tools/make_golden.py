#!/usr/bin/env python
"""Generate tests/golden/*.npz: small fixed inputs plus the oracle's outputs on them.

The reference's solver cannot be run in this environment (GTSAM 4.2a5, Boost, ROS1 absent) and holds no test
vectors of its own, so these files pin the ORACLE (oracle/, the CPU restatement of
ViconGraphSolver::build_and_solve) against regressions and give the GPU tests a fixture that does not depend on
the simulator or the oracle being rebuilt identically.  The reference's own factor-level code IS pinned, separately:
tools/make_ref_golden.py -> tests/golden/ref_pin.npz.

    python tools/make_golden.py        # rewrites tests/golden/c2_8s.npz and c1_6s.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def make(name, cfg, seed, duration, sigmas, lm_iters, trajectory="synth"):
    from oracle import binding as ob
    from vicon2gt_b200 import sim
    from vicon2gt_b200._cdefs import TrialInfo

    ds = sim.make_config(cfg, seed=seed, trajectory=trajectory, duration_s=duration, vicon_sigmas=sigmas)
    p = ob.default_params(lm_max_iterations=lm_iters)
    o = ob.OracleSolver.from_dataset(p, ds)
    assert o.clean() == 0 and o.build(True) == 0
    t0, x0 = o.states()
    pre, cov = o.preint()
    lin = o.linearize()
    delta, ok = o.solve(1e-5)
    assert ok
    rng = np.random.default_rng(seed)
    tq = np.concatenate([ds.vic_t[::9], rng.uniform(ds.vic_t[0] - 0.02, ds.vic_t[-1] + 0.02, 300), ds.vic_t[:2], ds.vic_t[-2:]])
    it = o.eval_interp(tq, 0.1)
    o2 = ob.OracleSolver.from_dataset(p, ds)
    assert o2.build_and_solve() == 0
    inf = o2.info(TrialInfo)
    t1, x1 = o2.states()
    out = dict(
        imu_t=ds.imu_t, imu_w=ds.imu_w, imu_a=ds.imu_a, vic_t=ds.vic_t, vic_q=ds.vic_q, vic_p=ds.vic_p,
        vic_R_const=ds.vic_R_const, state_t=ds.state_t, lm_max_iterations=np.int64(lm_iters),
        init_times=t0, init_states=x0, preint=pre[:, :63], preint_cov=cov,
        lin_cost=np.float64(lin["cost"]), lin_error0=np.float64(lin["lin_error0"]), lin_D=lin["D"][::25], lin_O=lin["O"][::25],
        lin_B=lin["B"][::25], lin_g=lin["g"], lin_Gamma=lin["Gamma"], lin_g_gamma=lin["g_gamma"], solve_delta=delta,
        interp_t=tq, interp_idx0=it["idx0"], interp_idx1=it["idx1"], interp_ok=it["ok"], interp_q=it["q"], interp_p=it["p"],
        final_times=t1, final_states=x1, final_globals=o2.globals10(), final_error=np.float64(inf.final_error),
        initial_error=np.float64(inf.initial_error), iterations=np.int64(inf.iterations), tries=np.int64(inf.tries),
        lm_trace=o2.lm_trace(),
    )
    path = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB;", inf.n_states, "states,", inf.iterations, "iterations,", inf.tries, "tries")


if __name__ == "__main__":
    from vicon2gt_b200 import build

    build.build_sim()
    build.build_oracle()
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    make("c2_8s", "C2", 11, 8.0, (1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2), 30)
    # sim.launch shape on the first 12 s of the reference's TUM trajectory file (the 6 s fixture of the first round left the
    # accelerometer bias unobservable -- cond(H) ~ 2e15 -- and could not hold the 1e-6 bar)
    make("c1_12s", "C1", 1, 12.0, (0.0174, 0.0174, 0.0174, 0.05, 0.05, 0.05), 30, trajectory="file")
