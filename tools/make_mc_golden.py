#!/usr/bin/env python
"""tests/golden/mc16_oracle.npz: the first 16 trials of the benchmark's Monte-Carlo batch (BASELINE.json configs[3]: sim.launch
defaults on data/tum_corridor1_512_16_okvis.txt, seeds 0..15, 29.8 k states each) solved by the CPU oracle -- converged
calibration, chi2, LM counters and every 128th state of every trial.  The GPU test solves the 16 trials as ONE batch and
compares each of them (1e-6 / 1e-8).  About 45 s of CPU per trial (threads: argv[1], default 6)."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
N_TRIALS, STRIDE = 16, 128


def main():
    from oracle import binding as ob
    from tools.make_ref_golden import dataset_digest
    from vicon2gt_b200 import sim
    from vicon2gt_b200._cdefs import TrialInfo

    nthr = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    out = {"n_trials": np.int64(N_TRIALS), "stride": np.int64(STRIDE)}
    lock = threading.Lock()

    def work(k):
        for seed in range(k, N_TRIALS, nthr):
            ds = sim.make_config("C1", seed=seed, trajectory="file")
            o = ob.OracleSolver.from_dataset(ob.default_params(), ds)
            assert o.build_and_solve() == 0
            inf = o.info(TrialInfo)
            t, x = o.states()
            with lock:
                out.update({f"s{seed}_digest": dataset_digest(ds), f"s{seed}_n": np.int64(len(t)), f"s{seed}_x": x[::STRIDE].copy(),
                            f"s{seed}_g": o.globals10(), f"s{seed}_final_error": np.float64(inf.final_error),
                            f"s{seed}_iterations": np.int64(inf.iterations), f"s{seed}_tries": np.int64(inf.tries)})
                print(seed, len(t), inf.iterations, inf.tries, inf.final_error, flush=True)
            o.close()

    ths = [threading.Thread(target=work, args=(k,)) for k in range(nthr)]
    for th in ths:
        th.start()
    for th in ths:
        th.join()
    path = os.path.join(ROOT, "tests", "golden", "mc16_oracle.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
