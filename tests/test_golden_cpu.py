"""CPU tests (no GPU): the oracle reproduces the committed golden vectors (tests/golden/*.npz, made by
tools/make_golden.py from the oracle itself -- a regression pin; the reference has no vectors and cannot run here)."""
import os

import numpy as np
import pytest

from conftest import default_params, load_golden


@pytest.mark.parametrize("name", ["c2_8s", "c1_12s"])
def test_oracle_reproduces_golden(name):
    from oracle import binding as ob
    from vicon2gt_b200._cdefs import TrialInfo

    g = load_golden(name)
    p = default_params(lm_max_iterations=int(g["lm_max_iterations"]))
    mk = lambda: ob.OracleSolver(p, g["imu_t"], g["imu_w"], g["imu_a"], g["vic_t"], g["vic_q"], g["vic_p"], g["state_t"],
                                 vic_R_const=g["vic_R_const"])
    o = mk()
    assert o.clean() == 0 and o.build(True) == 0
    t0, x0 = o.states()
    assert np.array_equal(t0, g["init_times"]) and np.allclose(x0, g["init_states"], rtol=0, atol=1e-13)
    pre, cov = o.preint()
    assert np.array_equal(pre[:, 0], g["preint"][:, 0])
    assert np.allclose(pre[:, :63], g["preint"], rtol=1e-12, atol=1e-15)
    assert np.allclose(cov, g["preint_cov"], rtol=1e-10, atol=1e-30)
    it = o.eval_interp(g["interp_t"], 0.1)
    for k in ("idx0", "idx1", "ok"):
        assert np.array_equal(it[k], g["interp_" + k])
    lin = o.linearize()
    assert abs(lin["cost"] - g["lin_cost"]) <= 1e-12 * g["lin_cost"]
    assert np.allclose(lin["g"], g["lin_g"], rtol=1e-9, atol=1e-9 * np.abs(g["lin_g"]).max())
    assert np.allclose(lin["Gamma"], g["lin_Gamma"], rtol=1e-10, atol=1e-10 * np.abs(g["lin_Gamma"]).max())
    o2 = mk()
    assert o2.build_and_solve() == 0
    inf = o2.info(TrialInfo)
    assert (inf.iterations, inf.tries) == (int(g["iterations"]), int(g["tries"]))
    assert abs(inf.final_error - g["final_error"]) <= 1e-10 * g["final_error"]
    _, x1 = o2.states()
    assert np.allclose(x1, g["final_states"], rtol=0, atol=1e-8)
    assert np.allclose(o2.globals10(), g["final_globals"], rtol=0, atol=1e-8)
