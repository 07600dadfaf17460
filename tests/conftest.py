"""pytest configuration: the `gpu` marker, native builds, and shared synthetic datasets."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on a B200)")


def _has_gpu():
    try:
        from vicon2gt_b200 import api

        return api.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _native_libs():
    """Build the host simulator + oracle here; the CUDA library must already be in-tree (build() makes it)."""
    from vicon2gt_b200 import build

    build.build_sim()
    build.build_oracle()
    if not os.path.exists(build.lib_path("cuda")):
        build.build_cuda()
    yield


def default_params(**kw):
    """v2gt_params with the reference defaults, filled by the ORACLE's copy of the defaults so that CPU-only
    tests do not depend on the CUDA library."""
    from oracle import binding as ob

    return ob.default_params(**kw)


@pytest.fixture(scope="session")
def ds_c1_small():
    """sim.launch-shaped (400 Hz IMU, 100 Hz vicon, 100 Hz states; every state time hits a vicon stamp) 8 s."""
    from vicon2gt_b200 import sim

    return sim.make_config("C1", seed=1, trajectory="synth", duration_s=8.0)


@pytest.fixture(scope="session")
def ds_c3_small():
    """TUM-shaped (200 Hz IMU, 100 Hz vicon, 20 Hz states at the trajectory stamps) 40 s, low vicon noise."""
    from vicon2gt_b200 import sim

    return sim.make_config("C3", seed=3, trajectory="synth", duration_s=40.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2))


@pytest.fixture(scope="session")
def ds_c2_small():
    """EuRoC-shaped (200 Hz IMU, 100 Hz vicon, 20 Hz states) 25 s."""
    from vicon2gt_b200 import sim

    return sim.make_config("C2", seed=7, trajectory="synth", duration_s=25.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2))


def load_golden(name):
    """tests/golden/<name>.npz: fixed inputs + the oracle's outputs on them (tools/make_golden.py)."""
    return dict(np.load(os.path.join(ROOT, "tests", "golden", name + ".npz")))


def assert_same_solution_or_toff_split(single, ref, other, tol_chi2=1e-8):
    """Two solves of the SAME problem whose arithmetic differs at round-off (other segment geometry, sharded reductions).
    `ref` = (info, states, globals10) of `single` (an api.BatchSolver holding trial 0), `other` the same triple of the
    second solve.  Always: estimates within 1e-6, and each reported chi2 is the cost of its own estimate (re-evaluated by
    `single`'s cost function).  chi2 itself within `tol_chi2` -- unless the two estimates sit on different PLATEAUS of the
    cost: the vicon factor queries the interpolator at `t_I - t_off`, ONE fp64 subtraction at epoch scale
    (MeasBased_ViconPoseTimeoffsetFactor.cpp:52; timestamps ~1.5e9 s, ulp 2.4e-7 s), so the cost is piecewise constant in
    t_off with steps every 2.4e-7 s at which the query times of ALL states move by one ulp together; and because the
    factor differentiates as if its whitening W(t_off) were constant (:52-66 against :126-132) Levenberg-Marquardt stops
    where the cost still slopes steeply along t_off (~3e5 per second on the 40 s fixture), so neighbouring plateaus differ
    by ~3e-5 relative (measured with the oracle: +0.0658 on 1997.3 between t_off + 1e-10 and t_off + 1e-9, flat over the
    rest of +-1e-7).  Two round-off-different solves that agree to 1e-8 in t_off land on different plateaus once in a
    while.  In that case the gap must be carried by the time offset alone: moving t_off from one estimate to the other
    moves the cost by the gap.
    Returns (relative chi2 gap, share of the gap carried by t_off or nan)."""
    i1, x1, g1 = ref
    ic, xc, gc = other
    assert ic.status == 0 and i1.status == 0 and xc.shape == x1.shape
    assert np.max(np.abs(xc[:, 4:] - x1[:, 4:])) < 1e-6 and np.max(np.abs(gc[4:] - g1[4:])) < 1e-6
    dq = np.abs(np.sum(xc[:, :4] * x1[:, :4], axis=1))
    assert (2 * np.arccos(np.clip(dq, 0, 1))).max() < 1e-6
    dqg = abs(float(np.dot(gc[:4], g1[:4])))
    assert 2 * np.arccos(min(1.0, dqg)) < 1e-6
    single.set_estimate(0, xc, gc)
    assert abs(single.cost(0) - ic.final_error) / ic.final_error < 1e-10
    single.set_estimate(0, x1, g1)
    assert abs(single.cost(0) - i1.final_error) / i1.final_error < 1e-10
    gap = ic.final_error - i1.final_error
    rel = abs(gap) / i1.final_error
    share = float("nan")
    if rel >= tol_chi2:
        g_sw = g1.copy()
        g_sw[7] = gc[7]
        single.set_estimate(0, x1, g_sw)
        share = (single.cost(0) - i1.final_error) / gap
        single.set_estimate(0, x1, g1)
        assert abs(gc[7] - g1[7]) > 1e-10 and 0.05 < share < 20.0, (rel, gc[7] - g1[7], share)
    return rel, share
