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
