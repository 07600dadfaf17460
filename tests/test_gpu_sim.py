"""GPU tests (-m gpu) of the batched simulator (csrc/k_sim.cu) against the host simulator, which is
pinned on the reference's Simulator / BsplineSE3 (tests/test_ref_pin_cpu.py), and of the device-to-device hand-over to
the batch solver.  The stamps, the per-seed truth and the accept / reject pattern of the noise streams are exact; the
measurements differ from the host's by the last bits of log() in the polar method and by FMA contraction in the spline
evaluation -- the tolerances below are those round-off levels, written next to each assert."""
import numpy as np
import pytest

from conftest import default_params

pytestmark = pytest.mark.gpu


def _rows(n_rows):
    from vicon2gt_b200 import sim

    return sim.make_trajectory(**{**sim.TRAJ_SHAPES["tum_corridor1"], "n_rows": n_rows})


def _quat_angle(a, b):
    return 2 * np.arccos(np.clip(np.abs(np.sum(a * b, axis=1)), 0, 1))


def test_device_simulator_matches_host_simulator():
    from vicon2gt_b200 import sim

    rows = _rows(500)  # 25 s: ~9.9 k IMU samples = ~76 k accepted pairs = ~620 regenerated MT19937 blocks per trial
    seeds = [0, 1, 7, 1023, 123456]
    sig = np.array([[0.0174] * 3 + [0.05] * 3, [1e-3] * 3 + [1e-2] * 3, [0.005] * 3 + [0.01] * 3, [0.05] * 3 + [0.10] * 3,
                    [0.0174] * 3 + [0.05] * 3])
    with sim.SimBatch(rows, len(seeds), seeds=seeds, vicon_sigmas=sig) as sb:
        sb.run()
        assert sb.generate_ms() > 0
        for k, seed in enumerate(seeds):
            h = sim.simulate(traj_rows=rows, seed=seed, vicon_sigmas=sig[k])
            d = sb.dataset(k)
            assert np.array_equal(d.imu_t, h.imu_t) and np.array_equal(d.vic_t, h.vic_t) and np.array_equal(d.state_t, h.state_t)
            assert h.n_imu > 9000 and h.n_vicon > 2000
            # IMU: |w| ~ 1 rad/s, |a| ~ 10 m/s^2; a wrong stream position would show up at the noise level (1e-3 .. 1e-1)
            assert np.abs(d.imu_w - h.imu_w).max() < 1e-11
            assert np.abs(d.imu_a - h.imu_a).max() < 1e-10
            assert _quat_angle(d.vic_q, h.vic_q).max() < 1e-7  # arccos near 1: 1e-7 rad is round-off of the dot product
            assert np.abs(d.vic_q - h.vic_q).max() < 1e-12
            assert np.abs(d.vic_p - h.vic_p).max() < 1e-11
            # Simulator::get_state_in_vicon at the state times: pose, velocity, interpolated true biases
            assert np.array_equal(d.truth["states_ok"], h.truth["states_ok"]) and h.truth["states_ok"].sum() > 100
            assert np.abs(d.truth["states"] - h.truth["states"]).max() < 1e-10
            assert np.abs(d.truth["states"][:, 11:]).max() > 0  # the bias random walk is there
    # a second batch of the same seeds reproduces the first bit for bit (fixed stream order, no atomics)
    with sim.SimBatch(rows, 2, seeds=[7, 1023], vicon_sigmas=sig[2:4]) as sb2:
        sb2.run()
        a = sb2.dataset(1, with_truth_states=False)
    with sim.SimBatch(rows, 1, seeds=[1023], vicon_sigmas=sig[3:4]) as sb3:
        sb3.run()
        b = sb3.dataset(0, with_truth_states=False)
    for x, y in ((a.imu_w, b.imu_w), (a.imu_a, b.imu_a), (a.vic_q, b.vic_q), (a.vic_p, b.vic_p)):
        assert np.array_equal(x, y)


def test_other_rates_and_capped_stream():
    from vicon2gt_b200 import sim

    rows = _rows(300)
    kw = dict(freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0, max_imu=1111)
    with sim.SimBatch(rows, 2, seeds=[3, 4], **kw) as sb:
        sb.run()
        for k, seed in enumerate([3, 4]):
            h = sim.simulate(traj_rows=rows, seed=seed, **kw)
            d = sb.dataset(k, with_truth_states=False)
            assert h.n_imu == 1111 and np.array_equal(d.imu_t, h.imu_t) and np.array_equal(d.vic_t, h.vic_t)
            assert np.abs(d.imu_w - h.imu_w).max() < 1e-11 and np.abs(d.imu_a - h.imu_a).max() < 1e-10
            assert np.abs(d.vic_p - h.vic_p).max() < 1e-11 and np.abs(d.vic_q - h.vic_q).max() < 1e-12


def test_simulated_batch_goes_to_the_solver_without_a_host_round_trip():
    """v2gt_batch_set_trial_sim: the rows copied device -> device give the same solve, bit for bit, as the same rows
    downloaded and fed through the host path; and the solve recovers the simulated truth."""
    from vicon2gt_b200 import api, sim

    rows = _rows(400)
    seeds = [11, 12, 13]
    p = default_params()
    with sim.SimBatch(rows, 3, seeds=seeds, vicon_sigmas=(1e-3,) * 3 + (1e-2,) * 3) as sb:
        sb.run()
        with api.BatchSolver(3) as dev_fed, api.BatchSolver(3) as host_fed:
            for k in range(3):
                dev_fed.set_trial_sim(k, p, sb)
                host_fed.set_trial_dataset(k, p, sb.dataset(k, with_truth_states=False))
            dev_fed.build_and_solve()
            host_fed.build_and_solve()
            for k in range(3):
                ia, ib = dev_fed.info(k), host_fed.info(k)
                assert ia.status == 0 and ib.status == 0 and ia.n_states == ib.n_states > 1000
                assert (ia.tries, ia.iterations) == (ib.tries, ib.iterations) and ia.final_error == ib.final_error
                ta, xa = dev_fed.states(k)
                tb, xb = host_fed.states(k)
                assert np.array_equal(ta, tb) and np.array_equal(xa, xb)
                assert np.array_equal(dev_fed.globals10(k), host_fed.globals10(k))
                # the solve recovers the simulated time offset (drawn with sigma 0.08 s) to a few milliseconds on this 20 s run
                tr = sb.truth(k)
                _, _, toff, _ = dev_fed.calibration(k)
                assert abs(toff - tr["viconimu_dt"]) < 1e-2
        # mixed batch: one trial from the simulator, one from host arrays
        with api.BatchSolver(2) as mixed:
            mixed.set_trial_dataset(0, p, sb.dataset(1, with_truth_states=False))
            mixed.set_trial_sim(1, p, sb, sim_trial=2)
            mixed.build_and_solve()
            assert mixed.info(0).status == 0 and mixed.info(1).status == 0
            with api.BatchSolver(1) as one:
                one.set_trial_sim(0, p, sb, sim_trial=2)
                one.build_and_solve()
                # (a batch of another size cuts its chains differently: same solution to round-off, not bit for bit)
                assert np.abs(one.states(0)[1] - mixed.states(1)[1]).max() < 1e-7


def test_full_size_batch_against_host_simulator_and_stream_properties():
    """BASELINE configs[3] shape (sim.launch: 400 Hz IMU, 100 Hz vicon, 100 Hz states, 299 s): one seed of a batch against the
    host simulator over all 119.6 k IMU samples / 29.9 k poses (5 860 regenerated generator blocks), and size-independent
    properties of the batch: trials with the same seed but other vicon sigmas share the IMU stream bit for bit and their
    vicon noise scales exactly with sigma; different seeds are uncorrelated; the noise has the configured spectral density."""
    from vicon2gt_b200 import sim

    rows = sim.make_trajectory(**sim.TRAJ_SHAPES["tum_corridor1"])
    seeds = [17, 17, 18, 19]
    sig = np.array([[0.0174] * 3 + [0.05] * 3, [0.0348] * 3 + [0.10] * 3, [0.0174] * 3 + [0.05] * 3, [0.0174] * 3 + [0.05] * 3])
    with sim.SimBatch(rows, 4, seeds=seeds, vicon_sigmas=sig) as sb:
        sb.run()
        d = [sb.dataset(k, with_truth_states=(k == 0)) for k in range(4)]
    h = sim.simulate(traj_rows=rows, seed=17, vicon_sigmas=sig[0])
    assert h.n_imu > 119000 and h.n_vicon > 29000 and h.n_times > 29000
    assert np.array_equal(d[0].imu_t, h.imu_t) and np.array_equal(d[0].vic_t, h.vic_t) and np.array_equal(d[0].state_t, h.state_t)
    assert np.abs(d[0].imu_w - h.imu_w).max() < 1e-11 and np.abs(d[0].imu_a - h.imu_a).max() < 1e-10
    assert np.abs(d[0].vic_q - h.vic_q).max() < 1e-12 and np.abs(d[0].vic_p - h.vic_p).max() < 1e-11
    assert np.abs(d[0].truth["states"] - h.truth["states"]).max() < 1e-10
    # same seed, doubled sigma: identical IMU rows; the position noise doubles exactly (power-of-two factor)
    assert np.array_equal(d[1].imu_w, d[0].imu_w) and np.array_equal(d[1].imu_a, d[0].imu_a)
    n1, nh = d[1].vic_p - d[0].vic_p, d[0].vic_p - h.vic_p  # sigma n (device), device - host round-off
    assert 0.04 < n1.std() < 0.06 and np.abs(nh).max() < 1e-11
    # ... so removing it twice from the doubled run gives the noise-free positions: smooth at the 1e-3 level of a 1 m/s walk
    clean = d[0].vic_p - n1
    assert np.abs(np.diff(clean, n=2, axis=0)).max() < 5e-3 and np.abs(np.diff(d[0].vic_p, n=2, axis=0)).max() > 5e-2
    # different seeds: gyro noise uncorrelated, each with the configured density sigma_w / sqrt(dt)
    e23 = d[2].imu_w - d[3].imu_w  # (bias + noise) differences of two seeds: white part dominates sample to sample
    de = np.diff(e23, axis=0)  # first differences remove the slow bias walk: var = 4 sigma^2
    sigma = 1.6968e-04 / np.sqrt(1.0 / 400.0)
    assert abs(de.std() / (2 * sigma) - 1.0) < 0.02
    # and the noise of the two seeds is uncorrelated: corr(n2 - n0, n3 - n0) = 1/2 for independent streams of equal variance
    a, b = np.diff(d[2].imu_w - d[0].imu_w, axis=0)[:, 0], np.diff(d[3].imu_w - d[0].imu_w, axis=0)[:, 0]
    assert abs(np.corrcoef(a, b)[0, 1] - 0.5) < 0.02


def test_cpp_monte_carlo_driver_matches_the_python_mirror(tmp_path):
    """examples/run_monte_carlo (C++ SimBatch + BatchSolver::set_trial_sim + get_all_states + the device error report) prints
    what the Python mirror computes for the same seeds: same batch, same kernels, same numbers."""
    import re
    import subprocess

    from vicon2gt_b200 import api, build, sim

    exe = build.build_examples().replace("run_simulation", "run_monte_carlo")
    r = subprocess.run([exe, "--seeds", "3", "--duration", "10", "--ori_noise", "0.001", "--pos_noise", "0.01", "--out-dir", str(tmp_path)],
                       stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    pat = re.compile(r"run\s+(\d+): status (\d+), (\d+) states, (\d+) iterations, cost (\S+), mean ATE (\S+) deg / (\S+) m, toff error (\S+) ms")
    got = [m.groups() for m in pat.finditer(r.stdout)]
    assert len(got) == 3 and "wrote 3 runs" in r.stdout
    rows = sim.make_trajectory(t_start=1520531829.301144123, n_rows=201, dt=0.05, extent_m=20.0)
    p = api.default_params()
    with sim.SimBatch(rows, 3, vicon_sigmas=(1e-3,) * 3 + (1e-2,) * 3) as sb:
        sb.run()
        with api.BatchSolver(3) as bs:
            for k in range(3):
                bs.set_trial_sim(k, p, sb)
            bs.build_and_solve()
            for k in range(3):
                bs.set_truth_from_sim(k, sb)
            st = bs.error_stats()
            for k, g in enumerate(got):
                inf = bs.info(k)
                assert (int(g[0]), int(g[1]), int(g[2]), int(g[3])) == (k, 0, inf.n_states, inf.iterations)
                assert abs(float(g[4]) - inf.final_error) <= 1e-6 * inf.final_error  # printed with seven digits
                assert abs(float(g[5]) - st[k, 0, 1]) < 1e-5 and abs(float(g[6]) - st[k, 1, 1]) < 1e-5
                assert abs(float(g[7]) - 1e3 * abs(bs.calibration(k)[2] - sb.truth(k)["viconimu_dt"])) < 1e-3
            # the files of run 1 are the single-trial writer's, byte for byte
            bs.write_to_file(1, str(tmp_path / "py.csv"), str(tmp_path / "py.txt"))
    assert (tmp_path / "00001_states.csv").read_bytes() == (tmp_path / "py.csv").read_bytes()
    assert (tmp_path / "00001_info.txt").read_bytes() == (tmp_path / "py.txt").read_bytes()


@pytest.mark.parametrize("k,tag", [(0, "sim5"), (1, "sim1234")])
def test_device_simulator_against_the_reference_simulator(k, tag):
    """The device simulator against outputs of the reference's OWN Simulator / BsplineSE3 (compiled where they lie, frozen in
    tests/golden/ref_pin.npz: every 8th IMU sample, every 4th vicon pose, get_state_in_vicon at 98 off-grid times)."""
    from conftest import load_golden
    from vicon2gt_b200 import sim

    G = load_golden("ref_pin")
    with sim.SimBatch(G["sim_traj"], 2, seeds=[5, 1234], vicon_sigmas=(1e-3,) * 3 + (1e-2,) * 3, freq_imu=400.0, freq_cam=20.0,
                      freq_vicon=100.0, state_freq=100) as sb:
        sb.run()
        d = sb.dataset(k, with_truth_states=False)
        st, ok = sb.state_in_vicon(k, G[tag + "_state_t"])
    imu = np.column_stack([d.imu_t, d.imu_w, d.imu_a])[::8]
    vic = np.column_stack([d.vic_t, d.vic_q, d.vic_p])[::4]
    assert (d.n_imu, d.n_vicon) == (int(G[tag + "_n_imu"]), int(G[tag + "_n_vicon"]))
    assert np.array_equal(imu[:, 0], G[tag + "_imu"][:, 0]) and np.array_equal(vic[:, 0], G[tag + "_vic"][:, 0])  # stamps
    assert np.max(np.abs(imu[:, 1:4] - G[tag + "_imu"][:, 1:4])) < 1e-12  # rad/s
    assert np.max(np.abs(imu[:, 4:7] - G[tag + "_imu"][:, 4:7])) < 1e-10  # m/s^2 (second derivative of the spline)
    assert np.max(np.abs(vic[:, 1:] - G[tag + "_vic"][:, 1:])) < 1e-12
    assert np.array_equal(ok, G[tag + "_states_ok"])
    assert np.max(np.abs(st - G[tag + "_states"])) < 1e-10
