"""CPU tests of the batched simulator's host plan (v2gt_simbatch_plan: clock schedule, per-seed truth, vicon stamps, state
grids) against the host simulator -- which is itself pinned on the reference's Simulator / BsplineSE3
(tests/test_ref_pin_cpu.py) -- and of the stream layout the device noise kernel relies on (csrc/k_sim.cu: MT19937
regenerated in three dependent thirds, 4 words per polar attempt, accepted pairs in stream order), restated in numpy
and checked against what libstdc++'s std::normal_distribution gives the host simulator."""
import numpy as np
import pytest

from vicon2gt_b200 import api, sim


def _rows(n_rows=300):
    return sim.make_trajectory(**{**sim.TRAJ_SHAPES["tum_corridor1"], "n_rows": n_rows})


def test_plan_matches_the_host_simulator_bit_for_bit():
    rows = _rows()
    seeds = [1, 5, 1023]
    sig = np.array([[0.0174] * 3 + [0.05] * 3, [1e-3] * 3 + [1e-2] * 3, [0.01] * 3 + [0.10] * 3])
    with sim.SimBatch(rows, 3, seeds=seeds, vicon_sigmas=sig) as sb:
        sb.plan()
        for k, seed in enumerate(seeds):
            ds = sim.simulate(traj_rows=rows, seed=seed, vicon_sigmas=sig[k])
            imu_t, vic_t, state_t = sb.stamps(k)
            assert np.array_equal(imu_t, ds.imu_t) and np.array_equal(vic_t, ds.vic_t) and np.array_equal(state_t, ds.state_t)
            tr = sb.truth(k)
            for key in ("R_BtoI", "p_BinI", "R_GtoV"):
                assert np.array_equal(tr[key], ds.truth[key])
            assert tr["viconimu_dt"] == ds.truth["viconimu_dt"]
            assert np.array_equal(sb.vic_R_const(k), ds.vic_R_const)
            assert sb.sizes(k) == (ds.n_imu, ds.n_vicon, ds.n_times)


def test_plan_other_rates_and_state_times_at_the_trajectory_rows():
    rows = _rows(200)
    with sim.SimBatch(rows, 1, seeds=[7], freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0, max_imu=1500) as sb:
        sb.plan()
        ds = sim.simulate(traj_rows=rows, seed=7, freq_imu=200.0, freq_cam=20.0, freq_vicon=100.0, state_freq=0, max_imu=1500)
        imu_t, vic_t, state_t = sb.stamps(0)
        assert len(imu_t) == 1500
        assert np.array_equal(imu_t, ds.imu_t) and np.array_equal(vic_t, ds.vic_t) and np.array_equal(state_t, ds.state_t)


def test_mixed_rates_are_rejected_and_no_device_fails_loudly():
    rows = _rows(100)
    sb = sim.SimBatch(rows, 2)
    p = sim.default_sim_params(seed=1, sim_freq_imu=200.0)
    api._check(sb._L.v2gt_simbatch_set_trial(sb.h, 1, p), "set_trial")
    with pytest.raises(api.V2gtError) as e:
        sb.plan()
    assert e.value.code == 1
    sb.close()
    if api.device_count() == 0:
        with sim.SimBatch(rows, 1) as sb2:
            with pytest.raises(api.V2gtError) as e:
                sb2.run()
            assert e.value.code == 9  # V2GT_ERR_NO_DEVICE: nothing is generated on the CPU


# ---- numpy restatement of the device noise stream (k_sim_noise)
def _mt_seed(seed):
    mt = [seed & 0xFFFFFFFF]
    for i in range(1, 624):
        mt.append((1812433253 * (mt[-1] ^ (mt[-1] >> 30)) + i) & 0xFFFFFFFF)
    return np.array(mt, dtype=np.uint32)


def _twist(a, b, c):
    y = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7FFFFFFF))
    return c ^ (y >> np.uint32(1)) ^ np.where(y & np.uint32(1), np.uint32(0x9908B0DF), np.uint32(0))


def _regenerate(o):
    n = np.zeros(624, dtype=np.uint32)
    t = np.arange(227)
    n[t] = _twist(o[t], o[t + 1], o[t + 397])
    i = 227 + t
    n[i] = _twist(o[i], o[i + 1], n[i - 227])
    i = 454 + np.arange(170)
    n[i] = _twist(o[i], np.where(i == 623, n[0], o[np.minimum(i + 1, 623)]), n[i - 227])
    return n


def _temper(y):
    y = y ^ (y >> np.uint32(11))
    y = y ^ ((y << np.uint32(7)) & np.uint32(0x9D2C5680))
    y = y ^ ((y << np.uint32(15)) & np.uint32(0xEFC60000))
    return y ^ (y >> np.uint32(18))


def _normals(seed, n_blocks):
    mt, words = _mt_seed(seed), []
    for _ in range(n_blocks):
        mt = _regenerate(mt)
        words.append(_temper(mt))
    w = np.concatenate(words)
    f = w.astype(np.float64).reshape(-1, 4)

    def canon(w0, w1):
        c = (w0 + w1 * 4294967296.0) * 2.0 ** -64
        return np.where(c >= 1.0, np.nextafter(1.0, 0.0), c)

    x, y = 2.0 * canon(f[:, 0], f[:, 1]) - 1.0, 2.0 * canon(f[:, 2], f[:, 3]) - 1.0
    r2 = x * x + y * y
    acc = ~((r2 > 1.0) | (r2 == 0.0))
    mult = np.sqrt(-2 * np.log(r2[acc]) / r2[acc])
    return w, np.stack([y[acc] * mult, x[acc] * mult], 1).reshape(-1)


def test_noise_stream_layout_against_libstdcxx():
    seed = 5
    words, vals = _normals(seed, 12)
    # the three-thirds regeneration gives MT19937's words (numpy's legacy seeding is init_genrand, as std::mt19937(seed))
    ref = np.random.RandomState(seed).randint(0, 2 ** 32, size=len(words), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(words, ref)
    rows = _rows(100)
    a = sim.simulate(traj_rows=rows, seed=seed, vicon_sigmas=(0.0174,) * 3 + (0.05,) * 3)
    # truth draws 3..6 of the per-seed stream (Simulator.cpp:33-36): values in (returned, saved) order per accepted pair
    assert np.array_equal(a.truth["p_BinI"], 0.2 * vals[3:6]) and a.truth["viconimu_dt"] == 0.08 * vals[6]
    # pose j owns values [6j, 6j+6): isolate the position noise through two runs that differ only in sigma_p
    b = sim.simulate(traj_rows=rows, seed=seed, vicon_sigmas=(0.0174,) * 3 + (0.10,) * 3)
    n = min(a.n_vicon, len(vals) // 6)
    assert n > 100
    j = np.arange(n)
    got = (b.vic_p[:n] - a.vic_p[:n]) / 0.05
    want = np.stack([vals[6 * j + 3], vals[6 * j + 4], vals[6 * j + 5]], 1)
    assert np.abs(got - want).max() < 1e-12
    # sample i owns values [12i, 12i+12): isolate the gyro noise through two runs that differ only in sigma_w
    c = sim.simulate(traj_rows=rows, seed=seed, sigma_w=2 * 1.6968e-04)
    m = min(a.n_imu, len(vals) // 12)
    i = np.arange(m)
    got = (c.imu_w[:m] - a.imu_w[:m]) / (1.6968e-04 / np.sqrt(1.0 / 400.0))
    want = np.stack([vals[12 * i], vals[12 * i + 1], vals[12 * i + 2]], 1)
    assert m > 100 and np.abs(got - want).max() < 1e-9


def test_plan_against_the_reference_simulator():
    """The host plan against outputs of the reference's OWN Simulator / BsplineSE3 (compiled where they lie and frozen in
    tests/golden/ref_pin.npz by tools/make_ref_golden.py): same sample counts, same stamps, the same std::mt19937 draws
    for the perturbed calibration (scalars bit for bit, rotations to an ulp)."""
    from conftest import load_golden

    G = load_golden("ref_pin")
    tags = [("sim5", 5), ("sim1234", 1234)]
    with sim.SimBatch(G["sim_traj"], 2, seeds=[s for _, s in tags], vicon_sigmas=(1e-3,) * 3 + (1e-2,) * 3, freq_imu=400.0, freq_cam=20.0,
                      freq_vicon=100.0, state_freq=100) as sb:
        sb.plan()
        for k, (tag, _) in enumerate(tags):
            imu_t, vic_t, _ = sb.stamps(k)
            assert (len(imu_t), len(vic_t)) == (int(G[tag + "_n_imu"]), int(G[tag + "_n_vicon"]))
            assert np.array_equal(imu_t[::8], G[tag + "_imu"][:, 0]) and np.array_equal(vic_t[::4], G[tag + "_vic"][:, 0])
            tr = sb.truth(k)
            mine = np.concatenate([tr["R_BtoI"].reshape(-1), tr["p_BinI"], [tr["viconimu_dt"]], tr["R_GtoV"].reshape(-1)])
            assert np.max(np.abs(mine - G[tag + "_truth"])) < 1e-15
            assert mine[12] == G[tag + "_truth"][12] and np.array_equal(mine[9:12], G[tag + "_truth"][9:12])
