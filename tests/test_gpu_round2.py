"""GPU tests (-m gpu) of the second round's host-side changes: per-trial relinearisation budgets inside one batch, an
IMU store fed out of time order, CUDA-graph replay of a damped solve, two devices / two host threads, and the
in-library driver of the chain-sharded solve."""
import threading

import numpy as np
import pytest

from conftest import default_params

pytestmark = pytest.mark.gpu


def _solve_single(ds, p, **opts):
    from vicon2gt_b200 import api

    bs = api.BatchSolver(1, **opts)
    bs.set_trial_dataset(0, p, ds)
    bs.build_and_solve()
    out = (bs.info(0), bs.states(0), bs.globals10(0), bs.lm_trace(0))
    bs.close()
    return out


def test_mixed_relinearisation_budgets_in_one_batch(ds_c2_small):
    """ViconGraphSolver.cpp:163 loops `num_loop_relin + 1` times PER SOLVER: a trial with budget 0 inside a batch whose
    other trials relinearise must keep the result (and the counters) of its only pass."""
    from vicon2gt_b200 import api

    ds = ds_c2_small
    p0 = default_params(num_loop_relin=0, lm_max_iterations=6)
    p1 = default_params(num_loop_relin=1, lm_max_iterations=6)
    # fixed segment count: the batch and the single-trial solves then run the same arithmetic
    i0, (t0, x0), g0, _ = _solve_single(ds, p0, segments=8)
    i1, (t1, x1), g1, _ = _solve_single(ds, p1, segments=8)
    assert i0.status == 0 and i1.status == 0
    assert np.max(np.abs(x0 - x1)) > 0  # the second pass does move the estimate
    bs = api.BatchSolver(3, segments=8)
    for k, p in enumerate((p0, p1, p0)):
        bs.set_trial_dataset(k, p, ds)
    bs.build_and_solve()
    for k, (ir, xr, gr) in enumerate(((i0, x0, g0), (i1, x1, g1), (i0, x0, g0))):
        inf = bs.info(k)
        _, x = bs.states(k)
        assert inf.status == 0
        assert (inf.iterations, inf.tries, inf.linearizations) == (ir.iterations, ir.tries, ir.linearizations)
        assert inf.final_error == ir.final_error
        assert np.array_equal(x, xr) and np.array_equal(bs.globals10(k), gr)
    bs.close()


def test_imu_store_fed_out_of_time_order(ds_c2_small):
    """Propagator::feed_imu never sorts and propagate() scans from sample 0 (Propagator.cpp:21-31, :46).  A stale sample
    in the middle of the store changes what the reference selects for the interval around it; the CUDA path must select
    the same samples (scan from 0 for such a trial) -- compared with the oracle in its reference-faithful scan mode."""
    from oracle import binding as ob
    from vicon2gt_b200 import api

    ds = ds_c2_small
    p = default_params()
    # a stale copy of sample 1000 re-delivered 1500 samples later, sitting between two samples that bracket a state time
    k = None
    for i in range(2500, 2600):
        inside = (ds.state_t > ds.imu_t[i]) & (ds.state_t < ds.imu_t[i + 1])
        if inside.any():
            k = i
            break
    assert k is not None
    ins = lambda a: np.concatenate([a[: k + 1], a[1000:1001], a[k + 1:]])  # noqa: E731
    imu_t, imu_w, imu_a = ins(ds.imu_t), ins(ds.imu_w), ins(ds.imu_a)
    assert not np.all(np.diff(imu_t) >= 0)
    orc = ob.OracleSolver(p, imu_t, imu_w, imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t, vic_R_const=ds.vic_R_const)
    orc.set_option("faithful_scan", 1)
    assert orc.clean() == 0
    rc = orc.build(True)
    bs = api.BatchSolver(1, keep_cov=1)
    bs.set_trial(0, p, imu_t, imu_w, imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t, vic_R_const=ds.vic_R_const)
    bs.upload()
    bs.build(True)
    inf = bs.info(0)
    assert (inf.status == 0) == (rc == 0), (inf.status, rc)
    if rc == 0:
        po, _ = orc.preint()
        pg = bs.preint(0)
        assert np.array_equal(po[:, 0], pg[:, 0])
        assert np.max(np.abs(po[:, 1:11] - pg[:, 1:11])) < 1e-12
        # and it is NOT what the time-ordered store gives: the interval around the stale sample differs
        ref = api.BatchSolver(1)
        ref.set_trial_dataset(0, p, ds)
        ref.upload()
        ref.build(True)
        assert np.max(np.abs(ref.preint(0)[:, 1:11] - pg[:, 1:11])) > 1e-6
        ref.close()
    bs.close()
    orc.close()


@pytest.mark.parametrize("dsname", ["ds_c1_small", "ds_c3_small"])
def test_graph_replay_is_bitwise_the_plain_launch_sequence(dsname, request):
    """One damped solve captured as a CUDA graph and replayed = the same kernels with the same arguments."""
    ds = request.getfixturevalue(dsname)
    p = default_params()
    ia, (ta, xa), ga, tra = _solve_single(ds, p, graph=0)
    ib, (tb, xb), gb, trb = _solve_single(ds, p, graph=1)
    assert ia.status == 0 and ib.status == 0
    assert (ia.iterations, ia.tries) == (ib.iterations, ib.tries) and ia.final_error == ib.final_error
    assert np.array_equal(xa, xb) and np.array_equal(ga, gb) and np.array_equal(tra, trb)


def test_two_handles_on_two_host_threads(ds_c2_small, ds_c3_small):
    """Per-device function attributes and handle-private streams: two batches driven concurrently from two host threads
    (on two devices when the box has them) give the single-threaded answers."""
    from vicon2gt_b200 import api

    p = default_params(lm_max_iterations=5)
    dss = [ds_c2_small, ds_c3_small]
    want = [_solve_single(ds, p, segments=6) for ds in dss]
    ndev = api.device_count()
    got, errs = [None, None], []

    def work(k):
        try:
            dev = k % ndev
            bs = api.BatchSolver(1, device=dev, segments=6)
            bs.set_trial_dataset(0, p, dss[k])
            bs.build_and_solve()
            got[k] = (bs.info(0), bs.states(0)[1])
            bs.close()
        except Exception as e:  # noqa: BLE001
            errs.append(e)

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errs, errs
    for k in range(2):
        assert got[k][0].status == 0 and got[k][0].final_error == want[k][0].final_error
        assert np.array_equal(got[k][1], want[k][1][1])


def test_empty_batch_error_stats_rows_are_defined(ds_c2_small):
    """v2gt_batch_error_stats with no solved state must hand back defined rows (zeros, count 0), not device garbage."""
    from vicon2gt_b200 import api

    ds = ds_c2_small
    p = default_params(lm_max_iterations=2)
    bs = api.BatchSolver(1)
    bs.set_trial(0, p, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t + 1e4, vic_R_const=ds.vic_R_const)
    bs.build_and_solve()
    assert bs.info(0).status == 5
    bs.set_truth(0, np.zeros((0, 10)))
    st = bs.error_stats()
    assert np.all(np.asarray(st) == 0.0)
    bs.close()


@pytest.mark.parametrize("graph", [False, True])
def test_chain_in_library_driver_world_of_one(ds_c3_small, graph):
    """v2gt_chain_optimize on a world of one rank (no NCCL): the chain-mode kernels, both exchanges (as device copies) and
    the graph replay against the ordinary single-GPU solve."""
    from vicon2gt_b200 import api
    from vicon2gt_b200.chain import ChainSolver, NativeComm

    from conftest import assert_same_solution_or_toff_split

    ds = ds_c3_small
    p = default_params()
    single = api.BatchSolver(1)
    single.set_trial_dataset(0, p, ds)
    single.build_and_solve()
    i1, (t1, x1), g1 = single.info(0), single.states(0), single.globals10(0)
    cs = ChainSolver(p, ds, NativeComm(None), segs_per_rank=6, use_graph=graph)
    ic = cs.optimize()
    tc, xc = cs.states()
    qc, pc, toffc, thc = cs.calibration()
    assert ic.status == 0 and np.array_equal(tc, t1)
    rel, share = assert_same_solution_or_toff_split(single, (i1, x1, g1), (ic, xc, np.concatenate([qc, pc, [toffc], thc])))
    print("chain (in-library, world 1) vs single: rel chi2", rel, "share of a gap carried by t_off", share, "tries", ic.tries, i1.tries)
    single.close()
    assert cs.check_brackets()
    # a second solve from a fresh build replays the cached graph
    cs.rebuild()
    ic2 = cs.optimize()
    assert ic2.final_error == ic.final_error and ic2.tries == ic.tries
    cs.close()


def test_sixteen_full_size_monte_carlo_trials_in_one_batch_against_the_oracle():
    """The first 16 trials of the benchmark's Monte-Carlo batch (seeds 0..15, 29.8 k states each, the reference's TUM
    trajectory file) solved as ONE batch: every trial against the CPU oracle's solution of the same trial
    (tests/golden/mc16_oracle.npz, tools/make_mc_golden.py): states / calibration 1e-6, chi2 1e-8 relative."""
    import os

    from conftest import ROOT
    from vicon2gt_b200 import api, sim

    path = os.path.join(ROOT, "tests", "golden", "mc16_oracle.npz")
    if not os.path.exists(path):
        pytest.skip("tests/golden/mc16_oracle.npz missing (tools/make_mc_golden.py)")
    import importlib.util

    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(ROOT, "tools", "make_ref_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    G = np.load(path)
    n, stride = int(G["n_trials"]), int(G["stride"])
    p = default_params()
    dss = [sim.make_config("C1", seed=s, trajectory="file") for s in range(n)]
    bs = api.BatchSolver(n)
    for s, ds in enumerate(dss):
        assert np.array_equal(mod.dataset_digest(ds), G[f"s{s}_digest"]), "simulator output changed: rerun tools/make_mc_golden.py"
        bs.set_trial_dataset(s, p, ds)
    bs.build_and_solve()
    worst = [0.0, 0.0, 0.0]
    for s in range(n):
        inf = bs.info(s)
        _, x = bs.states(s)
        assert inf.status == 0 and inf.n_states == int(G[f"s{s}_n"])
        dx = float(np.max(np.abs(x[::stride] - G[f"s{s}_x"])))
        dg = float(np.max(np.abs(bs.globals10(s) - G[f"s{s}_g"])))
        rel = abs(inf.final_error - float(G[f"s{s}_final_error"])) / float(G[f"s{s}_final_error"])
        worst = [max(worst[0], dx), max(worst[1], dg), max(worst[2], rel)]
        assert dx < 1e-6 and dg < 1e-6 and rel < 1e-8, (s, dx, dg, rel, inf.iterations, int(G[f"s{s}_iterations"]))
    print("16 full-size trials vs oracle: max |dx| %.2e, max |d calibration| %.2e, max rel chi2 %.2e" % tuple(worst))
    bs.close()


def test_wave_runner_matches_one_batch(ds_c1_small, ds_c2_small, ds_c3_small):
    """api.WaveRunner (two resident handles, staging / read-back threads overlapped with the solve): five trials in waves of
    two give, trial by trial, what one five-trial batch gives (fixed segment count: the same arithmetic, bit for bit)."""
    from vicon2gt_b200 import api

    p = default_params(lm_max_iterations=5)
    dss = [ds_c2_small, ds_c1_small, ds_c3_small, ds_c2_small, ds_c3_small]
    with api.BatchSolver(len(dss), segments=6) as bs:
        for k, ds in enumerate(dss):
            bs.set_trial_dataset(k, p, ds)
        bs.build_and_solve()
        want = [(bs.info(k).final_error, bs.states(k)[1], bs.globals10(k)) for k in range(len(dss))]
    seen = 0
    with api.WaveRunner(wave=2, segments=6) as runner:
        for first, off, t, x, g, infos in runner.run(p, dss):
            for i, inf in enumerate(infos):
                k = first + i
                assert inf.status == 0 and inf.final_error == want[k][0]
                assert np.array_equal(x[off[i]:off[i + 1]], want[k][1]) and np.array_equal(g[i], want[k][2])
                seen += 1
        assert runner.device_ms > 0 and runner.wall_s > 0
    assert seen == len(dss)


def test_cpp_wave_runner_example():
    """examples/run_monte_carlo --wave: the C++ WaveRunner of include/vicon2gt_b200.hpp against the one-batch run of the
    same example (same seeds: the printed costs agree)."""
    import os
    import re
    import subprocess

    from conftest import ROOT
    from vicon2gt_b200 import build

    build.build_examples()
    exe = os.path.join(ROOT, "examples", "run_monte_carlo")
    args = [exe, "--seeds", "5", "--duration", "10", "--ori_noise", "0.001", "--pos_noise", "0.01"]
    a = subprocess.run(args, capture_output=True, text=True, timeout=300)
    b = subprocess.run(args + ["--wave", "2"], capture_output=True, text=True, timeout=300)
    assert a.returncode == 0 and b.returncode == 0, a.stderr + b.stderr
    pat = re.compile(r"run\s+(\d+): status (\d+), (\d+) states, (\d+) iterations, cost (\S+?)[,\n]")
    ga, gb = pat.findall(a.stdout), pat.findall(b.stdout + "\n")
    assert len(ga) == 5 and len(gb) == 5
    for x, y in zip(ga, gb):
        assert x[:3] == y[:3] and x[1] == "0"
        assert abs(float(x[4]) - float(y[4])) <= 2e-6 * float(x[4])
