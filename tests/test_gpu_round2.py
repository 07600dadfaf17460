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

    ds = ds_c3_small
    p = default_params()
    i1, (t1, x1), g1, _ = _solve_single(ds, p)
    cs = ChainSolver(p, ds, NativeComm(None), segs_per_rank=6, use_graph=graph)
    ic = cs.optimize()
    tc, xc = cs.states()
    assert ic.status == 0 and np.array_equal(tc, t1)
    assert abs(ic.final_error - i1.final_error) / i1.final_error < 1e-8
    assert np.max(np.abs(xc[:, 4:] - x1[:, 4:])) < 1e-6
    dq = np.abs(np.sum(xc[:, :4] * x1[:, :4], axis=1))
    assert (2 * np.arccos(np.clip(dq, 0, 1))).max() < 1e-6
    assert cs.check_brackets()
    # a second solve from a fresh build replays the cached graph
    cs.rebuild()
    ic2 = cs.optimize()
    assert ic2.final_error == ic.final_error and ic2.tries == ic.tries
    cs.close()
