#!/usr/bin/env python
"""Generate tests/golden/ref_pin.npz: outputs of the REFERENCE'S OWN SOURCES on fixed inputs.

oracle/_ref/libv2gt_ref.so is built by oracle/Makefile.ref from the files where they lie under
/root/reference/src (quat_ops.h, rpy_ops.h, CpiBase.h, CpiV1.h, Propagator.cpp, Interpolator.cpp,
ImuFactorCPIv1.cpp, MeasBased_ViconPoseTimeoffsetFactor.cpp, JPLNavState.cpp, JPLQuaternion.cpp,
RotationXY.cpp) against the interface stand-ins of oracle/shim/ (Eigen / GTSAM / Boost are not in
this image).  /root/reference does not exist on the GPU box, so the reference's outputs are frozen
here; tests/test_ref_pin_cpu.py checks the oracle against them (and, where the library can be
rebuilt, that the library still reproduces this file bit for bit), tests/test_gpu_parity.py checks
the CUDA path against them.

Also frozen: end-to-end runs of the reference's ViconGraphSolver class (src/solver/ViconGraphSolver.cpp,
compiled where it lies: timestamp cleaning, state erasure, initial guess, graph construction,
relinearisation loop, CSV writer are the reference's) -- with optimizer.optimize() served by the
generic skyline-Cholesky restatement of GTSAM's LM in oracle/shim/gtsam/mini_gtsam_lm.h.

Not covered (GTSAM's own code, absent from /root/reference): Gaussian::Covariance whitening, the
Huber weights, LevenbergMarquardtOptimizer.  Those are restated from upstream twice, independently
(oracle/solver.hpp block-tridiagonal; mini_gtsam_lm.h generic), and the two are compared end to end.

    python tools/make_ref_golden.py       # needs /root/reference
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

SIGMAS = np.array([1.6968e-04, 1.9393e-05, 2.0e-3, 3.0e-3])  # run_simulation.cpp:79-82


def rand_quat(rng):
    q = rng.normal(size=4)
    return q / np.linalg.norm(q)


def rand_state(rng):
    return np.concatenate([rand_quat(rng), rng.normal(size=3) * 0.01, rng.normal(size=3), rng.normal(size=3) * 0.1,
                           rng.normal(size=3) * 3])


def helpers(rb, rng, out):
    n = 24
    q = np.array([rand_quat(rng) for _ in range(n)])
    p = np.array([rand_quat(rng) for _ in range(n)])
    w = rng.normal(size=(n, 3))
    w[0] = 0.0                                  # exp_so3 / Jl_so3 small-angle branches
    w[1] *= 1e-9
    w[2] *= 1e-5
    w[3] = w[3] / np.linalg.norm(w[3]) * (np.pi - 1e-7)   # log_so3 near pi
    w[4] = w[4] / np.linalg.norm(w[4]) * 3.0
    q[0] = [0, 0, 0, 1]
    q[1] = [1, 0, 0, 0]                         # rot_2_quat branches: trace < 0 with each diagonal dominant
    q[2] = [0, 1, 0, 0]
    q[3] = [0, 0, 1, 0]
    q[4] = np.array([0.5, 0.5, 0.5, -0.5])
    R = np.array([rb.quat_2_Rot(x) for x in q])
    Rw = np.array([rb.exp_so3(x) for x in w])
    out.update(h_q=q, h_p=p, h_w=w, h_q2R=R, h_R2q=np.array([rb.rot_2_quat(x) for x in R]),
               h_qmul=np.array([rb.quat_multiply(a, b) for a, b in zip(q, p)]), h_exp=Rw,
               h_log=np.array([rb.log_so3(x) for x in Rw]), h_log_q=np.array([rb.log_so3(x) for x in R]),
               h_Jl=np.array([rb.Jl_so3(x) for x in w]), h_rpy=np.array([rb.rot2rpy(x) for x in R]))
    th = rng.uniform(-4, 4, size=(n, 2))
    d = rng.normal(size=(n, 2)) * 10.0 ** rng.integers(-8, 2, size=(n, 1))
    out.update(rx_th=th, rx_d=d, rx_out=np.array([rb.retract_rotxy(a, b) for a, b in zip(th, d)]),
               rx_rot=np.array([rb.rotxy_rot(a) for a in th]))
    x = np.array([rand_state(rng) for _ in range(n)])
    y = np.array([rand_state(rng) for _ in range(n)])
    xi = rng.normal(size=(n, 15)) * 10.0 ** rng.integers(-12, 1, size=(n, 1))
    xi[0] = 0.0                                 # retract(0): 0/0 -> NaN branch -> identity quaternion
    xi[1, :3] = 0.0
    xi[2, :3] *= 7.0                            # large rotation (dq4 may turn negative: sign flip branch)
    out.update(rs_x=x, rs_y=y, rs_xi=xi, rs_out=np.array([rb.retract_state(a, b) for a, b in zip(x, xi)]),
               rs_local=np.array([rb.local_state(a, b) for a, b in zip(x, y)]),
               rq_out=np.array([rb.retract_quat(a[:4], b[:3]) for a, b in zip(x, xi)]))


def propagation(rb, rng, out):
    """Propagator::propagate + CpiV1::feed_IMU over a store with irregular stamps and one repeated stamp."""
    n = 360
    dt = np.full(n, 0.0025)
    dt[50:60] = 0.004
    dt[100] = 0.0                                # repeated stamp -> "Zero DT ... removing it"
    dt[200:210] = 0.0011
    t = 5.0 + np.cumsum(dt)
    w = rng.normal(size=(n, 3)) * 0.6
    w[300:] *= 1e-3                              # |w_hat| < 0.008726646: the small-angle branch of feed_IMU
    a = rng.normal(size=(n, 3)) * 2 + np.array([0.3, -0.2, 9.81])
    cases = []
    for (t0, t1) in [(t[10] + 0.0003, t[22] - 0.0007), (t[10], t[20]), (t[10], t[20] + 1e-4), (t[10] - 1e-4, t[20]),
                     (t[48], t[63]), (t[95], t[108]), (t[99], t[102]), (t[198], t[215] + 2e-4), (t[305] + 1e-4, t[340]),
                     (t[0], t[5]), (t[0] - 0.01, t[5]), (t[-6], t[-1]), (t[-6], t[-1] + 0.01), (t[-1] + 0.1, t[-1] + 0.2),
                     (t[30], t[30] + 1e-13), (t[30] + 1e-4, t[30] + 3e-4), (t[150], t[190])]:
        bg = rng.normal(size=3) * 0.01 if len(cases) % 2 else np.zeros(3)
        ba = rng.normal(size=3) * 0.1 if len(cases) % 2 else np.zeros(3)
        cases.append(np.concatenate([[t0, t1], bg, ba]))
    cases = np.array(cases)
    ok = np.zeros(len(cases), dtype=np.int32)
    rows = np.full((len(cases), 62), np.nan)
    covs = np.full((len(cases), 225), np.nan)
    for k, c in enumerate(cases):
        r = rb.propagate(SIGMAS, t, w, a, c[0], c[1], c[2:5], c[5:8])
        if r is not None:
            ok[k] = 1
            rows[k] = r[0][:62]
            covs[k] = r[1].reshape(-1)
    tq = np.concatenate([t[:2], t[-2:], [t[0] - 1e-9, t[-1] + 1e-9, t[100], 0.5 * (t[3] + t[4])]])
    out.update(pr_imu_t=t, pr_imu_w=w, pr_imu_a=a, pr_cases=cases, pr_ok=ok, pr_row=rows, pr_cov=covs,
               hb_q=tq, hb_out=rb.has_bounding_imu(t, tq))


def interpolation(rb, rng, out):
    """Interpolator: store fed out of order with repeated stamps, per-pose covariances, a 0.3 s and a 6 s dropout."""
    n = 420
    t = 20.0 + 0.01 * np.arange(n)
    t[200:] += 0.3                               # > 0.1 s: get_pose_with_jacobian refuses, get_pose interpolates
    t[320:] += 6.0                               # > 5 s: both refuse
    ang = 0.4 * np.sin(0.7 * (t - 20.0))[:, None] * np.array([0.3, -0.5, 0.8]) + rng.normal(size=(n, 3)) * 1e-3
    q = np.array([rb.rot_2_quat(rb.exp_so3(a)) for a in ang])
    p = np.stack([np.sin(t), np.cos(0.5 * t), 0.1 * t], axis=1) + rng.normal(size=(n, 3)) * 1e-3
    A = rng.normal(size=(n, 3, 3)) * 1e-2
    Rq = np.einsum("nij,nkj->nik", A, A) + 1e-6 * np.eye(3)
    A = rng.normal(size=(n, 3, 3)) * 3e-2
    Rp = np.einsum("nij,nkj->nik", A, A) + 1e-5 * np.eye(3)
    order = rng.permutation(n)
    dup = rng.choice(n, 12, replace=False)       # the same stamps fed again with other values: std::set keeps the first
    ft = np.concatenate([t[order], t[dup]])
    fq = np.concatenate([q[order], np.array([rand_quat(rng) for _ in dup])])
    fp = np.concatenate([p[order], rng.normal(size=(len(dup), 3))])
    fRq = np.concatenate([Rq[order], Rq[dup] * 2]).reshape(-1, 9)
    fRp = np.concatenate([Rp[order], Rp[dup] * 2]).reshape(-1, 9)
    ri = rb.RefInterpolator(ft, fq, fp, fRq, fRp)
    assert ri.size == n
    tq = np.concatenate([t[::5], t[:3], t[-3:], [t[0] - 1.0, t[-1] + 1.0], rng.uniform(t[0] - 0.05, t[-1] + 0.05, 900),
                         np.nextafter(t[100:106], np.inf), np.nextafter(t[100:106], -np.inf),
                         t[199] + np.array([0.05, 0.15, 0.25]), t[319] + np.array([0.5, 3.0, 5.5])])
    a = ri.eval(tq, True)
    b = ri.eval(tq, False)
    out.update(in_t=ft, in_q=fq, in_p=fp, in_Rq=fRq, in_Rp=fRp, in_query=tq)
    for k, v in a.items():
        out["inj_" + k] = v
    for k in ("ok", "q", "p", "R"):
        out["inp_" + k] = b[k]
    # vicon factor on this store (per-pose covariance, dropouts, out-of-range states)
    m = 40
    xs = np.array([rand_state(rng) for _ in range(m)])
    gs = np.array([np.concatenate([rand_quat(rng), rng.normal(size=3) * 0.2, [rng.normal() * 0.02], rng.normal(size=2) * 0.1])
                   for _ in range(m)])
    ts = rng.uniform(t[0] - 0.02, t[-1] + 0.02, m)
    ts[:6] = t[50:56] + gs[:6, 7]                # t - t_off lands on a stamp to round-off
    fl = np.ones((m, 3), dtype=np.int32)
    fl[10:14] = [[0, 1, 1], [1, 0, 1], [1, 1, 0], [0, 0, 0]]
    res = [ri.eval_vicon_factor(ts[i], xs[i], gs[i], fl[i]) for i in range(m)]
    out.update(fv_t=ts, fv_x=xs, fv_g=gs, fv_flags=fl, fv_e=np.array([r["e"] for r in res]), fv_H1=np.array([r["H1"] for r in res]),
               fv_H2=np.array([r["H2"] for r in res]), fv_H3=np.array([r["H3"] for r in res]), fv_H4=np.array([r["H4"] for r in res]))
    ri.close()


def dataset(rb, rng, out):
    """A 3 s EuRoC-shaped trial: the reference's preintegration of every interval and both factors of every state at a
    perturbed estimate (what ViconGraphSolver::build_problem + one GTSAM linearize() evaluate)."""
    from oracle import binding as ob
    from vicon2gt_b200 import sim

    ds = sim.make_config("C2", seed=11, trajectory="synth", duration_s=3.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2))
    p = ob.default_params()
    o = ob.OracleSolver.from_dataset(p, ds)     # used only for the cleaned state times and the initial guess
    assert o.clean() == 0 and o.build(True) == 0
    t, x0 = o.states()
    g0 = o.globals10()
    n = len(t)
    x = x0.copy()
    x[:, 4:7] += 1e-3 * rng.standard_normal((n, 3))
    x[:, 10:13] += 1e-2 * rng.standard_normal((n, 3))
    x[:, 13:16] += 0.05 * rng.standard_normal((n, 3))
    x[::7, 13:16] += 0.5                          # vicon outliers: |e| well beyond the Huber threshold
    g = g0.copy()
    g[4:7] += 0.05
    g[7] += 0.0123
    g[8:10] += 0.02
    rows = np.zeros((n - 1, 62))
    covs = np.zeros((n - 1, 225))
    full = np.zeros((n - 1, 288))
    for k in range(n - 1):
        r = rb.propagate(SIGMAS, ds.imu_t, ds.imu_w, ds.imu_a, t[k], t[k + 1], x0[k, 4:7], x0[k, 10:13])
        assert r is not None and r[0][0] == t[k + 1] - t[k]
        full[k], rows[k], covs[k] = r[0], r[0][:62], r[1].reshape(-1)
    Rc = np.asarray(ds.vic_R_const)
    ri = rb.RefInterpolator(ds.vic_t, ds.vic_q, ds.vic_p, Rc[:9], Rc[9:])
    fi = [rb.eval_imu_factor(full[k], p.gravity_magnitude, x[k], x[k + 1], g[8:10]) for k in range(n - 1)]
    fv = [ri.eval_vicon_factor(t[k], x[k], g) for k in range(n)]
    ri.close()
    out.update(ds_imu_t=ds.imu_t, ds_imu_w=ds.imu_w, ds_imu_a=ds.imu_a, ds_vic_t=ds.vic_t, ds_vic_q=ds.vic_q, ds_vic_p=ds.vic_p,
               ds_vic_R_const=Rc, ds_state_t=ds.state_t, ds_t=t, ds_x0=x0, ds_g0=g0, ds_x=x, ds_g=g, ds_pre_row=rows, ds_pre_cov=covs,
               ds_imu_e=np.array([r["e"] for r in fi]), ds_imu_H1=np.array([r["H1"] for r in fi]),
               ds_imu_H2=np.array([r["H2"] for r in fi]), ds_imu_H3=np.array([r["H3"] for r in fi]),
               ds_vic_e=np.array([r["e"] for r in fv]), ds_vic_H1=np.array([r["H1"] for r in fv]),
               ds_vic_H2=np.array([r["H2"] for r in fv]), ds_vic_H3=np.array([r["H3"] for r in fv]),
               ds_vic_H4=np.array([r["H4"] for r in fv]))
    # the IMU factor away from the data: random states, large bias corrections
    m = 16
    xi = np.array([rand_state(rng) for _ in range(m)])
    xj = np.array([rand_state(rng) for _ in range(m)])
    th = rng.normal(size=(m, 2)) * 0.3
    res = [rb.eval_imu_factor(full[k], 9.81, xi[k], xj[k], th[k]) for k in range(m)]
    out.update(fi_row=full[:m, :62], fi_xi=xi, fi_xj=xj, fi_th=th, fi_e=np.array([r["e"] for r in res]),
               fi_H1=np.array([r["H1"] for r in res]), fi_H2=np.array([r["H2"] for r in res]), fi_H3=np.array([r["H3"] for r in res]))


E2E_CASES = {
    # name: (config, make_config kwargs, v2gt_params overrides)
    "e2e_c2": ("C2", dict(seed=7, trajectory="synth", duration_s=25.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2)), {}),
    "e2e_c2_relin": ("C2", dict(seed=7, trajectory="synth", duration_s=25.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2)), dict(num_loop_relin=1)),
    "e2e_c3_fixed": ("C3", dict(seed=3, trajectory="synth", duration_s=40.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2)),
                     dict(estimate_toff_vicon_to_imu=0, estimate_pos_vicon_to_imu=0)),
    "e2e_c1": ("C1", dict(seed=1, trajectory="synth", duration_s=8.0), {}),
    # 11 s without vicon in the middle: build_problem erases the states inside the hole (get_pose refuses a query more
    # than 5 s from either bracket end, Interpolator.cpp:119-123).  A shorter hole would keep states whose vicon factor
    # then reads uninitialised memory in the reference (see tests/test_ref_pin_cpu.py::test_vicon_factor).
    "e2e_c2_dropout": ("C2", dict(seed=5, trajectory="synth", duration_s=34.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2)), {}),
    # the first 40 s of the reference's own data/tum_corridor1_512_16_okvis.txt: jittered state times (0.043 ... 0.055 s apart)
    "e2e_c3_file": ("C3", dict(seed=4, trajectory="file", duration_s=40.0, vicon_sigmas=(1e-3, 1e-3, 1e-3, 1e-2, 1e-2, 1e-2)), {}),
}


# BASELINE.json configs[1] and configs[2] at full size ON THE REFERENCE'S OWN TRAJECTORY FILES (tests/golden/traj_*.txt =
# data/euroc_V1_01_easy.txt, data/tum_corridor1_512_16_okvis.txt; about 30 s and 70 s of CPU each to regenerate; every 4th
# state kept) and trial 0 of the benchmark's Monte-Carlo batch (configs[0] / configs[3]: sim.launch defaults on the TUM
# file, 29.8 k states, about 6 min of CPU; every 16th state kept)
E2E_FULL = {
    "e2e_c2_full": ("C2", dict(seed=21, trajectory="file"), {}),
    "e2e_c3_full": ("C3", dict(seed=21, trajectory="file"), {}),
    "e2e_c1_full": ("C1", dict(seed=0, trajectory="file"), {}),
    # BASELINE.json configs[4]: the synthesised one-hour sequence (72 k states; no reference data file is that long), the
    # input of the chain-sharded solve.  About 15 min of CPU: generated on its own (--only e2e_c5_full) and merged.
    "e2e_c5_full": ("C5", dict(seed=5), {}),
}
FULL_STRIDE = {"e2e_c2_full": 4, "e2e_c3_full": 4, "e2e_c1_full": 16, "e2e_c5_full": 32}


def dataset_digest(ds):
    """sha256 over a trial's input arrays: the end-to-end cases store their inputs as (simulator arguments, digest)."""
    import hashlib

    h = hashlib.sha256()
    for a in (ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.vic_R_const, ds.state_t):
        h.update(np.ascontiguousarray(a, dtype=np.float64).tobytes())
    return np.frombuffer(h.digest(), dtype=np.uint8).copy()


def e2e_dataset(name):
    from oracle import binding as ob
    from vicon2gt_b200 import sim

    cfg, kw, pk = (E2E_CASES[name] if name in E2E_CASES else E2E_FULL[name])
    ds = sim.make_config(cfg, **kw)
    if name.endswith("_dropout"):
        keep = (ds.vic_t < ds.vic_t[0] + 11.5) | (ds.vic_t > ds.vic_t[0] + 22.5)
        ds.vic_t, ds.vic_q, ds.vic_p = ds.vic_t[keep].copy(), ds.vic_q[keep].copy(), ds.vic_p[keep].copy()
    return ds, ob.default_params(**pk)


def end_to_end(rb, out):
    """ViconGraphSolver::build_and_solve of the reference class: cleaned times, initial values (optimiser capped at 0
    iterations), final states / calibration / cost, and the CSV of write_to_file."""
    import tempfile

    for name in E2E_CASES:
        ds, p = e2e_dataset(name)
        r0 = rb.RefSolver(p, ds, max_iterations=0)
        r0.build_and_solve()
        a = r0.result()
        r0.close()
        rs = rb.RefSolver(p, ds)
        rs.build_and_solve()
        b = rs.result()
        out.update({name + "_digest": dataset_digest(ds), name + "_times": b["times"], name + "_x0": a["states"],
                    name + "_g0": a["globals10"], name + "_initial_error": np.float64(a["initial_error"]), name + "_x": b["states"],
                    name + "_g": b["globals10"], name + "_final_error": np.float64(b["final_error"]),
                    name + "_iterations": np.int64(b["iterations"]), name + "_tries": np.int64(b["tries"])})
        assert np.array_equal(a["times"], b["times"])
        if name == "e2e_c2":
            with tempfile.TemporaryDirectory() as d:
                rs.write_to_file(os.path.join(d, "sub", "states.csv"), os.path.join(d, "sub", "info.txt"))
                out[name + "_csv"] = np.frombuffer(open(os.path.join(d, "sub", "states.csv"), "rb").read(), dtype=np.uint8).copy()
        rs.close()


def end_to_end_full(rb, out, names=None):
    import hashlib

    for name in (names or [n for n in E2E_FULL if n != "e2e_c5_full"]):
        ds, p = e2e_dataset(name)
        rs = rb.RefSolver(p, ds)
        rs.build_and_solve()
        b = rs.result()
        rs.close()
        out.update({name + "_digest": dataset_digest(ds), name + "_n": np.int64(len(b["times"])),
                    name + "_times_digest": np.frombuffer(hashlib.sha256(b["times"].tobytes()).digest(), dtype=np.uint8).copy(),
                    name + "_stride": np.int64(FULL_STRIDE[name]), name + "_x": b["states"][::FULL_STRIDE[name]], name + "_g": b["globals10"],
                    name + "_final_error": np.float64(b["final_error"]), name + "_iterations": np.int64(b["iterations"]),
                    name + "_tries": np.int64(b["tries"])})


SIM_STRIDE_IMU, SIM_STRIDE_VICON, SIM_STRIDE_TRUTH = 8, 4, 10


def write_trajectory(rows, path):
    with open(path, "w") as f:
        for r in rows:
            f.write(" ".join(repr(float(x)) for x in r) + "\n")


def simulator(rb, out):
    """src/sim/Simulator.cpp + BsplineSE3.cpp driven as run_simulation.cpp:104-137 drives them, on a 10 s trajectory file:
    the IMU / vicon streams (every 8th / 4th sample kept), the perturbed calibration it drew, get_state_in_vicon."""
    import tempfile

    from vicon2gt_b200 import sim

    rows = sim.make_trajectory(t_start=1403715273.26214, n_rows=200, dt=0.05, extent_m=2.5)
    with tempfile.TemporaryDirectory() as d:
        path = os.path.join(d, "traj.txt")
        write_trajectory(rows, path)
        for seed, tag in ((5, "sim5"), (1234, "sim1234")):
            r = rb.simulate(path, seed=seed, freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0)
            tt = r["vic_t"][::SIM_STRIDE_TRUTH] + 0.0031
            r2 = rb.simulate(path, seed=seed, freq_imu=400.0, freq_cam=20.0, freq_vicon=100.0, truth_times=tt)
            out.update({tag + "_n_imu": np.int64(len(r["imu_t"])), tag + "_n_vicon": np.int64(len(r["vic_t"])),
                        tag + "_imu": np.column_stack([r["imu_t"], r["imu_w"], r["imu_a"]])[::SIM_STRIDE_IMU],
                        tag + "_vic": np.column_stack([r["vic_t"], r["vic_q"], r["vic_p"]])[::SIM_STRIDE_VICON],
                        tag + "_truth": np.concatenate([r["R_BtoI"].reshape(-1), r["p_BinI"], [r["viconimu_dt"]], r["R_GtoV"].reshape(-1)]),
                        tag + "_state_t": tt, tag + "_states": r2["states"], tag + "_states_ok": r2["states_ok"]})
    out["sim_traj"] = rows


def statistics(rb, rng, out):
    """Stats::calculate (src/utils/stats.h) on samples of 1, 2, odd and even length."""
    for n in (1, 2, 7, 500, 2871):
        v = np.abs(rng.normal(size=n)) * 10.0 ** rng.integers(-3, 2)
        out["st_v%d" % n] = v
        out["st_o%d" % n] = rb.stats_calculate(v)


def generate(full=True):
    from oracle import ref_binding as rb
    from vicon2gt_b200 import build

    build.build_sim()
    build.build_oracle()
    rb.build()
    out = {}
    rng = np.random.default_rng(20261020)
    helpers(rb, rng, out)
    propagation(rb, rng, out)
    interpolation(rb, rng, out)
    dataset(rb, rng, out)
    end_to_end(rb, out)
    simulator(rb, out)
    statistics(rb, rng, out)
    if full:
        end_to_end_full(rb, out)
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "ref_pin.npz")
    if len(sys.argv) > 2 and sys.argv[1] == "--only":
        # regenerate the named full-size cases only and merge them into the committed file
        from oracle import ref_binding as rb
        from vicon2gt_b200 import build

        build.build_sim()
        build.build_oracle()
        rb.build()
        out = dict(np.load(path))
        end_to_end_full(rb, out, sys.argv[2:])
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")
        sys.exit(0)
    out = generate()
    if os.path.exists(path):  # keep the separately generated one-hour case
        old = np.load(path)
        out.update({k: old[k] for k in old.files if k.startswith("e2e_c5_full_")})
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path) // 1024, "KiB,", len(out), "arrays")
