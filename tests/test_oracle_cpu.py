"""CPU tests (no GPU): the oracle -- the C++ restatement of the reference's build-and-solve path in oracle/ --
against the known-answer identities, reference quirks and finite differences listed in SURVEY.md 8(c).

The reference ships no tests and cannot be built here (Eigen, Boost, GTSAM 4.2a5, ROS1 are absent), so these
identities plus the committed golden vectors (tests/golden, made by tools/make_golden.py) are what pins the oracle.
"""
import numpy as np
import pytest

from conftest import default_params

SIG = (1.6968e-4, 1.9393e-5, 2.0e-3, 3.0e-3)


@pytest.fixture(scope="module")
def ob():
    from oracle import binding

    return binding


def _rand_quat(rng):
    q = rng.standard_normal(4)
    q /= np.linalg.norm(q)
    return q if q[3] >= 0 else -q


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


# ---------------------------------------------------------------------------------------- quat_ops.h / rpy_ops.h
def test_quaternion_helpers(ob):
    rng = np.random.default_rng(0)
    for _ in range(50):
        q = _rand_quat(rng)
        R = ob.quat_2_Rot(q)
        assert np.allclose(R @ R.T, np.eye(3), atol=1e-14) and abs(np.linalg.det(R) - 1) < 1e-13
        # quat_2_Rot (quat_ops.h:153-158): (2w^2-1) I - 2w [v]x + 2 v v^T
        v, w = q[:3], q[3]
        assert np.allclose(R, (2 * w * w - 1) * np.eye(3) - 2 * w * _skew(v) + 2 * np.outer(v, v), atol=1e-15)
        q2 = ob.rot_2_quat(R)  # largest-pivot branches, w >= 0, unit norm (quat_ops.h:85-120)
        assert q2[3] >= 0 and abs(np.linalg.norm(q2) - 1) < 1e-15
        assert np.allclose(q2, q, atol=1e-14)
        # quat_multiply(q, p) = L(q) p, w forced >= 0 and renormalised (quat_ops.h:181-196); R(q (x) p) = R(q) R(p)
        p = _rand_quat(rng)
        qp = ob.quat_multiply(q, p)
        assert qp[3] >= 0 and abs(np.linalg.norm(qp) - 1) < 1e-15
        assert np.allclose(ob.quat_2_Rot(qp), R @ ob.quat_2_Rot(p), atol=1e-14)


def test_so3_helpers(ob):
    rng = np.random.default_rng(1)
    assert np.array_equal(ob.exp_so3(np.zeros(3)), np.eye(3))  # theta == 0 branch (quat_ops.h:232-253)
    assert np.array_equal(ob.log_so3(np.eye(3)), np.zeros(3))  # exact-identity test (quat_ops.h:268-289)
    for _ in range(50):
        w = rng.standard_normal(3)
        w *= rng.uniform(1e-3, 3.0) / np.linalg.norm(w)  # |w| < pi
        R = ob.exp_so3(w)
        th = np.linalg.norm(w)
        K = _skew(w / th)
        assert np.allclose(R, np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K, atol=1e-14)
        assert np.allclose(ob.log_so3(R), w, atol=1e-9)  # acos form: accuracy degrades, reproduced not improved
        # Jl_so3 (quat_ops.h:492-502) against a finite difference of exp: exp(w + d) ~ exp(Jl d) exp(w)
        J = ob.Jl_so3(w)
        d = 1e-6 * rng.standard_normal(3)
        lhs = ob.exp_so3(w + d) @ R.T
        assert np.allclose(ob.log_so3(lhs), J @ d, atol=5e-11)
    # small-angle limit
    assert np.allclose(ob.Jl_so3(np.array([1e-9, 0, 0])), np.eye(3), atol=1e-8)


def test_rpy(ob):
    # rot2rpy(rot_y(ty) rot_x(tx)) recovers roll, pitch (rpy_ops.h:68-74), used for G(0) (ViconGraphSolver.cpp:403-404)
    tx, ty = -0.2556, 0.0063
    cx, sx, cy, sy = np.cos(tx), np.sin(tx), np.cos(ty), np.sin(ty)
    R = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]]) @ np.array([[1, 0, 0], [0, cx, -sx], [0, sx, cx]])
    rpy = ob.rot2rpy(R)
    assert np.allclose(rpy[:2], [tx, ty], atol=1e-14)


# ---------------------------------------------------------------------------------------- retract (a17, a18)
def test_retract(ob):
    rng = np.random.default_rng(2)
    x = np.concatenate([_rand_quat(rng), rng.standard_normal(12)])
    # delta = 0 hits the NaN branch of JPLNavState::retract (sin(0)/0) -> identity quaternion
    assert np.allclose(ob.retract_state(x, np.zeros(15)), x, atol=0)
    d = 1e-2 * rng.standard_normal(15)
    y = ob.retract_state(x, d)
    assert np.allclose(y[4:], x[4:] + d[3:], atol=0)
    th = np.linalg.norm(d[:3])
    dq = np.concatenate([np.sin(th / 2) * d[:3] / th, [np.cos(th / 2)]])
    assert np.allclose(y[:4], ob.quat_multiply(dq, x[:4]), atol=1e-15)  # left-multiplicative JPL
    # RotationXY: theta <- wrap2pi(theta + wrap2pi(delta))
    assert np.allclose(ob.retract_rotxy([3.0, -3.0], [0.5, -0.5]), [3.5 - 2 * np.pi, -3.5 + 2 * np.pi], atol=1e-15)


# ---------------------------------------------------------------------------------------- CpiV1 (a3-a8)
def _imu_store(n=81, dt=0.0025, t0=1.5e9):
    return t0 + np.arange(n) * dt


def test_preintegration_zero_rate_constant_acceleration(ob):
    t = _imu_store()
    a = np.array([0.3, -0.2, 9.81])
    out, P, ns = ob.propagate(SIG, t, np.zeros((len(t), 3)), np.tile(a, (len(t), 1)), t[3], t[23])
    T = out[0]
    assert T == t[23] - t[3] and ns == 20  # sequentially accumulated DT equals t1 - t0 exactly (epoch-scale stamps)
    assert np.allclose(out[1:4], 0.5 * a * T * T, rtol=1e-12)  # alpha
    assert np.allclose(out[4:7], a * T, rtol=1e-12)  # beta
    assert np.array_equal(out[7:11], [0, 0, 0, 1])  # q_k2tau
    I = np.eye(3).ravel()
    assert np.allclose(out[17:26], T * I, rtol=1e-12)  # J_q
    assert np.allclose(out[44:53], -T * I, rtol=1e-12)  # H_beta
    assert np.allclose(out[53:62], -0.5 * T * T * I, rtol=1e-10)  # H_alpha
    assert np.allclose(P, P.T, atol=0) and np.all(np.linalg.eigvalsh(P) > 0)
    # white-noise part of the covariance: P_theta = sigma_w^2 T I, P_bg = sigma_wb^2 T I (F has no feedback into them at w = 0)
    assert np.allclose(np.diag(P)[0:3], SIG[0] ** 2 * T, rtol=1e-6)
    assert np.allclose(np.diag(P)[3:6], SIG[1] ** 2 * T, rtol=1e-9)
    assert np.allclose(np.diag(P)[9:12], SIG[3] ** 2 * T, rtol=1e-9)


def test_preintegration_constant_rate_rotation(ob):
    t = _imu_store()
    w = np.array([0.4, -0.3, 0.5])
    out, _, _ = ob.propagate(SIG, t, np.tile(w, (len(t), 1)), np.zeros((len(t), 3)), t[0], t[40])
    T = out[0]
    R = ob.quat_2_Rot(out[7:11])
    assert np.allclose(R, ob.exp_so3(-w * T), atol=1e-13)  # R_k2tau = exp(-w T) in the JPL convention (CpiV1.h:113-118)
    assert np.allclose(out[1:7], 0, atol=1e-15)


def test_preintegration_endpoints_interpolated_and_selection_modes_agree(ob):
    rng = np.random.default_rng(3)
    t = _imu_store(n=200)
    w = 0.3 * rng.standard_normal((200, 3))
    a = np.array([0, 0, 9.81]) + rng.standard_normal((200, 3))
    t0, t1 = t[10] + 0.0011, t[35] + 0.0007  # both strictly inside a sample interval: CASE 1 and CASE 3 lerps
    fast = ob.propagate(SIG, t, w, a, t0, t1, bg_lin=(0.01, -0.02, 0.005), ba_lin=(0.1, 0.0, -0.1))
    slow = ob.propagate(SIG, t, w, a, t0, t1, bg_lin=(0.01, -0.02, 0.005), ba_lin=(0.1, 0.0, -0.1), faithful=True)
    assert fast is not None and slow is not None
    assert fast[2] == slow[2] == 26  # 2 end-point pieces + 24 whole samples
    assert fast[0][0] == t1 - t0
    # binary-search selection + block-sparse RK4 == the reference's linear scan + dense 15x15 RK4
    assert np.allclose(fast[0][:62], slow[0][:62], rtol=0, atol=1e-15)
    assert np.allclose(fast[1], slow[1], rtol=1e-11, atol=1e-30)
    # splitting at an existing sample: the two halves compose to the whole (Appendix A associativity)
    tm = t[20]
    A = ob.propagate(SIG, t, w, a, t0, tm)[0]
    B = ob.propagate(SIG, t, w, a, tm, t1)[0]
    W = ob.propagate(SIG, t, w, a, t0, t1)[0]
    RA, RB = ob.quat_2_Rot(A[7:11]), ob.quat_2_Rot(B[7:11])
    assert np.allclose(ob.quat_2_Rot(W[7:11]), RB @ RA, atol=1e-14)
    assert np.allclose(W[4:7], A[4:7] + RA.T @ B[4:7], atol=1e-15)  # beta
    assert np.allclose(W[1:4], A[1:4] + A[4:7] * B[0] + RA.T @ B[1:4], atol=1e-15)  # alpha
    assert np.allclose(W[17:26].reshape(3, 3), RB @ A[17:26].reshape(3, 3) + B[17:26].reshape(3, 3), atol=1e-15)  # J_q


def test_preintegration_fails_without_bounding_samples(ob):
    t = _imu_store(n=10)
    z = np.zeros((10, 3))
    assert ob.propagate(SIG, t, z, z, t[0] - 1.0, t[0] - 0.5) is None


# ---------------------------------------------------------------------------------------- Interpolator (a13, a14)
def _solver(ob, ds, **kw):
    return ob.OracleSolver.from_dataset(default_params(**kw), ds)


def test_interpolator_bracket_quirks(ob, ds_c2_small):
    ds = ds_c2_small
    o = _solver(ob, ds)
    vt = ds.vic_t
    k = 100
    q = np.array([vt[k], 0.5 * (vt[k] + vt[k + 1]), vt[0], vt[1], vt[-1], vt[0] - 1.0, vt[-1] + 1.0, np.nextafter(vt[k], np.inf)])
    r = o.eval_interp(q, 0.1)
    # exact hit of sample k: bounds are k-1 and k+1 (equal_range on a set with unique keys, Interpolator.cpp:175-237)
    assert (r["idx0"][0], r["idx1"][0], r["ok"][0]) == (k - 1, k + 1, 1)
    assert (r["idx0"][1], r["idx1"][1], r["ok"][1]) == (k, k + 1, 1)
    assert r["ok"][2] == 0  # lo == begin, even on an exact hit of the first sample
    assert (r["idx0"][3], r["idx1"][3], r["ok"][3]) == (0, 2, 1)
    assert r["ok"][4] == 0 and r["ok"][5] == 0 and r["ok"][6] == 0  # hi == end / outside
    assert (r["idx0"][7], r["idx1"][7], r["ok"][7]) == (k, k + 1, 1)
    # gap threshold: 0.1 s in get_pose_with_jacobian, 5 s in get_pose; a 7 ms threshold rejects a mid-interval query
    assert o.eval_interp(q[1:2], 0.004)["ok"][0] == 0
    # lambda -> 0 limit returns pose k (p linear, q = q_k)
    assert np.allclose(r["p"][7], ds.vic_p[k], atol=1e-9)
    assert abs(abs(r["q"][7] @ ds.vic_q[k]) - 1) < 1e-12
    # midpoint: p is the mean, and the reference's covariance bug: position block = (1-l)^2 R_q1 + l^2 R_p1
    assert np.allclose(r["p"][1], 0.5 * (ds.vic_p[k] + ds.vic_p[k + 1]), atol=1e-12)
    Rq, Rp = ds.vic_R_const[:9].reshape(3, 3), ds.vic_R_const[9:].reshape(3, 3)
    lam = (q[1] - vt[k]) / (vt[k + 1] - vt[k])
    assert np.allclose(r["R"][1].reshape(6, 6)[3:, 3:], (1 - lam) ** 2 * Rq + lam ** 2 * Rp, rtol=1e-12)
    # H_toff position rows = -(p1 - p0) / (t1 - t0)
    assert np.allclose(r["H_toff"][1][3:], -(ds.vic_p[k + 1] - ds.vic_p[k]) / (vt[k + 1] - vt[k]), rtol=1e-12)


def test_interpolator_duplicates_and_order(ob, ds_c2_small):
    """std::set semantics: poses fed out of order are sorted, duplicate timestamps keep the first one."""
    ds = ds_c2_small
    n = len(ds.vic_t)
    perm = np.random.default_rng(4).permutation(n)
    vt = np.concatenate([ds.vic_t[perm], ds.vic_t[:5]])
    vq = np.concatenate([ds.vic_q[perm], ds.vic_q[5:10]])  # the duplicates carry different poses: must be ignored
    vp = np.concatenate([ds.vic_p[perm], ds.vic_p[5:10] + 1.0])
    a = _solver(ob, ds)
    b = ob.OracleSolver(default_params(), ds.imu_t, ds.imu_w, ds.imu_a, vt, vq, vp, ds.state_t, vic_R_const=ds.vic_R_const)
    assert b.n_vicon == n
    q = np.linspace(ds.vic_t[2], ds.vic_t[40], 57)
    ra, rb = a.eval_interp(q, 0.1), b.eval_interp(q, 0.1)
    for key in ("idx0", "idx1", "ok", "q", "p"):
        assert np.array_equal(ra[key], rb[key])


# ---------------------------------------------------------------------------------------- factors (a10, a11, a15)
def _fd_state(f, x, ob, eps=1e-6):
    cols = []
    for k in range(15):
        d = np.zeros(15)
        d[k] = eps
        cols.append((f(ob.retract_state(x, d)) - f(ob.retract_state(x, -d))) / (2 * eps))
    return np.array(cols).T


def test_imu_factor_jacobians_match_finite_differences(ob, ds_c3_small):
    o = _solver(ob, ds_c3_small)
    assert o.clean() == 0 and o.build(True) == 0
    pre, _ = o.preint()
    _, X = o.states()
    rng = np.random.default_rng(5)
    g = 9.81
    for k in (3, 57, 120):
        xi, xj = X[k].copy(), X[k + 1].copy()
        xi[4:7] += 1e-3 * rng.standard_normal(3)  # non-zero bias offsets from the linearisation point
        xi[10:13] += 1e-2 * rng.standard_normal(3)
        th = np.array([0.1, -0.05])
        r = ob.eval_imu_factor(pre[k], g, xi, xj, th)
        H1 = _fd_state(lambda x: ob.eval_imu_factor(pre[k], g, x, xj, th)["e"], xi, ob)
        H2 = _fd_state(lambda x: ob.eval_imu_factor(pre[k], g, xi, x, th)["e"], xj, ob)
        H3 = np.array([(ob.eval_imu_factor(pre[k], g, xi, xj, ob.retract_rotxy(th, d))["e"] -
                        ob.eval_imu_factor(pre[k], g, xi, xj, ob.retract_rotxy(th, -d))["e"]) / 2e-6
                       for d in (np.array([1e-6, 0]), np.array([0, 1e-6]))]).T
        for A, F in ((r["H1"], H1), (r["H2"], H2), (r["H3"], H3)):
            assert np.max(np.abs(A - F)) < 2e-6 * max(1.0, np.max(np.abs(A)))


def test_vicon_factor_whitening_and_exact_jacobian_blocks(ob, ds_c3_small):
    ds = ds_c3_small
    o = _solver(ob, ds)
    assert o.clean() == 0 and o.build(True) == 0
    t, X = o.states()
    g10 = o.globals10()
    g10[4:7] = [0.05, -0.1, 0.02]
    k = 40
    r = o.eval_vicon_factor(t[k], X[k], g10)
    assert r["has_vicon"]
    # position columns of H1 and the p_BinI Jacobian are exact (W does not depend on p): finite differences agree
    def e_of_p(dp):
        x = X[k].copy()
        x[13:16] += dp
        return o.eval_vicon_factor(t[k], x, g10)["e"]
    for c in range(3):
        d = np.zeros(3)
        d[c] = 1e-6
        assert np.allclose((e_of_p(d) - e_of_p(-d)) / 2e-6, r["H1"][:, 12 + c], atol=1e-5 * np.max(np.abs(r["H1"])))
    def e_of_c1(dp):
        gg = g10.copy()
        gg[4:7] += dp
        return o.eval_vicon_factor(t[k], X[k], gg)["e"]
    for c in range(3):
        d = np.zeros(3)
        d[c] = 1e-6
        assert np.allclose((e_of_c1(d) - e_of_c1(-d)) / 2e-6, r["H3"][:, c], atol=1e-5 * np.max(np.abs(r["H3"])))
    # bg, v, ba columns are structurally zero
    assert np.array_equal(r["H1"][:, 3:12], np.zeros((6, 9)))
    # a query without bounding vicon poses contributes nothing
    far = o.eval_vicon_factor(ds.vic_t[-1] + 10.0, X[k], g10)
    assert not far["has_vicon"] and np.array_equal(far["e"], np.zeros(6)) and np.array_equal(far["H1"], np.zeros((6, 15)))


def test_estimate_flags_zero_the_calibration_columns(ob, ds_c3_small):
    o = _solver(ob, ds_c3_small, estimate_toff_vicon_to_imu=0, estimate_ori_vicon_to_imu=0, estimate_pos_vicon_to_imu=0)
    assert o.clean() == 0 and o.build(True) == 0
    lin = o.linearize()
    G = lin["Gamma"].reshape(9, 9)
    assert np.array_equal(G[:7, :7], np.zeros((7, 7)))  # fixed calibration = zero Jacobian columns (GtsamConfig.h)
    assert np.all(np.diag(G)[7:] > 0)  # gravity angles are always estimated
    assert np.array_equal(lin["B"].reshape(-1, 15, 9)[:, :, :7], np.zeros((len(lin["B"]), 15, 7)))


# ---------------------------------------------------------------------------------------- normal equations + LM (a21)
def test_linear_solve_satisfies_the_damped_normal_equations(ob, ds_c2_small):
    import scipy.sparse as sp
    import scipy.sparse.linalg as spl

    o = _solver(ob, ds_c2_small)
    assert o.clean() == 0 and o.build(True) == 0
    lin = o.linearize()
    n = o.n_states
    N = 15 * n + 9
    H = sp.lil_matrix((N, N))
    for i in range(n):
        s = slice(15 * i, 15 * i + 15)
        H[s, s] = lin["D"][i].reshape(15, 15)
        if i + 1 < n:
            H[s, 15 * i + 15:15 * i + 30] = lin["O"][i].reshape(15, 15)
            H[15 * i + 15:15 * i + 30, s] = lin["O"][i].reshape(15, 15).T
        H[s, 15 * n:] = lin["B"][i].reshape(15, 9)
        H[15 * n:, s] = lin["B"][i].reshape(15, 9).T
    H[15 * n:, 15 * n:] = lin["Gamma"].reshape(9, 9)
    H = H.tocsc()
    assert abs(H - H.T).max() < 1e-6 * abs(H).max()
    g = np.concatenate([lin["g"].ravel(), lin["g_gamma"]])
    for lam in (1e-5, 1.0, 1e6):
        delta, ok = o.solve(lam)
        assert ok
        A = H + lam * sp.identity(N, format="csc")
        ref = spl.spsolve(A, g)
        sc = np.abs(ref).reshape(-1)[: 15 * n].reshape(n, 15).max(axis=0)
        assert np.max(np.abs(delta[: 15 * n] - ref[: 15 * n]).reshape(n, 15) / sc) < 1e-6
        assert np.allclose(delta[15 * n:], ref[15 * n:], rtol=1e-5, atol=1e-12)
    # the gradient of the nonlinear cost is -g (b = -e_w): directional derivative check along a random direction
    rng = np.random.default_rng(6)
    d = rng.standard_normal(N) * 1e-7
    c0 = o.cost()
    assert abs((o.cost_at(d) - o.cost_at(-d)) / 2 + g @ d) < 1e-3 * abs(g @ d) + 1e-9 * c0


def test_lm_policy_and_truth_recovery(ob, ds_c2_small):
    from vicon2gt_b200._cdefs import TrialInfo

    ds = ds_c2_small
    o = _solver(ob, ds)
    assert o.build_and_solve() == 0
    inf = o.info(TrialInfo)
    tr = o.lm_trace()
    assert inf.status == 0 and 1 <= inf.iterations <= 30 and inf.tries == len(tr) and inf.final_error < inf.initial_error
    assert tr[0, 0] == 1e-5  # lambdaInitial
    for a, b in zip(tr[:-1], tr[1:]):
        # success: lambda / 10 ; failure: lambda * 10 (lambdaFactor = 10, useFixedLambdaFactor)
        assert np.isclose(b[0], a[0] / 10 if a[4] == 1 else a[0] * 10, rtol=1e-12)
        if a[4] == 1:
            assert b[1] == a[2]  # the accepted cost becomes the current one
    assert int(tr[:, 4].sum()) == inf.iterations
    # truth recovery at the tech report's level for this noise (1 mrad / 1 cm vicon): mm / ms / mrad
    q, p, toff, th = o.calibration()
    assert abs(toff - ds.truth["viconimu_dt"]) < 5e-3
    assert np.max(np.abs(p - ds.truth["p_BinI"])) < 1e-2
    assert 2 * np.arccos(min(1.0, abs(q @ ob.rot_2_quat(ds.truth["R_BtoI"])))) < 5e-3
    assert np.max(np.abs(th - ob.rot2rpy(ds.truth["R_GtoV"])[:2])) < 5e-3
    t, X = o.states()
    ts = ds.truth["states"]
    idx = np.searchsorted(ds.state_t, t)
    assert np.array_equal(ds.state_t[idx], t)
    truth = ts[idx]
    pos_err = np.linalg.norm(X[:, 13:16] - truth[:, 5:8], axis=1)
    assert np.sqrt(np.mean(pos_err ** 2)) < 1e-2


def test_failure_codes(ob, ds_c2_small):
    from vicon2gt_b200._cdefs import TrialInfo

    ds = ds_c2_small
    p = default_params()
    none = ob.OracleSolver(p, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, np.zeros(0), vic_R_const=ds.vic_R_const)
    assert none.build_and_solve() == 3  # empty timestamps (ViconGraphSolver.cpp:99-104)
    one = ob.OracleSolver(p, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t[:1], ds.vic_q[:1], ds.vic_p[:1], ds.state_t,
                          vic_R_const=ds.vic_R_const)
    assert one.build_and_solve() == 4  # < 2 vicon poses (:107-112)
    far = ob.OracleSolver(p, ds.imu_t, ds.imu_w, ds.imu_a, ds.vic_t, ds.vic_q, ds.vic_p, ds.state_t + 1e4, vic_R_const=ds.vic_R_const)
    assert far.build_and_solve() == 5  # every timestamp cleaned away (:145-150)
    ok = _solver(ob, ds)
    assert ok.clean() == 0 and ok.build(True) == 0
    inf = ok.info(TrialInfo)
    # 0.5 s trimmed at each end of the vicon window (2 * TIME_OFFSET) at 20 Hz
    assert inf.removed_before + inf.removed_after + inf.removed_no_imu + inf.removed_no_vicon + inf.n_states == len(ds.state_t)
    assert inf.removed_before >= 9 and inf.removed_after >= 9


# ------------------------------------------------------------------------------------------ error report
def test_stats_ref_known_answers():
    """Stats::calculate restatement (oracle/stats_ref.py) on hand-computable inputs."""
    from oracle import stats_ref as sr

    s = sr.stats_calculate([4.0, 1.0, 3.0, 2.0])
    assert np.allclose(s, [np.sqrt(7.5), 2.5, 2.5, 1.0, 4.0, np.sqrt(5.0 / 3.0), 2.5 + 2.326 * np.sqrt(5.0 / 3.0), 4.0], rtol=1e-15)
    s = sr.stats_calculate([2.0, 7.0, 3.0])
    assert s[2] == 3.0 and s[3] == 2.0 and s[4] == 7.0 and s[7] == 3.0
    s1 = sr.stats_calculate([5.0])
    assert s1[0] == 5.0 and s1[1] == 5.0 and s1[2] == 5.0 and np.isnan(s1[5])  # std of one value: 0 / 0 as in the reference
    assert sr.stats_calculate([])[7] == 0.0


def test_trajectory_errors_ref_known_answers():
    """A 0.2 rad rotation about z and (3, 4, 0) / (0, 0, 2) offsets: ori = 0.2 rad in degrees to first order of the
    2|vec(q)| form (2 sin(0.1)), pos = 5, vel = 2; the unique-quaternion flip makes the sign of q irrelevant."""
    from oracle import stats_ref as sr

    q_gt = np.array([0.0, 0.0, np.sin(0.1), np.cos(0.1)])
    x = np.zeros((2, 16))
    x[:, 3] = 1.0
    x[1, :4] = [0.0, 0.0, 0.0, -1.0]
    x[:, 13:16] = [3.0, 4.0, 0.0]
    x[:, 7:10] = [0.0, 0.0, 2.0]
    truth = np.zeros((2, 10))
    truth[:, 0:4] = q_gt
    e = sr.trajectory_errors(x, truth)
    assert np.allclose(e[:, 0], 180.0 / np.pi * 2.0 * np.sin(0.1), rtol=1e-14)
    assert np.allclose(e[:, 1], 5.0) and np.allclose(e[:, 2], 2.0)
    assert np.allclose(sr.quat_multiply(q_gt, sr.quat_inv(q_gt)), [0, 0, 0, 1])
