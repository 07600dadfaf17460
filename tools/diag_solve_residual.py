"""Diagnostic (GPU): normal-equation residual of the GPU and oracle steps on the golden fixtures."""
import sys
import numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from conftest import default_params, load_golden
from vicon2gt_b200 import api
import scipy.sparse as sp

def H_of(lin, n):
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
    return H.tocsc()

for name in ("c2_8s", "c1_3s"):
    g = load_golden(name)
    p = default_params()
    bs = api.BatchSolver(1)
    bs.set_trial(0, p, g["imu_t"], g["imu_w"], g["imu_a"], g["vic_t"], g["vic_q"], g["vic_p"], g["state_t"], vic_R_const=g["vic_R_const"])
    bs.upload(); bs.build(True)
    lin = bs.linearize(0)
    n = bs.info(0).n_states
    H = H_of(lin, n)
    rhs = np.concatenate([lin["g"].ravel(), lin["g_gamma"]])
    for lam in (1e-5, 1.0, 1e5):
        d, ok = bs.solve(0, lam)
        A = H + lam * sp.identity(H.shape[0], format="csc")
        r = A @ d - rhs
        import scipy.sparse.linalg as spl
        ref = spl.spsolve(A, rhs)
        rr = A @ ref - rhs
        line = f"{name} lam={lam:g} ok={ok} |r_gpu|/|g|={np.linalg.norm(r)/np.linalg.norm(rhs):.2e} |r_splu|/|g|={np.linalg.norm(rr)/np.linalg.norm(rhs):.2e} max|d-ref|/max|ref|={np.max(np.abs(d-ref))/np.max(np.abs(ref)):.2e}"
        if lam == 1e-5:
            go = g["solve_delta"]
            ro = A @ go - rhs
            line += f" |r_oracle|/|g|={np.linalg.norm(ro)/np.linalg.norm(rhs):.2e} max|d-orc|/max={np.max(np.abs(d-go))/np.max(np.abs(go)):.2e} max|ref-orc|={np.max(np.abs(ref-go))/np.max(np.abs(go)):.2e}"
        print(line)
    try:
        print("cond estimate (dense):", np.linalg.cond((H + 1e-5 * sp.identity(H.shape[0])).toarray()))
    except Exception as e:
        print(e)
