"""TEST INFRASTRUCTURE (oracle): numpy restatement of the reference's trajectory-error report.

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this; the product path
(vicon2gt_b200/csrc/k_stats.cu) never does.  stats_calculate is pinned on Stats::calculate of the reference's own
src/utils/stats.h (compiled into oracle/_ref; tests/test_ref_pin_cpu.py::test_error_statistics_against_stats_h), the
quaternion helpers on quat_ops.h the same way; trajectory_errors restates three lines of main() in run_simulation.cpp
(known-answer cases in tests/test_oracle_cpu.py).

  quat_multiply / quat_inv   src/utils/quat_ops.h:181-196, 445-450 (JPL, vector first; result forced to w >= 0 and normalised)
  trajectory_errors          src/run_simulation.cpp:216-218
  stats_calculate            src/utils/stats.h:64-112
"""
import numpy as np


def quat_multiply(q, p):
    """L(q) p with L = [[q4 I - skew(qv), qv], [-qv^T, q4]] (quat_ops.h:181-196)."""
    q, p = np.asarray(q, float), np.asarray(p, float)
    qv, q4 = q[:3], q[3]
    sk = np.array([[0, -qv[2], qv[1]], [qv[2], 0, -qv[0]], [-qv[1], qv[0], 0]])
    L = np.zeros((4, 4))
    L[:3, :3] = q4 * np.eye(3) - sk
    L[:3, 3] = qv
    L[3, :3] = -qv
    L[3, 3] = q4
    o = L @ p
    if o[3] < 0:
        o = -o
    return o / np.linalg.norm(o)


def quat_inv(q):
    q = np.asarray(q, float)
    return np.array([-q[0], -q[1], -q[2], q[3]])


def trajectory_errors(x_est16, truth10):
    """Per-state (ori deg, pos m, vel m/s) of estimated states [n][16] = (q4, bg3, v3, ba3, p3) against truth [n][10] =
    (q_VtoI 4, p_IinV 3, v_IinV 3): run_simulation.cpp:216-218."""
    n = len(x_est16)
    out = np.zeros((n, 3))
    for i in range(n):
        qe = quat_multiply(truth10[i, 0:4], quat_inv(x_est16[i, 0:4]))
        out[i, 0] = 180.0 / np.pi * 2.0 * np.linalg.norm(qe[:3])
        out[i, 1] = np.linalg.norm(x_est16[i, 13:16] - truth10[i, 4:7])
        out[i, 2] = np.linalg.norm(x_est16[i, 7:10] - truth10[i, 7:10])
    return out


def stats_calculate(values):
    """Stats::calculate (stats.h:64-112) -> (rmse, mean, median, min, max, std, ninetynine, count); sums in sorted order."""
    v = np.sort(np.asarray(values, float))
    n = len(v)
    if n == 0:
        return np.array([np.nan] * 7 + [0.0])
    if n == 1:
        median = v[0]
    elif n % 2 == 1:
        median = v[n // 2]
    else:
        median = 0.5 * (v[n // 2 - 1] + v[n // 2])
    mean = 0.0
    rmse = 0.0
    for x in v:
        mean += x
        rmse += x * x
    mean /= n
    rmse = np.sqrt(rmse / n)
    sd = 0.0
    for x in v:
        sd += (x - mean) ** 2
    with np.errstate(invalid="ignore", divide="ignore"):
        sd = np.sqrt(np.float64(sd) / np.float64(n - 1))
    return np.array([rmse, mean, median, v[0], v[-1], sd, mean + 2.326 * sd, float(n)])
