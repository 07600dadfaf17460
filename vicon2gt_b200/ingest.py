"""Non-ROS ingestion of recorded data: what src/estimate_vicon2gt.cpp reads from a rosbag, from plain
CSV / text files instead.

  IMU    `t, wx, wy, wz, ax, ay, az`                       (EuRoC imu0/data.csv; sensor_msgs/Imu, estimate_vicon2gt.cpp:172-180)
  vicon  `t, px, py, pz, qw, qx, qy, qz`                   (EuRoC vicon0/data.csv: position first, scalar-first quaternion)
         `t tx ty tz qx qy qz qw`                          (TUM / PoseStamped / TransformStamped text, :222-262)
         `t, px, py, pz, qx, qy, qz, qw, cov[36]`          (nav_msgs/Odometry: pose + 6x6 covariance, order x y z rx ry rz, :183-219)
Time stamps above 1e12 are nanoseconds.  Lines starting with '#' are comments; ',' or blanks separate the fields.
The state times are the fixed-rate grid of estimate_vicon2gt.cpp:266-276 (first vicon stamp, `t += 1 / state_freq`
while `t <` last vicon stamp).  The result is a `sim.Dataset` that `BatchSolver.set_trial_dataset` accepts.
"""
import argparse
import sys

import numpy as np

from .sim import Dataset

DEFAULT_VICON_SIGMAS = (1e-4, 1e-4, 1e-4, 1e-5, 1e-5, 1e-5)  # estimate_vicon2gt.cpp:126


def _read_table(path):
    rows = []
    with open(path) as f:
        for line in f:
            line = line.strip()
            if not line or line.startswith("#"):
                continue
            rows.append([float(x) for x in line.replace(",", " ").split()])
    if not rows:
        return np.zeros((0, 0))
    w = min(len(r) for r in rows)
    return np.array([r[:w] for r in rows], dtype=np.float64)


def _seconds(t):
    t = np.asarray(t, dtype=np.float64)
    return t * 1e-9 if (t.size and np.max(np.abs(t)) > 1e12) else t


def load_imu_csv(path):
    """-> t[M], w[M][3], a[M][3] in file order (Propagator::feed_imu appends, it never sorts)."""
    a = _read_table(path)
    if a.shape[1] < 7:
        raise ValueError(f"{path}: an IMU row needs 7 fields (t, w xyz, a xyz)")
    return _seconds(a[:, 0]), np.ascontiguousarray(a[:, 1:4]), np.ascontiguousarray(a[:, 4:7])


def load_vicon_csv(path, layout="auto"):
    """-> t[V], q[V][4] (JPL x y z w, as the reference stores the message's quaternion), p[V][3], cov[V][36] or None.

    layout: "euroc" (t, p, qw qx qy qz), "tum" (t, p, qx qy qz qw), "odom" (tum + 36 covariance entries), "auto": by width
    (44 -> odom, 8 -> euroc for ',' separated files whose header names q_w before q_x, else tum)."""
    a = _read_table(path)
    if a.shape[1] < 8:
        raise ValueError(f"{path}: a pose row needs 8 fields")
    if layout == "auto":
        if a.shape[1] >= 44:
            layout = "odom"
        else:
            with open(path) as f:
                head = f.readline().lower()
            layout = "euroc" if ("q_rs_w" in head or "qw,qx" in head.replace(" ", "") or "q_w" in head) else "tum"
    t, p = _seconds(a[:, 0]), np.ascontiguousarray(a[:, 1:4])
    if layout == "euroc":
        q = np.ascontiguousarray(a[:, [5, 6, 7, 4]])
        cov = None
    elif layout in ("tum", "odom"):
        q = np.ascontiguousarray(a[:, 4:8])
        cov = np.ascontiguousarray(a[:, 8:44]) if layout == "odom" else None
    else:
        raise ValueError(layout)
    return t, q, p, cov


def odom_cov_blocks(cov36):
    """R_q[V][9], R_p[V][9] from Odometry covariances exactly as estimate_vicon2gt.cpp:190-204 unpacks them:
    pose_cov(r, c) = covariance[6 c + r]; R_q = block(3, 3), R_p = block(0, 0)."""
    c = np.asarray(cov36, dtype=np.float64).reshape(-1, 6, 6).transpose(0, 2, 1)  # (r, c) <- [6 c + r]
    return np.ascontiguousarray(c[:, 3:6, 3:6]).reshape(-1, 9), np.ascontiguousarray(c[:, 0:3, 0:3]).reshape(-1, 9)


def make_state_times(vic_t, state_freq):
    """estimate_vicon2gt.cpp:266-276: fixed-rate grid between the first and the last vicon stamp IN FEED ORDER."""
    if len(vic_t) == 0:
        return np.zeros(0)
    start, end = float(vic_t[0]), float(vic_t[-1])
    out = []
    if start < end:
        t, step = start, 1.0 / float(state_freq)
        while t < end:
            out.append(t)
            t += step
    return np.array(out, dtype=np.float64)


def dataset_from_files(imu_csv, vicon_csv, state_freq=100, vicon_sigmas=DEFAULT_VICON_SIGMAS, use_manual_sigmas=False,
                       bag_start=0.0, bag_durr=-1.0, vicon_layout="auto"):
    """The stores estimate_vicon2gt.cpp fills from a bag (:140-301), from two files."""
    it, iw, ia = load_imu_csv(imu_csv)
    vt, vq, vp, cov = load_vicon_csv(vicon_csv, vicon_layout)
    # time window relative to the first message of either stream (:91-96)
    t_init = min(it[0] if len(it) else np.inf, vt[0] if len(vt) else np.inf) + bag_start
    t_fin = max(it[-1] if len(it) else -np.inf, vt[-1] if len(vt) else -np.inf) if bag_durr < 0 else t_init + bag_durr
    mi, mv = (it >= t_init) & (it <= t_fin), (vt >= t_init) & (vt <= t_fin)
    it, iw, ia = it[mi], iw[mi], ia[mi]
    vt, vq, vp = vt[mv], vq[mv], vp[mv]
    s = np.asarray(vicon_sigmas, dtype=np.float64)
    R_const = np.zeros(18)
    R_const[[0, 4, 8]] = s[:3] ** 2
    R_const[[9, 13, 17]] = s[3:] ** 2
    ds = Dataset(it, iw, ia, vt, vq, vp, R_const, make_state_times(vt, state_freq))
    if cov is not None and not use_manual_sigmas:
        ds.vic_Rq, ds.vic_Rp = odom_cov_blocks(cov[mv])
    return ds


def main(argv=None):
    """Non-ROS `estimate_vicon2gt`: files in, groundtruth CSV + calibration info out (needs a CUDA device)."""
    ap = argparse.ArgumentParser(description=main.__doc__)
    ap.add_argument("--imu", required=True)
    ap.add_argument("--vicon", required=True)
    ap.add_argument("--states-csv", required=True, help="stats_path_states of the reference's launch files")
    ap.add_argument("--info-txt", required=True, help="stats_path_info")
    ap.add_argument("--state-freq", type=int, default=100)
    ap.add_argument("--vicon-sigmas", type=float, nargs=6, default=list(DEFAULT_VICON_SIGMAS))
    ap.add_argument("--use-manual-sigmas", action="store_true")
    ap.add_argument("--bag-start", type=float, default=0.0)
    ap.add_argument("--bag-durr", type=float, default=-1.0)
    ap.add_argument("--num-loop-relin", type=int, default=0)
    ap.add_argument("--device", type=int, default=0)
    a = ap.parse_args(argv)
    from . import api

    ds = dataset_from_files(a.imu, a.vicon, a.state_freq, a.vicon_sigmas, a.use_manual_sigmas, a.bag_start, a.bag_durr)
    if ds.n_imu == 0 or ds.n_vicon == 0 or ds.n_times == 0:
        print("Not enough data to optimize with!", file=sys.stderr)  # estimate_vicon2gt.cpp:285-289
        return 1
    p = api.default_params(num_loop_relin=a.num_loop_relin)
    with api.BatchSolver(1, device=a.device) as bs:
        bs.set_trial_dataset(0, p, ds)
        bs.build_and_solve()
        inf = bs.info(0)
        if inf.status != 0:
            print(f"build_and_solve failed with status {inf.status}", file=sys.stderr)
            return 1
        bs.write_to_file(0, a.states_csv, a.info_txt)
        print(f"{inf.n_states} states, {inf.iterations} LM iterations, final error {inf.final_error:.6e}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
