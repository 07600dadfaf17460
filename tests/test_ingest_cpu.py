"""Non-ROS ingestion of recorded data (src/estimate_vicon2gt.cpp:140-301): file layouts, the Odometry
covariance unpacking, the time window and the fixed-rate state grid."""
import numpy as np
import pytest

from vicon2gt_b200 import ingest


def _write_euroc(tmp_path, ds):
    imu = tmp_path / "imu0.csv"
    vic = tmp_path / "vicon0.csv"
    with open(imu, "w") as f:
        f.write("#timestamp [ns],w_RS_S_x [rad s^-1],w_RS_S_y,w_RS_S_z,a_RS_S_x [m s^-2],a_RS_S_y,a_RS_S_z\n")
        for t, w, a in zip(ds.imu_t, ds.imu_w, ds.imu_a):
            f.write("%d,%s\n" % (int(round(t * 1e9)), ",".join(repr(float(x)) for x in (*w, *a))))
    with open(vic, "w") as f:
        f.write("#timestamp [ns], p_RS_R_x [m], p_RS_R_y [m], p_RS_R_z [m], q_RS_w [], q_RS_x [], q_RS_y [], q_RS_z []\n")
        for t, q, p in zip(ds.vic_t, ds.vic_q, ds.vic_p):
            f.write("%d,%s\n" % (int(round(t * 1e9)), ",".join(repr(float(x)) for x in (*p, q[3], q[0], q[1], q[2]))))
    return str(imu), str(vic)


def test_euroc_layout_round_trip(tmp_path, ds_c2_small):
    ds = ds_c2_small
    imu, vic = _write_euroc(tmp_path, ds)
    d2 = ingest.dataset_from_files(imu, vic, state_freq=20, vicon_sigmas=(1e-3,) * 3 + (1e-2,) * 3)
    assert np.allclose(d2.imu_t, ds.imu_t, rtol=0, atol=1e-6) and np.array_equal(d2.imu_w, ds.imu_w) and np.array_equal(d2.imu_a, ds.imu_a)
    assert np.array_equal(d2.vic_q, ds.vic_q) and np.array_equal(d2.vic_p, ds.vic_p)  # scalar-first file -> JPL x y z w
    assert d2.vic_Rq is None and np.allclose(d2.vic_R_const[[0, 4, 8]], 1e-6) and np.allclose(d2.vic_R_const[[9, 13, 17]], 1e-4)
    # state grid (estimate_vicon2gt.cpp:266-276): sequential t += 1 / freq from the first vicon stamp, strictly below the last
    st = d2.state_t
    assert st[0] == d2.vic_t[0] and st[-1] < d2.vic_t[-1] <= st[-1] + 1.0 / 20 + 1e-9
    t, ref = d2.vic_t[0], []
    while t < d2.vic_t[-1]:
        ref.append(t)
        t += 1.0 / 20.0
    assert np.array_equal(st, np.array(ref))


def test_odometry_covariance_and_window(tmp_path):
    rng = np.random.default_rng(0)
    V = 50
    t = 100.0 + 0.01 * np.arange(V)
    q = rng.normal(size=(V, 4))
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    p = rng.normal(size=(V, 3))
    cov = rng.normal(size=(V, 6, 6))  # deliberately NOT symmetric: the unpacking order is observable
    vic = tmp_path / "odom.txt"
    with open(vic, "w") as f:
        f.write("# t px py pz qx qy qz qw cov(36, row-major, order x y z rx ry rz)\n")
        for i in range(V):
            f.write(" ".join(repr(float(x)) for x in (t[i], *p[i], *q[i], *cov[i].reshape(-1))) + "\n")
    imu = tmp_path / "imu.txt"
    with open(imu, "w") as f:
        for ti in 99.9 + 0.0025 * np.arange(300):
            f.write(" ".join(repr(float(x)) for x in (ti, 0.1, 0.2, 0.3, 0.0, 0.0, 9.81)) + "\n")
    ds = ingest.dataset_from_files(str(imu), str(vic), state_freq=100)
    assert ds.vic_Rq.shape == (V, 9) and np.array_equal(ds.vic_q, q) and np.array_equal(ds.vic_p, p)
    for i in (0, 7, V - 1):  # pose_cov(r, c) = covariance[6 c + r]; R_q = block(3, 3), R_p = block(0, 0)
        pc = cov[i].reshape(-1)
        Rq = np.array([[pc[6 * c + r] for c in range(3, 6)] for r in range(3, 6)])
        Rp = np.array([[pc[6 * c + r] for c in range(0, 3)] for r in range(0, 3)])
        assert np.array_equal(ds.vic_Rq[i].reshape(3, 3), Rq) and np.array_equal(ds.vic_Rp[i].reshape(3, 3), Rp)
    # manual sigmas override the message covariances; the window is relative to the first message of either stream
    d2 = ingest.dataset_from_files(str(imu), str(vic), use_manual_sigmas=True, bag_start=0.2, bag_durr=0.25)
    assert d2.vic_Rq is None
    assert d2.imu_t[0] >= 99.9 + 0.2 - 1e-12 and d2.imu_t[-1] <= 99.9 + 0.45 + 1e-12
    assert d2.vic_t[0] >= 100.1 - 1e-9 and d2.vic_t[-1] <= 100.35 + 1e-9
    with pytest.raises(ValueError):
        ingest.load_imu_csv(str(vic).replace("odom", "missing") if False else _short(tmp_path))


def _short(tmp_path):
    pth = tmp_path / "short.csv"
    pth.write_text("1.0,2.0,3.0\n")
    return str(pth)
