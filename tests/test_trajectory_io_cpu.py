"""Trajectory writers (flvis_b200/host/trajectory_io.cpp) against the text the reference's recorder produces
(src/independ_modules/vo_repub_rec.cpp:80-126): C++ stream precision 6 == Python '%.6g', pose = T_c_w^-1, qw first."""
import ctypes as C
import os

import numpy as np

from oracle.vimotion_ref import SE3, q2R


def _bind(lib):
    lib.flv_traj_open.restype = C.c_void_p
    lib.flv_traj_open.argtypes = [C.c_char_p, C.c_int]
    lib.flv_traj_write.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    lib.flv_traj_close.argtypes = [C.c_void_p]


def _poses(n, seed):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        q = rng.normal(size=4); q /= np.linalg.norm(q)
        out.append((1403636579.763555527 + 0.05 * i, SE3(q, rng.normal(0, 3, 3))))
    return out


def test_recorder_pose_lines(lib, tmp_path):
    _bind(lib)
    path = os.path.join(tmp_path, "traj.txt")
    h = lib.flv_traj_open(path.encode(), 0)
    assert h
    poses = _poses(25, 1)
    for t, T in poses:
        a = np.ascontiguousarray(T.to7())
        assert lib.flv_traj_write(h, t, a.ctypes.data_as(C.c_void_p)) == 0
    lib.flv_traj_close(h)
    lines = open(path).read().strip().split("\n")
    assert len(lines) == 25
    for line, (t, T) in zip(lines, poses):
        f = line.split(" ")
        assert len(f) == 8
        sec, nsec = f[0].split(".")
        assert len(nsec) == 9 and abs(int(sec) + int(nsec) * 1e-9 - t) < 1e-6          # ros::Time formatting
        Twc = T.inverse()
        want = [Twc.t[0], Twc.t[1], Twc.t[2], Twc.q[0], Twc.q[1], Twc.q[2], Twc.q[3]]   # x y z qw qx qy qz
        for got, w in zip(f[1:], want):
            assert abs(float(got) - w) <= 5.1e-6 * max(1.0, abs(w))
            assert got == ("%.6g" % float(got))                                        # 6 significant digits, %g style


def test_kitti_rows(lib, tmp_path):
    _bind(lib)
    path = os.path.join(tmp_path, "kitti.txt")
    h = lib.flv_traj_open(path.encode(), 1)
    poses = _poses(10, 2)
    for t, T in poses:
        a = np.ascontiguousarray(T.to7())
        assert lib.flv_traj_write(h, t, a.ctypes.data_as(C.c_void_p)) == 0
    lib.flv_traj_close(h)
    rows = np.loadtxt(path)
    assert rows.shape == (10, 12)
    for r, (t, T) in zip(rows, poses):
        Twc = T.inverse()
        M = np.hstack([q2R(Twc.q), Twc.t.reshape(3, 1)])
        assert np.abs(r.reshape(3, 4) - M).max() <= 5.1e-6 * max(1.0, np.abs(M).max())
    assert lib.flv_traj_open(b"/nonexistent_dir/x.txt", 0) is None and lib.flv_traj_open(path.encode(), 7) is None
