"""The host stand-ins of the two RANSAC calls (flvis_b200/host/ransac.cpp, selectable instead of the K11 device kernels)
against cv2 on synthetic two-view / PnP scenes: statistical parity (same bars as tests/test_ransac_gpu.py)."""
import ctypes as C

import cv2
import numpy as np

K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
K4 = np.array([458.654, 457.296, 367.215, 248.375])


def _scene(rng, n, out_frac, noise=0.5):
    X = np.stack([rng.uniform(-4, 4, n), rng.uniform(-3, 3, n), rng.uniform(3, 12, n)], 1)
    rvec = rng.normal(0, 0.03, 3); t = np.array([0.15, 0.02, 0.05]) + rng.normal(0, 0.02, 3)
    R, _ = cv2.Rodrigues(rvec)
    proj = lambda Xc: (K @ (Xc / Xc[:, 2:3]).T).T[:, :2]
    u1 = proj(X) + rng.normal(0, noise, (n, 2))
    u2 = proj((R @ X.T).T + t) + rng.normal(0, noise, (n, 2))
    bad = rng.uniform(size=n) < out_frac
    u2[bad] += rng.uniform(-60, 60, (int(bad.sum()), 2)) + 15 * np.sign(rng.normal(size=(int(bad.sum()), 2)))
    return X.astype(np.float32), R, t, u1.astype(np.float32), u2.astype(np.float32), bad


def _p(a):
    return np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


def test_host_fundamental_ransac(lib):
    lib.flv_host_fundamental_ransac.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    rng = np.random.default_rng(4)
    for n, of in ((400, 0.25), (150, 0.2)):
        X, R, t, u1, u2, bad = _scene(rng, n, of)
        mask = np.zeros(n, np.uint8); F = np.zeros(9)
        assert lib.flv_host_fundamental_ransac(n, _p(u1), _p(u2), 5.0, 0.99, _p(mask), _p(F)) == 1
        _, mc = cv2.findFundamentalMat(u1, u2, cv2.FM_RANSAC, 5.0, 0.99)
        mc = mc.ravel().astype(bool); m = mask.astype(bool)
        # adaptive early exit + mask of the sample model (no local optimisation): a little below K11 / cv2
        assert m[~bad].mean() > 0.92 and m[bad].mean() <= max(0.2, mc[bad].mean() + 0.15)
        assert (m & mc).sum() / (m | mc).sum() > 0.88
    mask = np.zeros(5, np.uint8); F = np.zeros(9)
    assert lib.flv_host_fundamental_ransac(5, _p(u1[:5]), _p(u2[:5]), 5.0, 0.99, _p(mask), _p(F)) == 0 and not mask.any()


def test_host_pnp_ransac(lib):
    lib.flv_host_pnp_ransac.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                        C.c_void_p]
    rng = np.random.default_rng(6)
    X, R, t, u1, u2, bad = _scene(rng, 300, 0.2)
    w = np.sqrt(max(0.0, 1 + R[0, 0] + R[1, 1] + R[2, 2])) / 2
    q = np.array([(R[2, 1] - R[1, 2]) / (4 * w), (R[0, 2] - R[2, 0]) / (4 * w), (R[1, 0] - R[0, 1]) / (4 * w), w])
    T = np.r_[q, t + rng.normal(0, 0.02, 3)]                      # prior: true rotation, 2 cm off
    mask = np.zeros(300, np.uint8)
    ninl = lib.flv_host_pnp_ransac(300, _p(X), _p(u2), _p(K4), _p(T), 100, 3.0, 0.99, _p(mask))
    m = mask.astype(bool)
    assert ninl == m.sum() and ninl > 0.7 * (~bad).sum()
    assert m[~bad].mean() > 0.95 and m[bad].mean() < 0.1
    assert np.linalg.norm(T[4:] - t) < 0.02
