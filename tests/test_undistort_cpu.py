"""The STEREO_UNRECT per-point lens maps (flvis_b200/host/undistort.h) against the OpenCV calls the reference makes
(cv::undistortPoints / cv::projectPoints, src/processing/lkorb_tracking.cpp:59,87; camera_frame.cpp:116,130)."""
import ctypes as C

import cv2
import numpy as np
import pytest

# EuRoC cam0 / cam1 (launch/EuRoC_MAV/euroc.yaml: radtan k1 k2 p1 p2) and a rectification like cv::stereoRectify's
K0 = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
D0 = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05])
K1 = np.array([[457.587, 0, 379.999], [0, 456.134, 255.238], [0, 0, 1.0]])
D1 = np.array([-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05])


def _rect(k):
    r, _ = cv2.Rodrigues(np.array([0.003, -0.007, 0.002]) * k)
    P = np.array([[435.2, 0, 367.4, 0], [0, 435.2, 252.2, 0], [0, 0, 1, 0.0]])
    if k < 0:
        P[0, 3] = -47.9
    return r, P


def _d14(D):
    out = np.zeros(14)
    out[:len(D)] = D
    return out


def _p(a):
    return np.ascontiguousarray(a).ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize("K,D,k", [(K0, D0, 1), (K1, D1, -1), (K0, np.array([-0.3, 0.1, 1e-3, -2e-3, -0.02, 0.01, 0.002, -0.001]), 1),
                                   (K0, np.zeros(4), 0)])
def test_undistort_points_matches_cv2(lib, K, D, k):
    R, P = _rect(k)
    if k == 0:
        R = np.eye(3)
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(0, 752, 4000), rng.uniform(0, 480, 4000)], 1).astype(np.float32)
    ref = cv2.undistortPoints(pts.reshape(-1, 1, 2), K, D, R=R, P=P).reshape(-1, 2)
    out = np.zeros_like(pts)
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    lib.flv_host_undistort_points.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    assert lib.flv_host_undistort_points(_p(K4), _p(_d14(D)), _p(R.ravel()), _p(P.ravel()), len(pts), _p(pts), _p(out)) == 0
    # float outputs of the same double arithmetic: allow one float ulp where the double result sits on a rounding boundary
    assert np.abs(out - ref).max() <= 6.2e-5
    assert (out == ref).mean() > 0.999


@pytest.mark.parametrize("K,D", [(K0, D0), (K1, D1), (K0, np.array([-0.3, 0.1, 1e-3, -2e-3, -0.02, 0.01, 0.002, -0.001]))])
def test_project_points_matches_cv2(lib, K, D):
    rng = np.random.default_rng(5)
    X = np.stack([rng.uniform(-2, 2, 3000), rng.uniform(-1.5, 1.5, 3000), rng.uniform(1.0, 9, 3000)], 1).astype(np.float32)
    rvec = np.array([0.02, -0.05, 0.01]); t = np.array([0.11, -0.02, 0.05])
    R, _ = cv2.Rodrigues(rvec)
    ref, _ = cv2.projectPoints(X.reshape(-1, 1, 3).astype(np.float64), rvec, t, K, D)
    ref = ref.reshape(-1, 2).astype(np.float32)
    out = np.zeros((len(X), 2), np.float32)
    K4 = np.array([K[0, 0], K[1, 1], K[0, 2], K[1, 2]])
    lib.flv_host_project_points.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_void_p]
    assert lib.flv_host_project_points(_p(K4), _p(_d14(D)), _p(R.ravel()), _p(t), len(X), _p(X), _p(out)) == 0
    assert np.abs(out - ref).max() <= 1.3e-4        # rvec -> R round trip inside cv2 (1e-16) + float output rounding
    assert (out == ref).mean() > 0.99
