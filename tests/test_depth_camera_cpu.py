"""DepthCamera::setSteroCamInfo (src/processing/depth_camera.cpp:27-90) through the C ABI: the intrinsics the tracker uses come
from the rectified projections P0 / P1 (:73-82), T_cam1_cam0 = T_c0_c1^-1 (:60-61).  Input matrices = what the tracking nodelet
passes for euroc.yaml (vo_tracking.cpp:222-262, cv::stereoRectify through cv2)."""
import ctypes as C

import numpy as np

from flvis_b200 import batch, capi
from synthdata import sequences


def test_stereo_cam_info_matches_nodelet_setup():
    lib = capi.load_library()
    lib.flv_host_stereo_cam_info.argtypes = [C.POINTER(batch.F2FStereoConfig), C.c_void_p, C.c_void_p, C.c_void_p]
    seq = sequences.make_c1(2)
    cfg, lenses, equalize, K, m = batch.config_for(seq)
    sc = batch.stereo_config_for(seq)
    cam0 = np.zeros(4); cam1 = np.zeros(4); T10 = np.zeros(7)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.flv_host_stereo_cam_info(C.byref(sc), vp(cam0), vp(cam1), vp(T10)) == 0
    assert np.array_equal(cam0, np.array(K)) and np.array_equal(cam0, np.array(cfg.cam0[:]))
    assert np.array_equal(cam1, np.array(cfg.cam1[:]))
    assert sc.need_equal_hist == 1 and equalize
    ref = np.array(cfg.T_cam1_cam0[:])
    if ref[3] * T10[3] < 0:
        ref[:4] = -ref[:4]
    assert np.abs(T10 - ref).max() < 1e-12
    # rectified stereo: same fx, fy, cy on both sides and a pure x baseline in P1 (CALIB_ZERO_DISPARITY)
    assert cam0[0] == cam1[0] and cam0[3] == cam1[3] and cam0[2] == cam1[2] and m["P1"][0, 3] < 0
