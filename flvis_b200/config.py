"""FLVIS launch-file YAML -> tracker / local-map configuration (the key reader of the tracking nodelet).

Mirrors what TrackingNodeletClass::onInit does with its configFilePath (src/frontend/vo_tracking.cpp:105-306: the
get*VariableFromYaml helpers of src/utils/include/yamlRead.h) and LocalMapNodeletClass::onInit (src/backend/vo_localmap.cpp:
436-452: window_size clamped to [3, 100]):

  type_of_vi 0 / 2  D435(i) depth (+ pixhawk)  -> DEPTH_D435, skip the first 50 images, flv_f2f_config
  type_of_vi 3 / 5  D435(i) infrared stereo    -> STEREO_RECT through cv::stereoRectify, skip 50, flv_f2f_stereo_config
  type_of_vi 1      EuRoC MAV                  -> STEREO_UNRECT, T_c0_c1 = T_mavimu_cam0^-1 T_mavimu_cam1, T_i_c0 = T_imu_mavimu T_mavimu_cam0,
                                                  equalizeHist on, flv_f2f_stereo_config
  type_of_vi 4      KITTI stereo               -> STEREO_RECT from the two projection matrices (baseline = K^-1 P1), no IMU

cv::stereoRectify is called through cv2 exactly as the nodelet calls it (CALIB_ZERO_DISPARITY, alpha 0, same image size): node
initialisation, outside the hot path.  Poses are [qx qy qz qw tx ty tz].
"""
import ctypes as C
import math
import re

import numpy as np

from . import batch

TYPE_DEPTH = (0, 2)
TYPE_STEREO_D435 = (3, 5)
TYPE_EUROC = 1
TYPE_KITTI = 4


def load_yaml(path_or_text):
    """Parse a FLVIS yaml (path or text).  The launch files put comments directly behind values (`[...]#fx fy cx cy`), which strict
    YAML does not accept: a space is inserted in front of every such `#`."""
    import yaml
    text = path_or_text
    if "\n" not in path_or_text and ":" not in path_or_text.split("/")[-1]:
        with open(path_or_text) as f:
            text = f.read()
    text = re.sub(r"(?<=[^\s#])#", " #", text)
    # matrices are written as a flow sequence that starts on the line AFTER its key and continues at any indentation (yaml-cpp
    # accepts that, PyYAML does not): pull the sequence up to its key and onto one line
    text = re.sub(r"\[[^\]]*\]", lambda m: " ".join(m.group(0).split()), text)
    text = re.sub(r":[ \t]*\n[ \t]*\[", ": [", text)
    y = yaml.safe_load(text)
    if not isinstance(y, dict) or "type_of_vi" not in y:
        raise ValueError("not a FLVIS configuration: type_of_vi is missing")
    return y


def _mat44(y, key):
    m = np.asarray(y[key], np.float64).reshape(4, 4)
    return m


def _pose7(m):
    """4x4 homogeneous matrix -> [qx qy qz qw tx ty tz] (Eigen's / Sophus' rotation-matrix-to-quaternion branches)."""
    R = m[:3, :3]
    t = R[0, 0] + R[1, 1] + R[2, 2]
    q = np.zeros(4)                                            # w x y z
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q[:] = (0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s)
    else:
        i = int(np.argmax([R[0, 0], R[1, 1], R[2, 2]])); j = (i + 1) % 3; k = (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q[1 + i] = 0.25 * s; q[0] = (R[k, j] - R[j, k]) / s; q[1 + j] = (R[j, i] + R[i, j]) / s; q[1 + k] = (R[k, i] + R[i, k]) / s
    q /= np.linalg.norm(q)
    return np.array([q[1], q[2], q[3], q[0], m[0, 3], m[1, 3], m[2, 3]])


def _K(intr):
    fx, fy, cx, cy = [float(v) for v in intr]
    return np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])


def _d(vals, n):
    a = [float(v) for v in np.asarray(vals, np.float64).ravel()]
    return (C.c_double * n)(*(a + [0.0] * (n - len(a))))


def _paras(y):
    return ([y[f"feature_para{i}"] for i in range(1, 7)], [y[f"vifusion_para{i}"] for i in range(1, 7)],
            [y[f"dr_para{i}"] for i in range(1, 4)])


def window_size(y):
    """LocalMapNodeletClass::onInit (vo_localmap.cpp:441-447): window_size clamped to [3, 100]."""
    return int(min(100, max(3, int(y.get("window_size", 8)))))


def tracker_setup(y):
    """-> dict(kind = "depth" | "stereo", cfg = batch.F2FConfig | batch.F2FStereoConfig, window, has_imu).
    kind "depth": pass cfg to flv_f2f_create / flv_f2f_batch_create; kind "stereo": to flv_f2f_create_stereo."""
    vi = int(y["type_of_vi"])
    w, h = int(y["image_width"]), int(y["image_height"])
    fpar, vpar, dpar = _paras(y)
    if vi in TYPE_DEPTH:
        K = [float(v) for v in y["cam0_intrinsics"]]
        cfg = batch.F2FConfig(0, w, h, _d(K, 4), _d(K, 4), float(y["depth_factor"]), (C.c_double * 12)(), (C.c_double * 12)(),
                              _d([0, 0, 0, 1, 0, 0, 0], 7), _d(_pose7(_mat44(y, "T_imu_cam0")), 7), _d(fpar, 6), _d(vpar, 6), _d(dpar, 3), 50)
        return dict(kind="depth", cfg=cfg, window=window_size(y), has_imu=True)
    if vi in TYPE_STEREO_D435 or vi == TYPE_EUROC:
        import cv2
        K0, K1 = _K(y["cam0_intrinsics"]), _K(y["cam1_intrinsics"])
        D0 = np.asarray(y["cam0_distortion_coeffs"], np.float64); D1 = np.asarray(y["cam1_distortion_coeffs"], np.float64)
        if vi == TYPE_EUROC:
            T_mavi_c0, T_mavi_c1, T_i_mavi = _mat44(y, "T_mavimu_cam0"), _mat44(y, "T_mavimu_cam1"), _mat44(y, "T_imu_mavimu")
            T_c0_c1 = np.linalg.inv(T_mavi_c0) @ T_mavi_c1
            T_i_c0 = T_i_mavi @ T_mavi_c0
            cam_type, skip, equalize = 2, 0, 1
        else:
            T_c0_c1, T_i_c0 = _mat44(y, "T_cam0_cam1"), _mat44(y, "T_imu_cam0")
            cam_type, skip, equalize = 1, 50, 0
        T_c1_c0 = np.linalg.inv(T_c0_c1)
        R0, R1, P0, P1, _, _, _ = cv2.stereoRectify(K0, D0, K1, D1, (w, h), np.ascontiguousarray(T_c1_c0[:3, :3]),
                                                    np.ascontiguousarray(T_c1_c0[:3, 3]).reshape(3, 1), flags=cv2.CALIB_ZERO_DISPARITY,
                                                    alpha=0, newImageSize=(w, h))
        cfg = batch.F2FStereoConfig(cam_type, w, h, _d(K0, 9), _d(D0, 14), _d(R0, 9), _d(P0, 12), _d(K1, 9), _d(D1, 14), _d(R1, 9), _d(P1, 12),
                                    _d(_pose7(T_c0_c1), 7), _d(_pose7(T_i_c0), 7), _d(fpar, 6), _d(vpar, 6), _d(dpar, 3), skip, equalize)
        return dict(kind="stereo", cfg=cfg, window=window_size(y), has_imu=True)
    if vi == TYPE_KITTI:
        P0 = _mat44(y, "cam0_projection_matrix"); P1 = _mat44(y, "cam1_projection_matrix")
        Kinv = np.zeros((4, 4)); Kinv[:3, :3] = np.linalg.inv(P0[:3, :3])
        T = Kinv @ P1
        T[:3, :3] = np.eye(3); T[3] = (0, 0, 0, 1)
        K = P0[:3, :3]
        eye = np.eye(3)
        cfg = batch.F2FStereoConfig(1, w, h, _d(K, 9), _d([], 14), _d(eye, 9), _d(P0[:3, :4], 12), _d(K, 9), _d([], 14), _d(eye, 9),
                                    _d(P1[:3, :4], 12), _d(_pose7(T), 7), _d([0, 0, 0, 1, 0, 0, 0], 7), _d(fpar, 6), _d(vpar, 6), _d(dpar, 3), 0, 0)
        return dict(kind="stereo", cfg=cfg, window=window_size(y), has_imu=False)
    raise ValueError(f"type_of_vi {vi}: unknown sensor set-up")
