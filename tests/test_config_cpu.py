"""flvis_b200/config.py: the launch-file YAML reader builds the same tracker configuration TrackingNodeletClass::onInit builds
(vo_tracking.cpp:105-306).  The YAML texts below are written for the test (same keys and comment style as the reference's
launch files, public EuRoC / D435 calibration values); the expected structs come from the set-up the parity tests use."""
import glob
import os

import numpy as np
import pytest

from flvis_b200 import batch, config
from synthdata import sequences


def _arr(x):
    return np.array(x[:])


EUROC_YAML = """
#type_of_vi: 1 = euroc mav dataset
type_of_vi: 1
image_width: 752
image_height: 480
T_imu_mavimu:
[ 0.0,  0.0,  1.0,  0.0,
  0.0, -1.0,  0.0,  0.0,
  1.0,  0.0,  0.0,  0.0,
  0.0,  0.0,  0.0,  1.0]
cam0_intrinsics: [{K0}]#fx fy cx cy
cam0_distortion_coeffs: [{D0}]#k1 k2 r1 r2
T_mavimu_cam0:
[{T0}]
cam1_intrinsics: [{K1}]#fx fy cx cy
cam1_distortion_coeffs: [{D1}]#k1 k2 r1 r2
T_mavimu_cam1:
[{T1}]
vifusion_para1: 0.1
vifusion_para2: 0.01
vifusion_para3: 0.001
vifusion_para4: 0.001
vifusion_para5: 0.3
vifusion_para6: 0.1
feature_para1: 30
feature_para2: 20
feature_para3: 5
feature_para4: 1000
feature_para5: 0.01
feature_para6: 10
dr_para1: 0.90
dr_para2: 50
dr_para3: 1.0
window_size:       10
"""


def test_euroc_yaml_gives_the_stereo_config_of_the_parity_tests():
    c = sequences.EUROC
    j = lambda v: ", ".join(repr(float(x)) for x in v)
    text = EUROC_YAML.format(K0=j(c["K0"]), D0=j(c["D0"]), K1=j(c["K1"]), D1=j(c["D1"]), T0=j(c["T_mavimu_cam0"]), T1=j(c["T_mavimu_cam1"]))
    y = config.load_yaml(text)
    s = config.tracker_setup(y)
    assert s["kind"] == "stereo" and s["window"] == 10 and s["has_imu"]
    got = s["cfg"]
    ref = batch.stereo_config_for(sequences.make_c1(2))
    assert (got.cam_type, got.img_w, got.img_h, got.skip_first_n_imgs, got.need_equal_hist) == (2, 752, 480, 0, 1)
    for f in ("K0", "D0", "K1", "D1", "feature_para", "vi_para", "dc_para"):
        assert np.array_equal(_arr(getattr(got, f)), _arr(getattr(ref, f))), f
    for f in ("R0", "P0", "R1", "P1"):                          # cv2.stereoRectify on the same inputs up to the rounding of T_c1_c0
        assert np.abs(_arr(getattr(got, f)) - _arr(getattr(ref, f))).max() < 1e-9 * max(1.0, np.abs(_arr(getattr(ref, f))).max()), f
    for f in ("T_c0_c1", "T_i_c0"):
        a, b = _arr(getattr(got, f)), _arr(getattr(ref, f))
        if a[3] * b[3] < 0:
            a[:4] = -a[:4]
        assert np.abs(a - b).max() < 1e-12, f


def test_depth_yaml_and_window_clamp():
    c = sequences.D435
    text = ("type_of_vi: 0\nimage_width:  640\nimage_height: 480\n"
            f"cam0_intrinsics: [{', '.join(repr(float(v)) for v in c['K0'])}]#fx fy cx cy\n"
            "cam0_distortion_coeffs: [0.0, 0.0, 0.0, 0.0]#k1 k2 r1 r2\ndepth_factor: 1000.0\n"
            f"T_imu_cam0:\n[{', '.join(repr(float(v)) for v in c['T_imu_cam0'])}]\n"
            + "".join(f"vifusion_para{i + 1}: {v}\n" for i, v in enumerate(c["vi_para"]))
            + "".join(f"feature_para{i + 1}: {v}\n" for i, v in enumerate(c["feature_para"]))
            + "".join(f"dr_para{i + 1}: {v}\n" for i, v in enumerate(c["dc_para"])) + "window_size: 250\n")
    s = config.tracker_setup(config.load_yaml(text))
    ref, _, _, _, _ = batch.config_for(sequences.make_c0(2))
    got = s["cfg"]
    assert s["kind"] == "depth" and s["window"] == 100                      # vo_localmap.cpp:441-447 clamps to [3, 100]
    assert (got.cam_type, got.img_w, got.img_h, got.skip_first_n_imgs, got.depth_scale) == (0, 640, 480, 50, 1000.0)
    for f in ("cam0", "feature_para", "vi_para", "dc_para"):
        assert np.array_equal(_arr(getattr(got, f)), _arr(getattr(ref, f))), f
    a, b = _arr(got.T_i_c0), _arr(ref.T_i_c0)
    if a[3] * b[3] < 0:
        a[:4] = -a[:4]
    assert np.abs(a - b).max() < 1e-12


@pytest.mark.skipif(not os.path.isdir("/root/reference/launch"), reason="reference launch files only exist in the build container")
def test_every_reference_launch_yaml_parses():
    n = 0
    for path in sorted(glob.glob("/root/reference/launch/*/*.yaml")):
        try:
            y = config.load_yaml(path)
        except ValueError:
            continue                                                          # px4 plugin lists etc.: not a FLVIS configuration
        s = config.tracker_setup(y)
        assert s["kind"] in ("depth", "stereo") and 3 <= s["window"] <= 100
        n += 1
    assert n >= 5
