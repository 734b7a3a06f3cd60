"""BASELINE.json's configurations end to end, with the IMU in the loop: flv::F2FTracking (GPU, through the C ABI) against the
restated reference pipeline (oracle/f2f_ref.py), frame by frame, on sequences of the named shape (oracle/sequences.py):

  C0  640x480 D435i depth + IMU 200 Hz, 150 frames of which the first 50 are skipped (vo_tracking.cpp:171)
  C1  752x480 EuRoC raw stereo (euroc.yaml K/D/T, cv::stereoRectify, STEREO_UNRECT, equalizeHist) + IMU, 200 frames,
      10-keyframe local map chained on the tracker's keyframes, two unrelated frames force TrackingFail -> IMU re-init
  C3  1241x376 KITTI-shaped rectified stereo, no IMU, 20-keyframe local map

Covered reference code that no other test reaches: viGetCorrFrameState -> projected initial flow
(src/frontend/f2f_tracking.cpp:225, src/processing/lkorb_tracking.cpp:38-63), viVisionRPCompensation (:253),
viCorrectionFromVision (:281), the UnInit wait for the IMU (:153-175), TrackingFail + re-initialisation from the IMU pose
(:357-394).  Bars: landmark ids / order / flags and LK pixel positions bit-exact, poses <= 1e-6 (1e-5 through cv2's float
undistortPoints), per frame.  With the IMU pose guess in the loop the comparison is teacher-forced (the oracle is re-seeded
from the tracker's state after every frame, tests/seq_harness.py:sync_oracle_from_tracker explains why a free-running
comparison is chaotic in the last bits); a second, free-running oracle gives the trajectory bar: ATE between the two
paths <= 1 % of the path length, and both follow the synthetic ground truth equally well.
The two OpenCV RANSAC calls are injected on both sides (see tests/test_pipeline_gpu.py)."""
import numpy as np
import pytest

from synthdata import sequences

from .seq_harness import run_sequence

pytestmark = pytest.mark.gpu


def _ate_close(r):
    assert r["ate_vs_ref"] <= 0.01 * r["path"]
    if "ate_vs_free_ref" in r:                       # free-running reference path (IMU runs)
        assert r["free_frames_compared"] >= 60 and r["ate_vs_free_ref"] <= 0.01 * r["free_path"]
    else:
        assert abs(r["ate_gt"] - r["ate_gt_ref"]) <= 0.01 * max(r["ate_gt_ref"], 1e-9)


def test_c0_d435_depth_imu_150_frames(lib):
    seq = sequences.make_c0(150)
    r = run_sequence(lib, seq, tol_pose=1e-6, tol_und=0.0, tol_p3=1e-6, lockstep=True, free_frames=150)
    assert r["states"][:50] == ["UnInit"] * 50 and r["final_state"] == "Tracking"     # skip_first_n_imgs
    assert r["frames_tracked"] == 100 and r["has_imu"] == 1
    assert r["guess_used"] >= 95                                                        # IMU pose guess on every tracked frame
    assert np.abs(r["ref_acc_bias"]).max() > 1e-5                                       # bias feedback really ran
    assert np.abs(r["acc_bias"] - r["ref_acc_bias"]).max() <= 1e-9 and np.abs(r["gyro_bias"] - r["ref_gyro_bias"]).max() <= 1e-9
    assert r["path"] > 0.3
    _ate_close(r)
    assert r["ate_gt_ref"] < 0.05                                                       # and the reference path follows the truth


def test_c1_euroc_unrect_imu_200_frames_local_map_and_reinit(lib):
    seq = sequences.make_c1(200, blank_frames=(120, 121))
    r = run_sequence(lib, seq, tol_pose=1e-5, tol_und=6.2e-5, tol_p3=1e-5, window=seq.cfg["window"], lockstep=True, free_frames=120)
    st = r["states"]
    assert st[0] == "UnInit" and "Tracking" in st[:8]                                   # waits for imu_initialized, then inits
    assert st[121] == "TrackingFail" and st[124] == "Tracking" and r["reset"] >= 1      # forced failure, IMU-pose re-init
    assert r["final_state"] == "Tracking" and r["guess_used"] >= 180
    assert np.abs(r["ref_acc_bias"]).max() > 1e-5
    assert np.abs(r["acc_bias"] - r["ref_acc_bias"]).max() <= 1e-7 and np.abs(r["gyro_bias"] - r["ref_gyro_bias"]).max() <= 1e-7
    assert r["kf"] >= 20 and r["n_solved"] >= 10                                        # 10-KF window filled and slid
    assert r["path"] > 0.5
    _ate_close(r)


def test_c3_kitti_shaped_stereo_20kf_window(lib):
    seq = sequences.make_c3(34)
    r = run_sequence(lib, seq, tol_pose=1e-6, tol_und=0.0, tol_p3=1e-6, window=20)
    assert r["final_state"] == "Tracking" and r["has_imu"] == 0 and r["guess_used"] == 0
    assert r["kf"] >= 28 and r["n_solved"] >= 8                                         # W = 20 window solved and slid
    assert r["path"] > 8.0
    _ate_close(r)
    assert r["ate_gt_ref"] < 0.05
