"""Generates tests/golden/*.npz from the reference's own OpenCV library (cv2 4.13.0 in this container).

Run here (needs cv2); the fixtures are committed so the tests never need cv2 or /root/reference:
    python tests/golden/make_golden.py
Inputs are regenerated from seeds by oracle/synth.py (pure numpy), so only seeds/params and cv2's outputs
are stored.  Calls mirror the reference call sites:
  LK    src/processing/lkorb_tracking.cpp:64-73 (maxLevel 10), src/processing/camera_frame.cpp:124-128 (maxLevel 5)
  GFTT  src/processing/feature_dem.cpp:160, :221
  pyrDown: inside cv::buildOpticalFlowPyramid
"""
import os
import sys
import numpy as np
import cv2

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from synthdata import textures as synth  # noqa: E402

LK_CASES = [
    # name, seed, (h,w), shift, rot, scale, npts, max_level
    ("euroc_shift", 11, (480, 752), (5.3, -3.1), 0.0, 1.0, 480, 10),
    ("euroc_rot", 12, (480, 752), (-7.7, 4.4), 1.5, 1.01, 480, 10),
    ("kitti_stereo", 13, (376, 1241), (-18.5, 0.0), 0.0, 1.0, 480, 5),
    ("d435_big", 14, (480, 640), (21.0, 13.0), -2.0, 0.98, 300, 10),
]
GFTT_CASES = [
    # name, seed, (h,w), blur, N, q, d
    ("euroc", 21, (480, 752), 2, 1000, 0.01, 10),
    ("euroc_2n", 22, (480, 752), 1, 2000, 0.01, 10),
    ("kitti", 23, (376, 1241), 2, 2000, 0.0001, 10),
    ("kitti_2n", 24, (376, 1241), 1, 4000, 0.0001, 10),
    ("d435", 25, (480, 640), 2, 500, 0.01, 15),
    ("d435_2n", 26, (480, 640), 3, 1000, 0.01, 15),
]


def lk_case(seed, hw, shift, rot, scale, npts):
    h, w = hw
    I, J, flow = synth.frame_pair(seed, h, w, shift, rot, scale)
    corners = cv2.goodFeaturesToTrack(I, npts - 24, 0.01, 10).reshape(-1, 2).astype(np.float32)
    extra = synth.grid_points(h, w, 16, seed + 1000, border=0)
    edge = np.array([[0.5, 0.5], [w - 1.0, h - 1.0], [2.0, h - 10.0], [w - 3.25, 7.5], [w / 2.0, 1.25],
                     [5.3, 7.7], [w - 3.8, h - 3.1], [w / 2.0, h - 1.5]], np.float32)
    pts = np.vstack([corners, extra, edge]).astype(np.float32)
    init = pts.copy()
    # half the points start from a (noisy) motion prediction, like the IMU-guess path of lkorb_tracking.cpp:38-63
    rng = np.random.default_rng(seed + 7)
    pred = flow(pts).astype(np.float32) + rng.normal(0, 1.5, pts.shape).astype(np.float32)
    init[::2] = pred[::2]
    return I, J, pts, init


def main():
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001)
    for name, seed, hw, shift, rot, scale, npts, max_level in LK_CASES:
        I, J, pts, init = lk_case(seed, hw, shift, rot, scale, npts)
        nxt, st, err = cv2.calcOpticalFlowPyrLK(I, J, pts, init.copy(), winSize=(31, 31), maxLevel=max_level,
                                                criteria=crit, flags=cv2.OPTFLOW_USE_INITIAL_FLOW)
        pyr = [I]
        for _ in range(3):
            pyr.append(cv2.pyrDown(pyr[-1]))
        np.savez_compressed(os.path.join(HERE, f"lk_{name}.npz"), seed=seed, hw=hw, shift=shift, rot=rot, scale=scale,
                            npts=npts, max_level=max_level, pts=pts, init=init, next=nxt, status=st.ravel(),
                            err=err.ravel(), pyr_crc=np.array([int(p.astype(np.uint64).sum()) for p in pyr]),
                            pyr3=pyr[3], cv2_version=cv2.__version__)
        print(name, "tracked", int(st.sum()), "/", len(pts))
    for name, seed, hw, blur, N, q, d in GFTT_CASES:
        img = synth.texture(seed, hw[0], hw[1], blur)
        c = cv2.goodFeaturesToTrack(img, N, q, d).reshape(-1, 2)
        eig = cv2.cornerMinEigenVal(img, 3, ksize=3)
        # store the response map sparsely: row/col checksums + a strided sample + max
        np.savez_compressed(os.path.join(HERE, f"gftt_{name}.npz"), seed=seed, hw=hw, blur=blur, N=N, q=q, d=d,
                            corners=c.astype(np.float32), eig_max=eig.max(),
                            eig_sample=eig[::7, ::5].copy(), eig_rowsum=eig.astype(np.float64).sum(axis=1),
                            cv2_version=cv2.__version__)
        print(name, "corners", len(c))


if __name__ == "__main__":
    main()
