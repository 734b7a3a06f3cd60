"""Synthetic sliding-window BA problems (g2o ba_demo pattern): plain numpy arrays, no product or
oracle imports.

Geometry: P keyframes moving along +x looking down +z at a cloud of landmarks 4..12 m away; each landmark
is seen by a run of consecutive keyframes; 1 px Gaussian pixel noise, a fraction of gross outliers,
perturbed initial poses / points -- pattern of 3rdPartLib/g2o/g2o/examples/ba/ba_demo.cpp:126-250.
"""
import numpy as np

EUROC_K = (458.654, 457.296, 367.215, 248.375)


class Problem:
    def __init__(self, poses, lms, ep, el, uv, K, fixed_pose=0, fix_landmarks=0, gt=None):
        self.poses, self.lms, self.ep, self.el, self.uv, self.K = poses, lms, ep, el, uv, K
        self.fixed_pose, self.fix_landmarks, self.gt = fixed_pose, fix_landmarks, gt


def make_problem(window=10, n_landmarks=1500, obs_per_frame=480, seed=0, K=EUROC_K, w=752, h=480, noise_px=1.0,
                 outlier_frac=0.05, pose_noise=(0.01, 0.03), point_noise=0.05):
    rng = np.random.default_rng(seed)
    P = window
    gt_poses = np.zeros((P, 7)); gt_poses[:, 3] = 1.0
    gt_poses[:, 4] = -0.12 * np.arange(P)                   # T_c_w translation: camera moves along +x
    gt_poses[:, 5] = 0.01 * np.sin(np.arange(P))
    fx, fy, cx, cy = K
    ep, el, uv, lms = [], [], [], []
    # landmarks are created until every frame has ~obs_per_frame observations (each seen by 2..P consecutive KFs)
    counts = np.zeros(P, int)
    tries = 0
    while len(lms) < n_landmarks and tries < 50 * n_landmarks:
        tries += 1
        first = int(rng.integers(0, P - 1))
        run = int(rng.integers(2, P + 1))
        frames = [f for f in range(first, min(P, first + run)) if counts[f] < obs_per_frame]
        if len(frames) < 2:
            continue
        z = rng.uniform(4.0, 12.0)
        xc = (rng.uniform(40, w - 40) - cx) / fx * z; yc = (rng.uniform(40, h - 40) - cy) / fy * z
        Xw = np.array([xc - gt_poses[frames[0], 4], yc - gt_poses[frames[0], 5], z])
        obs = []
        for f in frames:
            Xc = Xw + gt_poses[f, 4:7]
            u = fx * Xc[0] / Xc[2] + cx; v = fy * Xc[1] / Xc[2] + cy
            if 0 < u < w - 1 and 0 < v < h - 1:
                obs.append((f, u, v))
        if len(obs) < 2:
            continue
        li = len(lms)
        lms.append(Xw)
        for f, u, v in obs:
            n = rng.normal(0, noise_px, 2) if noise_px > 0 else np.zeros(2)
            if rng.uniform() < outlier_frac:
                n = n + rng.uniform(-40, 40, 2)
            ep.append(f); el.append(li); uv.append((u + n[0], v + n[1])); counts[f] += 1
    lms = np.array(lms)
    poses = gt_poses.copy()
    for p in range(1, P):                                    # pose 0 is the fixed gauge
        aa = rng.normal(0, pose_noise[0], 3)
        q = np.concatenate([0.5 * aa, [1.0]]); q /= np.linalg.norm(q)
        poses[p, :4] = q
        poses[p, 4:7] += rng.normal(0, pose_noise[1], 3)
    lms_init = lms + rng.normal(0, point_noise, lms.shape)
    order = np.lexsort((np.array(el), np.array(ep)))         # keyframe-major = g2o insertion order (vo_localmap.cpp:185-208)
    return Problem(poses, lms_init, np.array(ep, np.int32)[order], np.array(el, np.int32)[order],
                   np.array(uv, np.float64)[order], K, gt=(gt_poses, lms))


def make_pose_only(n_pts=300, seed=0, K=EUROC_K, w=752, h=480, noise_px=0.7, outlier_frac=0.1):
    """OptimizeInFrame-shaped problem: one free pose, all points fixed (optimize_in_frame.cpp:35-63)."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    z = rng.uniform(2, 15, n_pts)
    X = np.stack([(rng.uniform(20, w - 20, n_pts) - cx) / fx * z, (rng.uniform(20, h - 20, n_pts) - cy) / fy * z, z], 1)
    gt = np.array([[0, 0, 0, 1.0, 0, 0, 0]])
    uv = np.stack([fx * X[:, 0] / X[:, 2] + cx, fy * X[:, 1] / X[:, 2] + cy], 1) + rng.normal(0, noise_px, (n_pts, 2))
    bad = rng.uniform(size=n_pts) < outlier_frac
    uv[bad] += rng.uniform(-30, 30, (int(bad.sum()), 2))
    aa = rng.normal(0, 0.01, 3); q = np.concatenate([0.5 * aa, [1.0]]); q /= np.linalg.norm(q)
    pose = np.concatenate([q, rng.normal(0, 0.05, 3)])[None]
    return Problem(pose, X, np.zeros(n_pts, np.int32), np.arange(n_pts, dtype=np.int32), uv, K, fixed_pose=-1,
                   fix_landmarks=1, gt=(gt, X))




def make_ba_demo(n_poses=15, n_points=500, seed=0, pixel_noise=1.0, outlier_ratio=0.0):
    """The vendored g2o demo problem (3rdPartLib/g2o/g2o/examples/ba/ba_demo.cpp:126-293): f = 1000, 640x480, principal
    point (320, 240), 15 cameras 4 cm apart along x looking down +z, 500 points in the box [-1.5,1.5] x [-0.5,0.5] x [3,4],
    a point is used when >= 2 cameras see it, Gaussian pixel noise, uniform outliers, points perturbed by N(0,1) per axis.
    The demo fixes its first TWO poses; FLVIS's call path fixes ONE (vo_localmap.cpp:149-166), which is what the C ABI
    exposes, so pose 0 is the gauge here."""
    rng = np.random.default_rng(seed)
    K = (1000.0, 1000.0, 320.0, 240.0)
    pts = np.stack([(rng.uniform(size=n_points) - 0.5) * 3, rng.uniform(size=n_points) - 0.5, rng.uniform(size=n_points) + 3], 1)
    gt_poses = np.zeros((n_poses, 7)); gt_poses[:, 3] = 1.0
    gt_poses[:, 4] = -(np.arange(n_poses) * 0.04 - 1.0)                     # T_c_w = inverse of the camera position
    ep, el, uv, keep = [], [], [], []
    for i, X in enumerate(pts):
        obs = []
        for p in range(n_poses):
            Xc = X + gt_poses[p, 4:7]
            z = np.array([K[0] * Xc[0] / Xc[2] + K[2], K[1] * Xc[1] / Xc[2] + K[3]])
            if 0 <= z[0] < 640 and 0 <= z[1] < 480:
                obs.append((p, z))
        if len(obs) < 2:
            continue
        li = len(keep); keep.append(i)
        for p, z in obs:
            if rng.uniform() < outlier_ratio:
                z = np.array([rng.uniform(0, 640), rng.uniform(0, 480)])
            z = z + rng.normal(0, pixel_noise, 2)
            ep.append(p); el.append(li); uv.append(z)
    lms_gt = pts[keep]
    lms = lms_gt + rng.normal(0, 1.0, lms_gt.shape) * 0.2
    order = np.lexsort((np.array(el), np.array(ep)))
    return Problem(gt_poses.copy(), lms, np.array(ep, np.int32)[order], np.array(el, np.int32)[order],
                   np.array(uv, np.float64)[order], K, gt=(gt_poses, lms_gt))
