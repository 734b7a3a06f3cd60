"""Restatement of CameraFrame's per-landmark geometry + Triangulation + DepthCamera -- TEST INFRASTRUCTURE.

Follows /root/reference/src/processing/camera_frame.cpp:43-91 (calReprjInlierOutlier), :93-180 (FromStereo, after
the LK call which is tested separately), :182-234 (FromDepthImg), :236-270 (FromTriangulation), :271-330
(depthInnovation); /root/reference/src/processing/triangulation.cpp:9-54,80-97; depth_camera.cpp:92-150.
The 4x4 JacobiSVD is numpy's LAPACK SVD here (the null vector is unique up to sign; V(:,3)/V(3,3) removes it).
`rand()` is glibc's TYPE_3 generator with the default seed 1 (the reference never calls srand).
No golden vectors exist in the reference for this path.
"""
import numpy as np

from .vimotion_ref import SE3, qrot, q2R


class GlibcRand:
    """glibc random_r TYPE_3 (r[i] = r[i-3] + r[i-31], output >> 1), what rand() runs with the default seed."""
    RAND_MAX = 2147483647

    def __init__(self, seed=1):
        r = [0] * 34
        r[0] = seed
        for i in range(1, 31):
            hi, lo = divmod(r[i - 1], 127773)
            w = 16807 * lo - 2836 * hi
            r[i] = w + 2147483647 if w < 0 else w
        for i in range(31, 34):
            r[i] = r[i - 31]
        self.r = [x & 0xffffffff for x in r]
        for _ in range(310):
            self._step()

    def _step(self):
        v = (self.r[-31] + self.r[-3]) & 0xffffffff
        self.r.append(v)
        self.r.pop(0)
        return v

    def rand(self):
        return self._step() >> 1

    def dummy_depth(self):
        """d_rand = 0.3 + static_cast<float>(rand())/(static_cast<float>(RAND_MAX/(0.4)))  (camera_frame.cpp:153)"""
        return float(np.float32(np.float64(0.3) + np.float64(np.float32(np.float32(self.rand()) / np.float32(self.RAND_MAX / 0.4)))))


def triangulation_pt(pt1, pt2, P1, P2):
    u1, v1 = pt1; u2, v2 = pt2
    A = np.stack([v1 * P1[2] - P1[1], P1[0] - u1 * P1[2], v2 * P2[2] - P2[1], P2[0] - u2 * P2[2]])
    _, _, Vt = np.linalg.svd(A)
    V3 = Vt[3]
    return V3[:3] / V3[3]


def proj_matrix(T, K):
    fx, fy, cx, cy = K
    Km = np.array([[fx, 0, cx], [0, fy, cy], [0, 0, 1.0]])
    return Km @ np.hstack([q2R(T.q), T.t.reshape(3, 1)])


def world2camera(p_w, T_c_w):
    return qrot(T_c_w.q, p_w) + T_c_w.t


def camera2world(p_c, T_c_w):
    Ti = T_c_w.inverse()
    return qrot(Ti.q, p_c) + Ti.t


def pixel2camera(p, K, depth):
    fx, fy, cx, cy = K
    return np.array([(p[0] - cx) * depth / fx, (p[1] - cy) * depth / fy, depth])


def camera2pixel(p_c, K):
    fx, fy, cx, cy = K
    return np.array([fx * p_c[0] / p_c[2] + cx, fy * p_c[1] / p_c[2] + cy])


class Frame:
    """SoA view of a CameraFrame's landmark list."""

    def __init__(self, T_c_w, K, plane, undist, p3d_w, has_3d, first_2d, first_pose, P0=None, P1=None, cam_type="stereo"):
        self.T_c_w, self.K = T_c_w, K
        self.plane = np.array(plane, float); self.undist = np.array(undist, float)
        self.p3d_w = np.array(p3d_w, float); self.p3d_c = np.zeros_like(self.p3d_w)
        self.has_3d = np.array(has_3d, bool)
        self.first_2d = np.array(first_2d, float); self.first_pose = first_pose     # list of SE3
        self.P0, self.P1, self.cam_type = P0, P1, cam_type
        self.inlier = np.ones(len(self.plane), bool)

    def n(self):
        return len(self.plane)


def recover_from_triangulation(fr, rng_range):
    pts, mask = [], []
    for i in range(fr.n()):
        T1 = fr.first_pose[i]
        if np.linalg.norm(T1.t - fr.T_c_w.t) >= 0.2:
            pw = triangulation_pt(fr.first_2d[i], fr.undist[i], proj_matrix(T1, fr.K), proj_matrix(fr.T_c_w, fr.K))
            pc = world2camera(pw, fr.T_c_w)
            if 0.5 <= pc[2] <= rng_range:
                pts.append(pc); mask.append(True); continue
        pts.append(np.zeros(3)); mask.append(False)
    return np.array(pts).reshape(-1, 3), np.array(mask, bool)


def recover_from_stereo(fr, pt1_undist, status, rng_range, rnd):
    pts, mask = [], []
    for i in range(fr.n()):
        u0 = np.float32(fr.undist[i]).astype(np.float64)          # cv::Point2f copies (getAll2dPlaneUndistort3d_cvPf)
        if status[i] == 1:
            pc = triangulation_pt(u0, np.float32(pt1_undist[i]).astype(np.float64), fr.P0, fr.P1)
            if not (pc[2] < 0 or pc[2] > rng_range):
                pts.append(pc); mask.append(True); continue
        pts.append(pixel2camera(u0, fr.K, rnd.dummy_depth())); mask.append(False)
    return np.array(pts).reshape(-1, 3), np.array(mask, bool)


def recover_from_depth(fr, depth_img, scale, rng_range, rnd):
    pts, mask = [], []
    f = np.float32
    for i in range(fr.n()):
        px = float(np.float32(np.floor(abs(fr.plane[i, 0]) + 0.5) * np.sign(fr.plane[i, 0])))     # C round(): half away from zero
        py = float(np.float32(np.floor(abs(fr.plane[i, 1]) + 0.5) * np.sign(fr.plane[i, 1])))
        z = float(f(f(depth_img[int(py), int(px)]) / f(scale)))
        if z >= 0.3 and z <= rng_range:
            fx, fy, cx, cy = fr.K
            pts.append(np.array([(px - cx) * z / fx, (py - cy) * z / fy, z])); mask.append(True)
        else:
            pts.append(pixel2camera(fr.plane[i], fr.K, rnd.dummy_depth())); mask.append(False)
    return np.array(pts).reshape(-1, 3), np.array(mask, bool)


def depth_innovation(fr, iir_ratio, rng_range, dummy_depth, rnd, stereo=None, depth=None):
    """stereo = (pt1_undist, status) or depth = (u16 image, scale).  Mutates fr.p3d_c / p3d_w / has_3d."""
    iir = float(np.float32(iir_ratio)); rr = float(np.float32(rng_range))
    tri, tri_mask = recover_from_triangulation(fr, rr)
    if depth is not None:
        cam, cam_mask = recover_from_depth(fr, depth[0], depth[1], rr, rnd)
    else:
        cam, cam_mask = recover_from_stereo(fr, stereo[0], stereo[1], rr, rnd)
    for i in range(fr.n()):
        if not cam_mask[i] and not tri_mask[i]:
            if depth is None:
                if not fr.has_3d[i] and dummy_depth:
                    fr.p3d_c[i] = cam[i]; fr.p3d_w[i] = camera2world(cam[i], fr.T_c_w); fr.has_3d[i] = True
                continue
        meas = cam[i] if cam_mask[i] else tri[i]
        if fr.has_3d[i]:
            lm_c = world2camera(fr.p3d_w[i], fr.T_c_w)
            upd = lm_c * iir + meas * (1 - iir)
            fr.p3d_c[i] = upd; fr.p3d_w[i] = camera2world(upd, fr.T_c_w)
        else:
            fr.p3d_c[i] = meas; fr.p3d_w[i] = camera2world(meas, fr.T_c_w); fr.has_3d[i] = True


def cal_reprj_inlier_outlier(fr, sh_over_med):
    """Returns (mean_prjerr, outlier plane points in the reference's reverse order); sets fr.inlier."""
    d = np.array([np.linalg.norm(fr.undist[i] - camera2pixel(world2camera(fr.p3d_w[i], fr.T_c_w), fr.K)) for i in range(fr.n())])
    valid = np.sort(d[d < 3.0])
    mean = valid.sum() / len(valid) if len(valid) else float("nan")
    sh = sh_over_med * valid[len(valid) // 2]
    if sh >= 3.0: sh = 3.0
    out = []
    for i in range(fr.n() - 1, -1, -1):
        if d[i] > sh:
            out.append(fr.plane[i]); fr.inlier[i] = False
        else:
            fr.inlier[i] = True
    return mean, np.array(out).reshape(-1, 2)
