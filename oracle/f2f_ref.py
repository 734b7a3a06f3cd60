"""Restatement of F2FTracking::image_feed / init_frame, LKORBTracking::tracking, OptimizeInFrame::optimize and the
CameraFrame glue -- TEST INFRASTRUCTURE (oracle).

Follows /root/reference/src/frontend/f2f_tracking.cpp:5-453, src/processing/lkorb_tracking.cpp:9-202,
optimize_in_frame.cpp:10-90, camera_frame.cpp (via oracle/camera_frame_ref.py), landmark.cpp:3-44.
OpenCV calls: LK = oracle/lk_ref.py through its C twin oracle/lk_ref.c (bit-identical, fast; bit-exact contract of the CUDA kernel; cv2 itself is <= 1e-3 px away),
FeatureDEM = oracle/feature_dem_ref.py, findFundamentalMat / solvePnPRansac = cv2 (the reference's real library).
Sensor types: "depth" (DEPTH_D435), "stereo" (STEREO_RECT, zero distortion) and "stereo_unrect" (STEREO_UNRECT: LK on the
raw images, cv2.undistortPoints / cv2.projectPoints per point with the raw lens models `lens0` / `lens1` =
(K 3x3, D, R 3x3)); `equalize` = need_equal_hist (cv2.equalizeHist on ingest).  correction_feed + the local-map feedback
step (f2f_tracking.cpp:40-44, :189-219; never called by the reference's own nodelet) are restated with their quirks.
"""
import math

import cv2
import numpy as np

from . import ba_ref, camera_frame_ref as cf, feature_dem_ref, lk_ref
from .vimotion_ref import SE3, VIMOTION, q2R, R2q, qn

f32 = np.float32


class LM:
    __slots__ = ("lm_id", "p3d_w", "p3d_c", "plane", "undist", "has_3d", "inlier", "first_2d", "first_pose")

    def copy(self):
        o = LM()
        for k in self.__slots__:
            v = getattr(self, k)
            setattr(o, k, v.copy() if isinstance(v, np.ndarray) else v)
        return o


class Frame:
    def __init__(self):
        self.frame_id = 0; self.time = 0.0; self.img0 = None; self.img1 = None; self.depth = None
        self.lms = []; self.T_c_w = SE3(); self.reproj_err = 0.0

    def clear(self):
        self.T_c_w = SE3(); self.lms = []; self.img0 = self.img1 = self.depth = None


def so3_log(q):
    n = math.sqrt(q[1] ** 2 + q[2] ** 2 + q[3] ** 2); w = q[0]
    f = (2. / w - 2. * (n * n) / (w * w * w)) if n < 1e-10 else 2 * math.atan(n / w) / n
    return np.array([f * q[1], f * q[2], f * q[3]])


class F2FTracking:
    def __init__(self, cam_type, w, h, K, feature_para, vi_para, dc_para, T_i_c=None, skip=0, depth_scale=1000.0,
                 K1=None, P0=None, P1=None, T_c1_c0=None, lens0=None, lens1=None, equalize=False):
        self.cam_type, self.w, self.h, self.K = cam_type, w, h, tuple(K)
        self.fd = feature_dem_ref.FeatureDEM(w, h, feature_para)
        self.vim = VIMOTION(T_i_c or SE3(), 9.81, vi_para[0], vi_para[1], vi_para[2], vi_para[3])
        self.iir = float(f32(dc_para[0])); self.range = float(f32(dc_para[1])); self.dummy = not (dc_para[2] < 0.5)
        self.skip = skip; self.depth_scale = depth_scale
        self.K1, self.P0, self.P1, self.T_c1_c0 = K1, P0, P1, T_c1_c0
        self.lens0, self.lens1, self.equalize = lens0, lens1, equalize
        self.curr, self.last = Frame(), Frame()
        self.state = "UnInit"; self.has_imu = False; self.frameCount = 0
        self.id_index = 100; self.rnd = cf.GlibcRand()
        self.T_kf = SE3(); self.fail_cnt = 0; self.tf_cnt = 0
        self.counts = (0, 0, 0)
        self.pose_records = []                     # [frame_id, T_c_w]
        self.feedback = None

    def correction_feed(self, corr):
        """corr: dict(frame_id, T_c_w (7,), lm_id, lm_3d, outlier_id) = the local map's CorrectionInf (f2f_tracking.cpp:40-44)."""
        self.feedback = corr

    def _apply_feedback(self):                     # f2f_tracking.cpp:189-219
        c = self.feedback
        self.feedback = None
        if not self.pose_records:
            return
        def as_int(v):                             # `int correct_lm_id = ids.at(i)` (camera_frame.cpp:349,366)
            return int(np.int64(v).astype(np.int32))
        corr_id = as_int(c["frame_id"])
        idx = 0
        for i in range(len(self.pose_records) - 1, -1, -1):
            if self.pose_records[i][0] == corr_id:
                idx = i
                break
        old_inv = self.pose_records[idx][1].inverse()
        upd = SE3.from7(np.asarray(c["T_c_w"], float))
        for i in range(idx, len(self.pose_records)):
            self.pose_records[i][1] = (self.pose_records[i][1] * old_inv) * upd
        self.last.T_c_w = (self.last.T_c_w * old_inv) * upd
        # correctLMP3DWByLMP3DCandT iterates by value in the reference: no effect
        for lid, p in zip(c["lm_id"], c["lm_3d"]):
            for l in self.last.lms:
                if l.lm_id == as_int(lid):
                    l.p3d_w = np.array(p, float)
                    break
        for lid in c["outlier_id"]:
            for l in self.last.lms:
                if l.lm_id == as_int(lid):
                    l.inlier = False

    def imu_feed(self, t, acc, gyro):
        if not self.vim.imu_initialized:
            self.has_imu = True
        return self.vim.imu_feed(t, acc, gyro)

    def _new_lm(self, pt, und, T, inlier=True):
        lm = LM()
        lm.lm_id = self.id_index; self.id_index += 1
        lm.first_2d = np.array(und, float); lm.undist = np.array(und, float); lm.plane = np.array(pt, float)
        lm.first_pose = T; lm.inlier = inlier; lm.p3d_w = np.zeros(3); lm.p3d_c = np.zeros(3); lm.has_3d = False
        return lm

    def _undist(self, lens, P, pts):
        """cv::undistortPoints(pts, K, D, R, P) on an (n,2) float32 array."""
        if len(pts) == 0:
            return np.zeros((0, 2), f32)
        K, D, R = lens
        return cv2.undistortPoints(np.asarray(pts, f32).reshape(-1, 1, 2), K, D, R=R, P=P[:3, :3]).reshape(-1, 2)

    def _project(self, lens, T, p3d):
        """cv::projectPoints(p3d, rvec(T), tvec(T), K, D) -> (n,2) float32."""
        K, D, _ = lens
        rvec, _ = cv2.Rodrigues(q2R(T.q))
        out, _ = cv2.projectPoints(np.asarray(p3d, np.float64).reshape(-1, 1, 3), rvec, T.t.reshape(3, 1), K, D)
        return out.reshape(-1, 2).astype(f32)

    # -- CameraFrame::depthInnovation through the SoA oracle
    def _depth_innovation(self, fr):
        n = len(fr.lms)
        if n == 0:
            return
        F = cf.Frame(fr.T_c_w, self.K, [l.plane for l in fr.lms], [l.undist for l in fr.lms], [l.p3d_w for l in fr.lms],
                     [l.has_3d for l in fr.lms], [l.first_2d for l in fr.lms], [l.first_pose for l in fr.lms], self.P0, self.P1)
        F.p3d_c = np.array([l.p3d_c for l in fr.lms], float).reshape(-1, 3)
        if self.cam_type == "depth":
            cf.depth_innovation(F, self.iir, self.range, self.dummy, self.rnd, depth=(fr.depth, self.depth_scale))
        else:
            prev = np.array([l.plane for l in fr.lms], f32)
            init = prev.copy()
            T1 = self.T_c1_c0 * fr.T_c_w
            if self.cam_type == "stereo_unrect":
                idx = [i for i, l in enumerate(fr.lms) if l.has_3d]
                if idx:
                    init[idx] = self._project(self.lens1, T1, np.array([f32(fr.lms[i].p3d_w) for i in idx]))
            else:
                for i, l in enumerate(fr.lms):
                    if l.has_3d:
                        pc = cf.world2camera(f32(l.p3d_w).astype(np.float64), T1)
                        init[i] = (f32(self.K1[0] * pc[0] / pc[2] + self.K1[2]), f32(self.K1[1] * pc[1] / pc[2] + self.K1[3]))
            nxt, st, _ = lk_ref.calc_optical_flow_pyr_lk_c(fr.img0, fr.img1, prev, init, max_level=5)
            if self.cam_type == "stereo_unrect":
                nxt = self._undist(self.lens1, self.P1, nxt)
            cf.depth_innovation(F, self.iir, self.range, self.dummy, self.rnd, stereo=(nxt.astype(np.float64), st))
        for i, l in enumerate(fr.lms):
            if F.has_3d[i]:
                l.p3d_w = F.p3d_w[i].copy(); l.p3d_c = F.p3d_c[i].copy(); l.has_3d = True

    def _init_frame(self):
        pts = self.fd.detect(self.curr.img0)
        und = self._undist(self.lens0, self.P0, pts) if self.cam_type == "stereo_unrect" else pts
        for p, u in zip(pts, und):
            self.curr.lms.append(self._new_lm(p, u, self.curr.T_c_w, True))
        self._depth_innovation(self.curr)
        self.curr.lms = [l for l in self.curr.lms if l.has_3d]
        if sum(1 for l in self.curr.lms if l.has_3d and l.inlier) > 30:
            self.pose_records.append([self.curr.frame_id, self.curr.T_c_w])
            self.T_kf = self.curr.T_c_w
            return True
        return False

    def _tracking(self, frm, to, guess, use_guess):
        n = len(frm.lms)
        fplane = np.array([l.plane for l in frm.lms], f32).reshape(-1, 2)
        fund = np.array([l.undist for l in frm.lms], f32).reshape(-1, 2)
        fp3 = np.array([l.p3d_w for l in frm.lms], f32).reshape(-1, 3)
        tplane = fplane.copy()
        if use_guess and self.cam_type == "stereo_unrect":
            if n:
                tplane = self._project(self.lens0, guess, fp3)
        elif use_guess:
            for i in range(n):
                pc = cf.world2camera(fp3[i].astype(np.float64), guess)
                px = cf.camera2pixel(pc, self.K)
                tplane[i] = (f32(px[0]), f32(px[1]))
        nxt, st, _ = lk_ref.calc_optical_flow_pyr_lk_c(frm.img0, to.img0, fplane, tplane, max_level=10)
        tplane = nxt
        tund = self._undist(self.lens0, self.P0, tplane) if self.cam_type == "stereo_unrect" else tplane.copy()
        to.lms = []
        keep = np.ones(n, bool)
        wl, hl = self.w - 1, self.h - 1
        of_cnt = 0
        for i in range(n - 1, -1, -1):
            if st[i] == 1 and tplane[i, 0] > 0 and tplane[i, 1] > 0 and tplane[i, 0] < wl and tplane[i, 1] < hl:
                of_cnt += 1
                lm = frm.lms[i].copy()
                lm.plane = tplane[i].astype(np.float64); lm.undist = tund[i].astype(np.float64)
                to.lms.append(lm)
            else:
                keep[i] = False
        fund_k, tund_k = fund[keep], tund[keep]
        self.counts = (of_cnt, 0, 0)
        if of_cnt < 10:
            return False
        _, maskF = cv2.findFundamentalMat(fund_k, tund_k, cv2.FM_RANSAC, 5.0, 0.99)
        maskF = np.zeros(len(fund_k), np.uint8) if maskF is None else maskF.ravel()
        for i in range(len(maskF)):
            if maskF[i] == 0:
                to.lms[i].inlier = False                       # mirrored index, as the reference (lkorb_tracking.cpp:138-149)
        f_cnt = sum(1 for l in to.lms if l.inlier)
        self.counts = (of_cnt, f_cnt, 0)
        if f_cnt < 10:
            return False
        sel = [l for l in to.lms if l.has_3d and l.inlier]
        p2d = np.array([l.undist for l in sel], f32).reshape(-1, 2)
        p3d = np.array([l.p3d_w for l in sel], f32).reshape(-1, 3)
        T, inl = pnp_cv2(p3d, p2d, self.K, guess if use_guess else None)
        mask = np.zeros(len(sel), np.uint8); mask[inl] = 1
        k = 0
        for l in to.lms:
            if l.has_3d and l.inlier:
                if mask[k] == 0:
                    l.inlier = False
                k += 1
        to.T_c_w = T
        self.counts = (of_cnt, f_cnt, len(inl))
        return len(inl) >= 10

    def _optimize_in_frame(self, fr):
        sel = [l for l in fr.lms if l.has_3d and l.inlier]
        if len(sel) < 10:
            return False
        n = len(sel)
        from .localmap_ref import _g2o_pose
        d = ba_ref.BAData(_g2o_pose(fr.T_c_w.to7())[None], np.array([l.p3d_w for l in sel]), np.zeros(n, np.int32),
                          np.arange(n, dtype=np.int32), np.array([l.undist for l in sel]), self.K, fixed_pose=-1, fix_landmarks=1)
        st = ba_ref.optimize(d, 2, 2, min_edges_after_cull=10)
        if not st.ok:
            return False
        fr.T_c_w = SE3.from7(d.poses[0])
        return True

    def _reprj(self, fr, sh):
        F = cf.Frame(fr.T_c_w, self.K, [l.plane for l in fr.lms], [l.undist for l in fr.lms], [l.p3d_w for l in fr.lms],
                     [l.has_3d for l in fr.lms], [l.first_2d for l in fr.lms], [l.first_pose for l in fr.lms])
        mean, _ = cf.cal_reprj_inlier_outlier(F, sh)
        for i, l in enumerate(fr.lms):
            l.inlier = bool(F.inlier[i])
        fr.reproj_err = mean

    def image_feed(self, t, img0, img1):
        """Returns (new_keyframe, reset_cmd)."""
        new_kf = reset = False
        self.frameCount += 1
        self.last, self.curr = self.curr, self.last
        self.curr.clear(); self.curr.frame_id = self.frameCount; self.curr.time = t
        self.curr.img0 = img0
        if self.cam_type == "depth": self.curr.depth = img1
        else: self.curr.img1 = img1
        if self.skip > 0:
            self.skip -= 1
            return new_kf, reset
        if self.equalize:                                   # f2f_tracking.cpp:125-145
            self.curr.img0 = cv2.equalizeHist(self.curr.img0)
            if self.cam_type != "depth":
                self.curr.img1 = cv2.equalizeHist(self.curr.img1)
        if self.state == "UnInit":
            R_w_c = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
            self.curr.T_c_w = SE3(R2q(R_w_c), np.zeros(3)).inverse()
            if self.has_imu:
                if self.vim.imu_initialized:
                    q = self.vim.vision_trigger()
                    Rwc = q2R(q) @ q2R(self.vim.T_i_c.q)
                    self.curr.T_c_w = SE3(R2q(Rwc), np.zeros(3)).inverse()
                else:
                    return new_kf, reset
            if self._init_frame():
                new_kf = True; self.state = "Tracking"
        elif self.state == "Tracking":
            if self.feedback is not None:
                self._apply_feedback()
            guess = None
            if self.has_imu:
                guess = self.vim.corr_frame_state(t)
            ok = self._tracking(self.last, self.curr, guess, guess is not None)
            if not ok:
                self.fail_cnt += 1
                self.last, self.curr = self.curr, self.last
                if self.fail_cnt >= 2: self.state = "TrackingFail"; self.fail_cnt = 0
                return new_kf, reset
            self.fail_cnt = 0
            if self.has_imu:
                self.curr.T_c_w = self.vim.rp_compensation(self.curr.time, self.curr.T_c_w)
            if not self._optimize_in_frame(self.curr):
                self.fail_cnt += 1
                self.last, self.curr = self.curr, self.last
                if self.fail_cnt >= 2: self.state = "TrackingFail"; self.fail_cnt = 0
                return new_kf, reset
            self._reprj(self.curr, 1.5)
            self.curr.lms = [l for l in self.curr.lms if l.inlier]
            if self.has_imu:
                self.vim.correction_from_vision(self.curr.time, self.curr.T_c_w, self.last.time, self.last.T_c_w)
            orig = len(self.curr.lms)
            new = self.fd.redetect(self.curr.img0, np.array([l.plane for l in self.curr.lms], np.float64).reshape(-1, 2))
            und = self._undist(self.lens0, self.P0, new) if self.cam_type == "stereo_unrect" else new
            for p, u in zip(new, und):
                self.curr.lms.append(self._new_lm(p, u, self.curr.T_c_w, orig < 60))
            self._depth_innovation(self.curr)
            self.curr.lms = [l for l in self.curr.lms if l.has_3d]
            self.pose_records.append([self.curr.frame_id, self.curr.T_c_w])
            if len(self.pose_records) >= 1000:
                self.pose_records.pop(0)
            Td = self.T_kf * self.curr.T_c_w.inverse()
            r = so3_log(Td.q)
            t_norm = abs(Td.t[0]) + abs(Td.t[1]) + abs(Td.t[2]); r_norm = abs(r[0]) + abs(r[1]) + abs(r[2])
            if self.frameCount < 40 and self.frameCount % 5 == 0:
                new_kf = True; self.T_kf = self.curr.T_c_w
            if t_norm >= 0.05 or r_norm >= 0.2:
                new_kf = True; self.T_kf = self.curr.T_c_w
        else:
            self.tf_cnt += 1
            if self.tf_cnt % 3 == 0:
                T = self.vim.corr_frame_state(self.curr.time)
                if T is not None:
                    self.curr.T_c_w = T
                    if self._init_frame():
                        new_kf = True; self.state = "Tracking"
                    else:
                        self.last, self.curr = self.curr, self.last
                else:
                    self.last, self.curr = self.curr, self.last
                self.tf_cnt = 0
            else:
                self.last, self.curr = self.curr, self.last
                if self.tf_cnt % 2 == 0: reset = True
        return new_kf, reset


def pnp_cv2(p3d, p2d, K, guess):
    """cv::solvePnPRansac exactly as lkorb_tracking.cpp:170-177 calls it. Returns (T_c_w SE3, inlier indices)."""
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    D = np.zeros(4)
    if guess is not None:
        rvec, _ = cv2.Rodrigues(q2R(guess.q)); tvec = guess.t.reshape(3, 1).copy()
        ok, rvec, tvec, inl = cv2.solvePnPRansac(p3d, p2d, Km, D, rvec, tvec, False, 100, 3.0, 0.99, flags=cv2.SOLVEPNP_ITERATIVE)
    else:
        ok, rvec, tvec, inl = cv2.solvePnPRansac(p3d, p2d, Km, D, None, None, False, 100, 3.0, 0.99, flags=cv2.SOLVEPNP_P3P)
    inl = np.zeros(0, np.int32) if inl is None else inl.ravel().astype(np.int32)
    R, _ = cv2.Rodrigues(rvec)
    return SE3(R2q(R), np.asarray(tvec, float).ravel()), inl


def fmat_cv2(from_xy, to_xy):
    _, m = cv2.findFundamentalMat(from_xy, to_xy, cv2.FM_RANSAC, 5.0, 0.99)
    return np.zeros(len(from_xy), np.uint8) if m is None else m.ravel().astype(np.uint8)


def make_stereo_sequence(n_frames, seed=0, w=640, h=480, K=(384.16455, 384.16455, 320.21445, 238.94403), Z=3.0, baseline=0.05):
    """Rectified stereo views of the same fronto-parallel plane: right image = left image shifted by the constant
    disparity fx*b/Z.  Returns (left images, right images, P0, P1, T_cam1_cam0 as SE3)."""
    from synthdata import textures as synth
    margin = 96
    canvas = synth.texture(seed, h + 2 * margin, w + 2 * margin, blur=2)
    lefts, rights = [], []
    cx, cy = w / 2.0, h / 2.0
    disp = K[0] * baseline / Z
    for k in range(n_frames):
        tx = 0.012 * k; ty = 0.006 * math.sin(0.5 * k); th = math.radians(0.15 * k)
        sx, sy = K[0] * tx / Z, K[1] * ty / Z
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        Rinv = np.linalg.inv(R)
        for dx, out in ((0.0, lefts), (-disp, rights)):
            tvec = np.array([cx, cy]) + np.array([sx + dx, sy])
            A = np.zeros((2, 3)); A[:, :2] = Rinv; A[:, 2] = -Rinv @ tvec + np.array([cx, cy]) + margin
            out.append(synth.warp_affine(canvas, A, h, w))
    P0 = np.array([[K[0], 0, K[2], 0], [0, K[1], K[3], 0], [0, 0, 1, 0.0]])
    P1 = np.array([[K[0], 0, K[2], -K[0] * baseline], [0, K[1], K[3], 0], [0, 0, 1, 0.0]])
    return lefts, rights, P0, P1, SE3([1.0, 0, 0, 0], [-baseline, 0, 0])


def make_depth_sequence(n_frames, seed=0, w=640, h=480, K=(384.16455, 384.16455, 320.21445, 238.94403), Z=3.0):
    """Fronto-parallel textured plane at depth Z seen by a camera that translates parallel to it and rolls slightly:
    the image motion is an exact similarity, the depth image is constant (mm)."""
    from synthdata import textures as synth
    margin = 96
    canvas = synth.texture(seed, h + 2 * margin, w + 2 * margin, blur=2)
    imgs, depths = [], []
    cx, cy = w / 2.0, h / 2.0
    for k in range(n_frames):
        tx = 0.012 * k; ty = 0.006 * math.sin(0.5 * k); th = math.radians(0.15 * k)
        sx, sy = K[0] * tx / Z, K[1] * ty / Z
        R = np.array([[math.cos(th), -math.sin(th)], [math.sin(th), math.cos(th)]])
        Rinv = np.linalg.inv(R)
        tvec = np.array([cx, cy]) + np.array([sx, sy])
        A = np.zeros((2, 3)); A[:, :2] = Rinv; A[:, 2] = -Rinv @ tvec + np.array([cx, cy]) + margin
        imgs.append(synth.warp_affine(canvas, A, h, w))
        depths.append(np.full((h, w), int(Z * 1000), np.uint16))
    return imgs, depths
