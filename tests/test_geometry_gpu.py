"""K6 per-landmark geometry (depthInnovation + calReprjInlierOutlier) against the Python restatement of
camera_frame.cpp / triangulation.cpp.  fp64; the 4x4 SVD differs (one-sided Jacobi vs LAPACK), so positions are
compared to 1e-7 m relative to metres-scale points; flags (has_3d, inlier) and the rand() draw count bit-exact."""
import ctypes as C

import numpy as np
import pytest

from oracle import camera_frame_ref as cf
from oracle.vimotion_ref import SE3, q2R

pytestmark = pytest.mark.gpu
K = (458.654, 457.296, 367.215, 248.375)
BASE = 0.11


def _make_frame(n, seed, cam_type):
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    aa = rng.normal(0, 0.05, 3); q = np.concatenate([[1.0], 0.5 * aa]); T = SE3(q, rng.normal(0, 0.3, 3))
    z = rng.uniform(1.0, 20.0, n)
    und = np.stack([rng.uniform(5, 747, n), rng.uniform(5, 475, n)], 1)
    Xc = np.stack([(und[:, 0] - cx) / fx * z, (und[:, 1] - cy) / fy * z, z], 1)
    Xw = np.array([cf.camera2world(x, T) for x in Xc])
    plane = und + rng.normal(0, 0.3, und.shape)
    has = rng.uniform(size=n) < 0.6
    p3d_w = Xw + rng.normal(0, 0.05, Xw.shape)
    # first observations: half of them from a pose with a >= 0.2 m baseline
    first_pose, first_2d = [], []
    for i in range(n):
        if i % 2 == 0:
            T1 = SE3(T.q, T.t + np.array([0.35, 0.02, -0.01]))
        else:
            T1 = SE3(T.q, T.t + np.array([0.05, 0.0, 0.0]))
        first_pose.append(T1)
        first_2d.append(cf.camera2pixel(cf.world2camera(Xw[i], T1), K) + rng.normal(0, 0.2, 2))
    P0 = np.array([[fx, 0, cx, 0], [0, fy, cy, 0], [0, 0, 1, 0.0]])
    P1 = np.array([[fx, 0, cx, -fx * BASE], [0, fy, cy, 0], [0, 0, 1, 0.0]])
    fr = cf.Frame(T, K, plane, und, p3d_w, has, np.array(first_2d), first_pose, P0, P1, cam_type)
    # stereo measurements: disparity from the true depth (+ noise), some LK failures, some absurd matches
    pt1 = und.copy(); pt1[:, 0] -= fx * BASE / z; pt1 += rng.normal(0, 0.1, pt1.shape)
    status = (rng.uniform(size=n) < 0.85).astype(np.uint8)
    pt1[::17, 0] += 30.0                                    # negative depth -> range gate fails -> dummy draw
    depth = np.zeros((480, 752), np.uint16)
    for i in range(n):
        depth[int(round(plane[i, 1])), int(round(plane[i, 0]))] = 0 if i % 11 == 0 else int(z[i] * 1000)
    return fr, pt1.astype(np.float32).astype(np.float64), status, depth


@pytest.mark.parametrize("cam_type", ["stereo", "depth"])
def test_depth_innovation_matches_oracle(cam_type):
    from flvis_b200 import capi
    S, M = 3, 512
    ctx = capi.Context(S, 752, 480, M)
    ns = [480, 137, 0]
    frames = [_make_frame(n, 10 + s, cam_type) for s, n in enumerate(ns)]
    cam = capi.Camera(*K, (C.c_double * 12)(*frames[0][0].P0.ravel()), (C.c_double * 12)(*frames[0][0].P1.ravel()),
                      0 if cam_type == "depth" else 1, 1000.0)
    prm = capi.DepthParams(0.9, 50.0 if cam_type == "stereo" else 40.0, 1)
    z = lambda *sh, dt=np.float64: np.zeros(sh, dt)
    n_lms = np.array(ns, np.int32); T = z(S, 7); plane = z(S, M, 2); und = z(S, M, 2); p3w = z(S, M, 3); p3c = z(S, M, 3)
    has = z(S, M, dt=np.uint8); f2 = z(S, M, 2); fp = z(S, M, 7); fp[:, :, 3] = 1; pt1 = z(S, M, 2); stt = z(S, M, dt=np.uint8)
    dat = z(S, M, dt=np.uint16); rnd = z(S, M, dt=np.float32); used = z(S, dt=np.int32)
    gens = [cf.GlibcRand() for _ in range(S)]
    for s, (fr, p1, st, depth) in enumerate(frames):
        n = ns[s]
        T[s] = fr.T_c_w.to7(); plane[s, :n] = fr.plane.reshape(-1, 2); und[s, :n] = fr.undist.reshape(-1, 2)
        p3w[s, :n] = fr.p3d_w.reshape(-1, 3); has[s, :n] = fr.has_3d
        f2[s, :n] = np.asarray(fr.first_2d).reshape(-1, 2); fp[s, :n] = np.array([t.to7() for t in fr.first_pose]).reshape(-1, 7)
        pt1[s, :n] = p1.reshape(-1, 2); stt[s, :n] = st
        for i in range(n):
            dat[s, i] = depth[int(round(fr.plane[i, 1])), int(round(fr.plane[i, 0]))]
        g2 = cf.GlibcRand()
        rnd[s] = [g2.dummy_depth() for _ in range(M)]
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = ctx.lib.flv_depth_innovation(ctx.h, S, vp(n_lms), C.byref(cam), C.byref(prm), vp(T), vp(plane), vp(und), vp(p3w),
                                      vp(p3c), vp(has), vp(f2), vp(fp), vp(pt1), vp(stt), vp(dat), vp(rnd), vp(used),
                                      capi.MEM_HOST)
    assert rc == 0, ctx.lib.flv_last_error(ctx.h)
    for s, (fr, p1, st, depth) in enumerate(frames):
        n = ns[s]
        before = gens[s]
        if cam_type == "stereo":
            cf.depth_innovation(fr, 0.9, 50.0, True, before, stereo=(p1, st))
        else:
            cf.depth_innovation(fr, 0.9, 40.0, True, before, depth=(depth, 1000.0))
        drawn = 0
        g3 = cf.GlibcRand()
        while g3.r != before.r:
            g3.rand(); drawn += 1
        assert used[s] == drawn                                              # rand() draws: exact count, in order
        assert np.array_equal(has[s, :n].astype(bool), fr.has_3d)             # flags bit-exact
        m = fr.has_3d
        if m.any():
            scale = np.maximum(1.0, np.abs(fr.p3d_c[m]).max())
            assert np.abs(p3c[s, :n][m] - fr.p3d_c[m]).max() <= 1e-7 * scale
            assert np.abs(p3w[s, :n][m] - fr.p3d_w[m]).max() <= 1e-7 * scale
    # reprojection inlier test on the updated landmarks
    inl = z(S, M, dt=np.uint8); mean = z(S)
    n2 = np.array([ns[0], ns[1], 0], np.int32)
    rc = ctx.lib.flv_reprojection_inliers(ctx.h, 2, vp(n2), C.byref(cam), vp(T), vp(und), vp(p3w), 1.5, vp(inl), vp(mean),
                                          capi.MEM_HOST)
    assert rc == 0
    for s in range(2):
        fr = frames[s][0]
        fr.p3d_w = p3w[s, :ns[s]].copy()          # same inputs as the device
        mo, _ = cf.cal_reprj_inlier_outlier(fr, 1.5)
        assert np.array_equal(inl[s, :ns[s]].astype(bool), fr.inlier)
        assert abs(mean[s] - mo) <= 1e-9 * max(1.0, mo)
    ctx.close()
