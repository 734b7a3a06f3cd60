"""Frame-by-frame parity of the C++ host pipeline (flv::F2FTracking over the GPU C ABI) with the Python restatement
of the reference pipeline (oracle/f2f_ref.py).  The two OpenCV RANSAC calls are injected through the tracker's
hooks (hard part 1): both sides get cv2's masks for the same float inputs, so landmark id
lists must be bit-exact and poses agree to the fp64 BA tolerance."""
import ctypes as C

import numpy as np
import pytest

from oracle import f2f_ref
from oracle.vimotion_ref import SE3

pytestmark = pytest.mark.gpu


class Cfg(C.Structure):
    _fields_ = [("cam_type", C.c_int), ("img_w", C.c_int), ("img_h", C.c_int), ("cam0", C.c_double * 4),
                ("cam1", C.c_double * 4), ("depth_scale", C.c_double), ("P0", C.c_double * 12), ("P1", C.c_double * 12),
                ("T_cam1_cam0", C.c_double * 7), ("T_i_c0", C.c_double * 7), ("feature_para", C.c_double * 6),
                ("vi_para", C.c_double * 6), ("dc_para", C.c_double * 3), ("skip_first_n_imgs", C.c_int)]


FMAT = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_uint8))
PNP = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_double), C.c_int,
                  C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_int))


@FMAT
def fmat_hook(user, n, a, b, mask):
    fa = np.ctypeslib.as_array(a, (n, 2)).copy(); fb = np.ctypeslib.as_array(b, (n, 2)).copy()
    m = f2f_ref.fmat_cv2(fa, fb)
    for i in range(n):
        mask[i] = int(m[i])
    return 0


@PNP
def pnp_hook(user, n, p3d, p2d, K, use_guess, T, inl, ninl):
    a3 = np.ctypeslib.as_array(p3d, (n, 3)).copy(); a2 = np.ctypeslib.as_array(p2d, (n, 2)).copy()
    Kt = tuple(K[i] for i in range(4))
    guess = SE3.from7(np.array([T[i] for i in range(7)])) if use_guess else None
    Tn, idx = f2f_ref.pnp_cv2(a3, a2, Kt, guess)
    t7 = Tn.to7()
    for i in range(7):
        T[i] = float(t7[i])
    for i, v in enumerate(idx):
        inl[i] = int(v)
    ninl[0] = len(idx)
    return 0


def _setup(lib):
    lib.flv_f2f_create.restype = C.c_void_p
    lib.flv_f2f_create.argtypes = [C.POINTER(Cfg), C.c_int]
    lib.flv_f2f_destroy.argtypes = [C.c_void_p]
    lib.flv_f2f_last_error.restype = C.c_char_p
    lib.flv_f2f_last_error.argtypes = [C.c_void_p]
    lib.flv_f2f_set_ransac_hooks.argtypes = [C.c_void_p, FMAT, PNP, C.c_void_p]
    lib.flv_f2f_image_feed.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
    lib.flv_f2f_state.argtypes = [C.c_void_p]
    lib.flv_f2f_get_frame.argtypes = [C.c_void_p] + [C.c_void_p] * 7 + [C.c_int]
    lib.flv_f2f_tracking_counts.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3


def _cam_centre(T7):
    """camera centre in the world frame of a T_c_w pose [qx qy qz qw tx ty tz]: -R^T t."""
    x, y, z, w = T7[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return -R.T @ np.asarray(T7[4:], float)


def test_depth_sequence_matches_oracle_frame_by_frame(lib):
    _setup(lib)
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.98, 40.0, 1.0]
    n_frames = 12
    imgs, depths = f2f_ref.make_depth_sequence(n_frames, seed=7)
    cfg = Cfg(0, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(), (C.c_double * 12)(),
              (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 6)(*fpara),
              (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    assert h and lib.flv_f2f_last_error(h) == b"", lib.flv_f2f_last_error(h)
    lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
    ref = f2f_ref.F2FTracking("depth", 640, 480, K, fpara, vpara, dpara)
    cap = 600
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_kf = 0
    traj, traj_ref = [], []
    for k in range(n_frames):
        kf = C.c_int(0); rs = C.c_int(0)
        img = np.ascontiguousarray(imgs[k]); dep = np.ascontiguousarray(depths[k])
        rc = lib.flv_f2f_image_feed(h, 0.05 * k, vp(img), vp(dep), C.byref(kf), C.byref(rs))
        assert rc == 0, lib.flv_f2f_last_error(h)
        rkf, rrs = ref.image_feed(0.05 * k, imgs[k], depths[k])
        state = {0: "UnInit", 1: "Tracking", 2: "TrackingFail"}[lib.flv_f2f_state(h)]
        assert state == ref.state and bool(kf.value) == rkf and bool(rs.value) == rrs, (k, state, ref.state)
        T = np.zeros(7); ids = np.zeros(cap, np.int64); pl = np.zeros((cap, 2)); un = np.zeros((cap, 2)); p3 = np.zeros((cap, 3))
        has = np.zeros(cap, np.uint8); inl = np.zeros(cap, np.uint8)
        n = lib.flv_f2f_get_frame(h, vp(T), vp(ids), vp(pl), vp(un), vp(p3), vp(has), vp(inl), cap)
        cur = ref.curr
        assert n == len(cur.lms), (k, n, len(cur.lms))
        assert list(ids[:n]) == [l.lm_id for l in cur.lms]                       # landmark ids + order: bit-exact
        assert list(inl[:n].astype(bool)) == [bool(l.inlier) for l in cur.lms]
        assert list(has[:n].astype(bool)) == [bool(l.has_3d) for l in cur.lms]
        if n:
            assert np.array_equal(pl[:n], np.array([l.plane for l in cur.lms]))   # tracked pixel positions: bit-exact (LK contract)
            assert np.abs(p3[:n] - np.array([l.p3d_w for l in cur.lms])).max() <= 1e-6
        rT = cur.T_c_w.to7()
        assert np.abs(T[4:] - rT[4:]).max() <= 1e-6                               # per-frame pose: 1e-6 m
        assert 2 * np.arccos(min(1.0, abs(float(np.dot(T[:4], rT[:4]))))) <= 1e-6
        traj.append(_cam_centre(T)); traj_ref.append(_cam_centre(rT))
        if k > 0 and ref.state == "Tracking":
            of = C.c_int(); fi = C.c_int(); pn = C.c_int()
            lib.flv_f2f_tracking_counts(h, C.byref(of), C.byref(fi), C.byref(pn))
            assert (of.value, fi.value, pn.value) == ref.counts
            assert pn.value >= 30                                                 # the synthetic plane is trackable
        n_kf += int(rkf)
    assert ref.state == "Tracking" and n_kf >= 2
    # absolute trajectory error against the reference path (BASELINE: ATE within 1 % of the reference's)
    traj, traj_ref = np.array(traj), np.array(traj_ref)
    ate = float(np.sqrt(np.mean(np.sum((traj - traj_ref) ** 2, axis=1))))
    path = float(np.sum(np.linalg.norm(np.diff(traj_ref, axis=0), axis=1)))
    assert path > 0.05 and ate <= 1e-6 and ate <= 0.01 * path
    lib.flv_f2f_destroy(h)


def test_stereo_rect_sequence_matches_oracle(lib):
    """STEREO_RECT: exercises the left->right LK + DLT triangulation path of depthInnovation inside the pipeline."""
    _setup(lib)
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.9, 50.0, 0.0]
    n_frames = 6
    L, R, P0, P1, T10 = f2f_ref.make_stereo_sequence(n_frames, seed=5)
    cfg = Cfg(1, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(*P0.ravel()),
              (C.c_double * 12)(*P1.ravel()), (C.c_double * 7)(*T10.to7()), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0),
              (C.c_double * 6)(*fpara), (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    assert h and lib.flv_f2f_last_error(h) == b""
    lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
    ref = f2f_ref.F2FTracking("stereo", 640, 480, K, fpara, vpara, dpara, K1=K, P0=P0, P1=P1, T_c1_c0=T10)
    cap = 600
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in range(n_frames):
        kf = C.c_int(0); rs = C.c_int(0)
        rc = lib.flv_f2f_image_feed(h, 0.05 * k, vp(np.ascontiguousarray(L[k])), vp(np.ascontiguousarray(R[k])), C.byref(kf), C.byref(rs))
        assert rc == 0, lib.flv_f2f_last_error(h)
        rkf, rrs = ref.image_feed(0.05 * k, L[k], R[k])
        assert {0: "UnInit", 1: "Tracking", 2: "TrackingFail"}[lib.flv_f2f_state(h)] == ref.state
        T = np.zeros(7); ids = np.zeros(cap, np.int64); p3 = np.zeros((cap, 3)); has = np.zeros(cap, np.uint8)
        n = lib.flv_f2f_get_frame(h, vp(T), vp(ids), None, None, vp(p3), vp(has), None, cap)
        cur = ref.curr
        assert n == len(cur.lms) and list(ids[:n]) == [l.lm_id for l in cur.lms]
        if n:
            ref3 = np.array([l.p3d_w for l in cur.lms])
            assert np.abs(p3[:n] - ref3).max() <= 1e-6 * max(1.0, np.abs(ref3).max())
        rT = cur.T_c_w.to7()
        assert np.abs(T[4:] - rT[4:]).max() <= 1e-6
    assert ref.state == "Tracking"
    # the triangulated depth of the plane is ~3 m
    z = np.array([l.p3d_c[2] for l in ref.curr.lms])
    assert abs(np.median(z) - 3.0) < 0.1
    lib.flv_f2f_destroy(h)


def test_stereo_unrect_equalize_sequence_matches_oracle(lib):
    """STEREO_UNRECT (EuRoC raw): LK on the raw images, cv::undistortPoints / cv::projectPoints per point with raw lens
    models, cv::equalizeHist on ingest -- frame by frame against oracle/f2f_ref.py (which calls cv2 for all three)."""
    import cv2
    _setup(lib)
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.9, 50.0, 0.0]
    n_frames = 6
    L, R, P0, P1, T10 = f2f_ref.make_stereo_sequence(n_frames, seed=9)
    # the images are rectified views; treating them as raw with mild lens models keeps the geometry consistent to ~1 px
    Km = np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]])
    K0r = Km + np.array([[1.3, 0, -0.8], [0, 0.9, 0.6], [0, 0, 0]])
    K1r = Km + np.array([[-0.7, 0, 0.5], [0, 1.1, -0.4], [0, 0, 0]])
    D0 = np.array([-0.004, 0.0015, 1e-4, -8e-5]); D1 = np.array([-0.0035, 0.001, -6e-5, 5e-5])
    R0, _ = cv2.Rodrigues(np.array([4e-4, -6e-4, 3e-4])); R1, _ = cv2.Rodrigues(np.array([-3e-4, 5e-4, -2e-4]))
    cfg = Cfg(2, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(*P0.ravel()),
              (C.c_double * 12)(*P1.ravel()), (C.c_double * 7)(*T10.to7()), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0),
              (C.c_double * 6)(*fpara), (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    assert h and lib.flv_f2f_last_error(h) == b""
    vp = lambda a: np.ascontiguousarray(a, np.float64).ctypes.data_as(C.c_void_p)
    d14 = lambda D: np.concatenate([D, np.zeros(14 - len(D))])
    lib.flv_f2f_set_lens.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    lib.flv_f2f_set_equalize_hist.argtypes = [C.c_void_p, C.c_int]
    for cam, (Kr, D, Rr) in enumerate([(K0r, D0, R0), (K1r, D1, R1)]):
        assert lib.flv_f2f_set_lens(h, cam, vp([Kr[0, 0], Kr[1, 1], Kr[0, 2], Kr[1, 2]]), vp(d14(D)), vp(Rr.ravel())) == 0
    assert lib.flv_f2f_set_equalize_hist(h, 1) == 0
    lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
    ref = f2f_ref.F2FTracking("stereo_unrect", 640, 480, K, fpara, vpara, dpara, K1=K, P0=P0, P1=P1, T_c1_c0=T10,
                              lens0=(K0r, D0, R0), lens1=(K1r, D1, R1), equalize=True)
    cap = 600
    vq = lambda a: a.ctypes.data_as(C.c_void_p)
    for k in range(n_frames):
        kf = C.c_int(0); rs = C.c_int(0)
        rc = lib.flv_f2f_image_feed(h, 0.05 * k, vq(np.ascontiguousarray(L[k])), vq(np.ascontiguousarray(R[k])), C.byref(kf), C.byref(rs))
        assert rc == 0, lib.flv_f2f_last_error(h)
        ref.image_feed(0.05 * k, L[k], R[k])
        assert {0: "UnInit", 1: "Tracking", 2: "TrackingFail"}[lib.flv_f2f_state(h)] == ref.state
        T = np.zeros(7); ids = np.zeros(cap, np.int64); p3 = np.zeros((cap, 3)); und = np.zeros((cap, 2)); pl = np.zeros((cap, 2))
        n = lib.flv_f2f_get_frame(h, vq(T), vq(ids), vq(pl), vq(und), vq(p3), None, None, cap)
        cur = ref.curr
        assert n == len(cur.lms) and list(ids[:n]) == [l.lm_id for l in cur.lms]
        if n:
            assert np.array_equal(pl[:n], np.array([l.plane for l in cur.lms]))                 # LK positions: bit-exact
            assert np.abs(und[:n] - np.array([l.undist for l in cur.lms])).max() <= 6.2e-5      # cv2 float rounding
            ref3 = np.array([l.p3d_w for l in cur.lms])
            assert np.abs(p3[:n] - ref3).max() <= 1e-5 * max(1.0, np.abs(ref3).max())
        assert np.abs(T[4:] - cur.T_c_w.to7()[4:]).max() <= 1e-5
    assert ref.state == "Tracking"
    und_shift = np.abs(np.array([l.undist for l in ref.curr.lms]) - np.array([l.plane for l in ref.curr.lms])).max()
    assert und_shift > 0.3          # the lens model really moves points
    lib.flv_f2f_destroy(h)


def test_system_tracker_keyframes_feed_local_map(lib):
    """Both hot paths chained as in FLVIS: the tracker's keyframes (KeyFrame message = ids / undistorted 2-d / world 3-d
    of the inlier landmarks with depth + T_c_w, keyframe_msg.cpp:30-110) drive the sliding-window local map; every
    CorrectionInf is compared with the chained oracles (f2f_ref -> localmap_ref)."""
    from flvis_b200 import capi
    from oracle import localmap_ref
    _setup(lib)
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.9, 50.0, 0.0]
    n_frames, window = 26, 5
    imgs, depths = f2f_ref.make_depth_sequence(n_frames, seed=21)
    cfg = Cfg(0, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(), (C.c_double * 12)(),
              (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0),
              (C.c_double * 6)(*fpara), (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    assert h and lib.flv_f2f_last_error(h) == b""
    lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
    ref = f2f_ref.F2FTracking("depth", 640, 480, K, fpara, vpara, dpara)
    ctx = capi.Context(1, 640, 480)
    clib = ctx.lib
    clib.flv_localmap_create.restype = C.c_void_p
    clib.flv_localmap_create.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
    clib.flv_localmap_destroy.argtypes = [C.c_void_p]
    clib.flv_localmap_add_keyframe.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 5 + \
        [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(capi.BAStats)]
    lm = clib.flv_localmap_create(ctx.h, window, *K)
    lm_ref = localmap_ref.LocalMap(window, K)
    cap = 600
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_solved = 0
    for k in range(n_frames):
        kf = C.c_int(0); rs = C.c_int(0)
        assert lib.flv_f2f_image_feed(h, 0.05 * k, vp(np.ascontiguousarray(imgs[k])), vp(np.ascontiguousarray(depths[k])),
                                      C.byref(kf), C.byref(rs)) == 0
        rkf, _ = ref.image_feed(0.05 * k, imgs[k], depths[k])
        assert bool(kf.value) == rkf
        if not rkf:
            continue
        T = np.zeros(7); ids = np.zeros(cap, np.int64); und = np.zeros((cap, 2)); p3 = np.zeros((cap, 3))
        has = np.zeros(cap, np.uint8); inl = np.zeros(cap, np.uint8)
        n = lib.flv_f2f_get_frame(h, vp(T), vp(ids), None, vp(und), vp(p3), vp(has), vp(inl), cap)
        sel = (has[:n] == 1) & (inl[:n] == 1)                                   # CameraFrame::getKeyFrameInf
        kids = np.ascontiguousarray(ids[:n][sel]); kuv = np.ascontiguousarray(und[:n][sel]); k3 = np.ascontiguousarray(p3[:n][sel])
        rsel = [l for l in ref.curr.lms if l.has_3d and l.inlier]
        okf = {"frame_id": ref.curr.frame_id, "lm_id": [l.lm_id for l in rsel], "lm_2d": np.array([l.undist for l in rsel]),
               "lm_3d": np.array([l.p3d_w for l in rsel]), "T_c_w": ref.curr.T_c_w.to7()}
        assert list(kids) == okf["lm_id"]
        o = lm_ref.frame_callback(okf)
        fid = np.zeros(1, np.int64); oT = np.zeros(7); nlm = np.zeros(1, np.int32); olm = np.zeros(8192, np.int64)
        o3 = np.zeros((8192, 3)); nout = np.zeros(1, np.int32); oout = np.zeros(8192, np.int64)
        st = capi.BAStats()
        rc = clib.flv_localmap_add_keyframe(lm, int(ref.curr.frame_id), len(kids), vp(kids), vp(kuv), vp(k3), vp(T), vp(fid), vp(oT),
                                            vp(nlm), vp(olm), vp(o3), 8192, vp(nout), vp(oout), 8192, C.byref(st))
        if o is None:
            assert rc == 0
            continue
        assert rc == 1, clib.flv_last_error(ctx.h)
        n_solved += 1
        assert fid[0] == o["frame_id"] and list(olm[:nlm[0]]) == o["lm_id"]
        assert sorted(oout[:nout[0]]) == sorted(o["outlier_id"])
        assert np.abs(oT[4:] - o["T_c_w"][4:]).max() <= 1e-5
        if nlm[0]:
            assert np.abs(o3[:nlm[0]] - o["lm_3d"]).max() <= 1e-4
    assert n_solved >= 2
    clib.flv_localmap_destroy(lm)
    ctx.close()
    lib.flv_f2f_destroy(h)


def test_builtin_ransac_tracks_without_hooks(lib):
    """The product's own host RANSAC stand-ins (no OpenCV): the sequence must track with most points as inliers and
    the recovered camera translation must follow the synthetic motion (12 mm per frame along the plane)."""
    _setup(lib)
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.98, 40.0, 1.0]
    imgs, depths = f2f_ref.make_depth_sequence(8, seed=9)
    cfg = Cfg(0, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(), (C.c_double * 12)(),
              (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 6)(*fpara),
              (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    T0 = None
    for k in range(8):
        kf = C.c_int(0); rs = C.c_int(0)
        assert lib.flv_f2f_image_feed(h, 0.05 * k, vp(np.ascontiguousarray(imgs[k])), vp(np.ascontiguousarray(depths[k])),
                                      C.byref(kf), C.byref(rs)) == 0
        assert lib.flv_f2f_state(h) == 1
        T = np.zeros(7)
        lib.flv_f2f_get_frame(h, vp(T), None, None, None, None, None, None, 0)
        if k == 0:
            T0 = T.copy()
        if k > 0:
            of = C.c_int(); fi = C.c_int(); pn = C.c_int()
            lib.flv_f2f_tracking_counts(h, C.byref(of), C.byref(fi), C.byref(pn))
            assert pn.value >= 0.8 * fi.value >= 0.5 * of.value > 50
    # camera centre moved ~7 * 12 mm in total
    c0 = -SE3.from7(T0).inverse().t if False else SE3.from7(T0).inverse().t
    c7 = SE3.from7(T).inverse().t
    assert abs(np.linalg.norm(c7 - c0) - 0.012 * 7) < 0.01
    lib.flv_f2f_destroy(h)


def test_correction_feed_applies_local_map_feedback_like_the_reference(lib):
    """F2FTracking::correction_feed + STEP1 of image_feed (src/frontend/f2f_tracking.cpp:40-44, :189-219), which the reference's
    own nodelet never wires: every CorrectionInf of the chained local map is fed back to the tracker on both sides; poses are
    re-based, the last frame's landmarks take the optimised world points, the local map's outliers are flagged -- frame by
    frame against the oracle with the same feedback."""
    from flvis_b200 import capi
    from oracle import localmap_ref
    _setup(lib)
    lib.flv_f2f_correction_feed.argtypes = [C.c_void_p, C.c_double, C.c_int64, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    K = (384.16455, 384.16455, 320.21445, 238.94403)
    fpara = [30, 15, 5, 500, 0.01, 15]; vpara = [0.1, 0.01, 0.001, 0.001, 0.5, 0.1]; dpara = [0.9, 50.0, 0.0]
    n_frames, window = 30, 4
    imgs, depths = f2f_ref.make_depth_sequence(n_frames, seed=33)
    cfg = Cfg(0, 640, 480, (C.c_double * 4)(*K), (C.c_double * 4)(*K), 1000.0, (C.c_double * 12)(), (C.c_double * 12)(),
              (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0), (C.c_double * 7)(0, 0, 0, 1, 0, 0, 0),
              (C.c_double * 6)(*fpara), (C.c_double * 6)(*vpara), (C.c_double * 3)(*dpara), 0)
    h = lib.flv_f2f_create(C.byref(cfg), 0)
    assert h and lib.flv_f2f_last_error(h) == b""
    lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
    ref = f2f_ref.F2FTracking("depth", 640, 480, K, fpara, vpara, dpara)
    ctx = capi.Context(1, 64, 64)
    clib = ctx.lib
    clib.flv_localmap_create.restype = C.c_void_p
    clib.flv_localmap_create.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
    clib.flv_localmap_destroy.argtypes = [C.c_void_p]
    clib.flv_localmap_add_keyframe.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 5 + \
        [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(capi.BAStats)]
    lm = clib.flv_localmap_create(ctx.h, window, *K)
    lm_ref = localmap_ref.LocalMap(window, K)
    cap = 600
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    n_fed = 0
    for k in range(n_frames):
        kf = C.c_int(0); rs = C.c_int(0)
        assert lib.flv_f2f_image_feed(h, 0.05 * k, vp(np.ascontiguousarray(imgs[k])), vp(np.ascontiguousarray(depths[k])),
                                      C.byref(kf), C.byref(rs)) == 0
        rkf, _ = ref.image_feed(0.05 * k, imgs[k], depths[k])
        assert bool(kf.value) == rkf
        T = np.zeros(7); ids = np.zeros(cap, np.int64); und = np.zeros((cap, 2)); p3 = np.zeros((cap, 3))
        has = np.zeros(cap, np.uint8); inl = np.zeros(cap, np.uint8)
        n = lib.flv_f2f_get_frame(h, vp(T), vp(ids), None, vp(und), vp(p3), vp(has), vp(inl), cap)
        cur = ref.curr
        assert n == len(cur.lms) and list(ids[:n]) == [l.lm_id for l in cur.lms], k
        assert list(inl[:n].astype(bool)) == [bool(l.inlier) for l in cur.lms], k          # the fed-back outlier flags show up here
        if n:
            assert np.abs(p3[:n] - np.array([l.p3d_w for l in cur.lms])).max() <= 1e-4     # and the fed-back world points here
        assert np.abs(T[4:] - cur.T_c_w.to7()[4:]).max() <= 1e-5, k
        if not rkf:
            continue
        sel = (has[:n] == 1) & (inl[:n] == 1)
        kids = np.ascontiguousarray(ids[:n][sel]); kuv = np.ascontiguousarray(und[:n][sel]); k3 = np.ascontiguousarray(p3[:n][sel])
        rsel = [l for l in cur.lms if l.has_3d and l.inlier]
        okf = {"frame_id": cur.frame_id, "lm_id": [l.lm_id for l in rsel], "lm_2d": np.array([l.undist for l in rsel]),
               "lm_3d": np.array([l.p3d_w for l in rsel]), "T_c_w": cur.T_c_w.to7()}
        o = lm_ref.frame_callback(okf)
        fid = np.zeros(1, np.int64); oT = np.zeros(7); nlm = np.zeros(1, np.int32); olm = np.zeros(8192, np.int64)
        o3 = np.zeros((8192, 3)); nout = np.zeros(1, np.int32); oout = np.zeros(8192, np.int64)
        st = capi.BAStats()
        rc = clib.flv_localmap_add_keyframe(lm, int(cur.frame_id), len(kids), vp(kids), vp(kuv), vp(k3), vp(T), vp(fid), vp(oT),
                                            vp(nlm), vp(olm), vp(o3), 8192, vp(nout), vp(oout), 8192, C.byref(st))
        if o is None:
            assert rc == 0
            continue
        assert rc == 1
        # feed both trackers with their own local map's CorrectionInf
        assert lib.flv_f2f_correction_feed(h, 0.05 * k, int(fid[0]), vp(oT), int(nlm[0]), vp(olm), vp(o3), int(nout[0]), vp(oout)) == 0
        ref.correction_feed(dict(frame_id=o["frame_id"], T_c_w=o["T_c_w"], lm_id=o["lm_id"], lm_3d=o["lm_3d"], outlier_id=o["outlier_id"]))
        n_fed += 1
    assert n_fed >= 3 and ref.state == "Tracking"
    clib.flv_localmap_destroy(lm)
    ctx.close()
    lib.flv_f2f_destroy(h)


def test_create_stereo_equals_create_plus_set_lens(lib):
    """flv_f2f_create_stereo (DepthCamera::setSteroCamInfo + F2FTracking::init, vo_tracking.cpp:245-268) builds the same tracker
    as flv_f2f_create + flv_f2f_set_lens + flv_f2f_set_equalize_hist: identical frames on an EuRoC-shaped raw stereo sequence."""
    from flvis_b200 import batch
    from synthdata import sequences
    from .seq_harness import SingleTrackers
    seq = sequences.make_c1(14)
    a = SingleTrackers(lib, [seq], hooks=False)
    lib.flv_f2f_create_stereo.restype = C.c_void_p
    lib.flv_f2f_create_stereo.argtypes = [C.POINTER(batch.F2FStereoConfig), C.c_int]
    sc = batch.stereo_config_for(seq)
    hb = lib.flv_f2f_create_stereo(C.byref(sc), 0)
    assert hb and lib.flv_f2f_last_error(hb) == b""
    n_checked = 0
    for t_img, img0, img1, imu in seq.frames():
        for (t, acc, gyro) in imu:
            acc = np.ascontiguousarray(acc, np.float64); gyro = np.ascontiguousarray(gyro, np.float64)
            a.imu_feed(0, t, acc, gyro)
            assert lib.flv_f2f_imu_feed(hb, t, acc.ctypes.data_as(C.c_void_p), gyro.ctypes.data_as(C.c_void_p)) == 0
        a.image_feed([t_img], [img0], [img1])
        kf = C.c_int(0); rs = C.c_int(0)
        assert lib.flv_f2f_image_feed(hb, float(t_img), np.ascontiguousarray(img0).ctypes.data_as(C.c_void_p),
                                      np.ascontiguousarray(img1).ctypes.data_as(C.c_void_p), C.byref(kf), C.byref(rs)) == 0
        fa = a.get_frame(0)
        a.h.append(hb); fb = a.get_frame(1); a.h.pop()
        assert fa[0] == fb[0]
        for x, y in zip(fa[1:], fb[1:]):
            assert np.array_equal(x, y)
        n_checked += fa[0] > 0
    assert n_checked >= 6
    lib.flv_f2f_destroy(hb)
