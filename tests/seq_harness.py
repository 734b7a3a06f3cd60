"""Drives flv::F2FTracking (C handle, GPU) and oracle/f2f_ref.py side by side over a BASELINE-shaped sequence
(oracle/sequences.py): IMU samples first, then the image pair, exactly the order TrackingNodeletClass delivers them
(/root/reference/src/frontend/vo_tracking.cpp:326-371, :387-429); optional local map chained on the keyframes
(vo_localmap.cpp:87-380).  Shared by tests/test_configs_gpu.py."""
import ctypes as C

import cv2
import numpy as np

from oracle import f2f_ref, localmap_ref
from oracle.vimotion_ref import SE3 as OSE3, q2R

from .test_pipeline_gpu import FMAT, PNP, Cfg, _cam_centre, _setup, fmat_hook, pnp_hook

STATE = {0: "UnInit", 1: "Tracking", 2: "TrackingFail"}


def _vp(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(vals, n):
    return (C.c_double * n)(*[float(v) for v in vals])


def stereo_rectify(cfg, T_c1_c0):
    """cv::stereoRectify as vo_tracking.cpp:236-247 calls it (CALIB_ZERO_DISPARITY, alpha 0, same size)."""
    K0 = np.array([[cfg["K0"][0], 0, cfg["K0"][2]], [0, cfg["K0"][1], cfg["K0"][3]], [0, 0, 1.0]])
    K1 = np.array([[cfg["K1"][0], 0, cfg["K1"][2]], [0, cfg["K1"][1], cfg["K1"][3]], [0, 0, 1.0]])
    D0, D1 = np.array(cfg["D0"]), np.array(cfg["D1"])
    size = (cfg["w"], cfg["h"])
    R0, R1, P0, P1, _, _, _ = cv2.stereoRectify(K0, D0, K1, D1, size, q2R(T_c1_c0.q), T_c1_c0.t.reshape(3, 1),
                                                flags=cv2.CALIB_ZERO_DISPARITY, alpha=0, newImageSize=size)
    return K0, D0, R0, P0, K1, D1, R1, P1


def make_cfg(seq):
    """-> (Cfg struct, [(K4, D14, R9) per camera] or None, equalize flag, K of the rectified camera, oracle factory) configured
    like TrackingNodeletClass::onInit does for the sequence's sensor (vo_tracking.cpp:140-306)."""
    c = seq.cfg
    Ti = seq.T_i_c0.to7()
    T_i_c0 = OSE3.from7(Ti)                                   # the oracle's own SE3 type (arithmetic of sophus, not synthdata's)
    lenses, equalize = None, False
    if seq.cam_type == "depth":
        K = c["K0"]
        cfg = Cfg(0, c["w"], c["h"], _c(K, 4), _c(K, 4), c["depth_factor"], (C.c_double * 12)(), (C.c_double * 12)(),
                  _c([0, 0, 0, 1, 0, 0, 0], 7), _c(Ti, 7), _c(c["feature_para"], 6), _c(c["vi_para"], 6), _c(c["dc_para"], 3), c["skip"])
        mk = lambda: f2f_ref.F2FTracking("depth", c["w"], c["h"], K, c["feature_para"], c["vi_para"], c["dc_para"], T_i_c=T_i_c0,
                                         skip=c["skip"], depth_scale=c["depth_factor"])
    elif seq.cam_type == "stereo_unrect":
        T10 = OSE3.from7(seq.T_c0_c1.to7()).inverse()
        K0, D0, R0, P0, K1, D1, R1, P1 = stereo_rectify(c, T10)
        K = (P0[0, 0], P0[1, 1], P0[0, 2], P0[1, 2]); Kr1 = (P1[0, 0], P1[1, 1], P1[0, 2], P1[1, 2])     # depth_camera.cpp:76-84
        cfg = Cfg(2, c["w"], c["h"], _c(K, 4), _c(Kr1, 4), 1000.0, _c(P0.ravel(), 12), _c(P1.ravel(), 12), _c(T10.to7(), 7),
                  _c(Ti, 7), _c(c["feature_para"], 6), _c(c["vi_para"], 6), _c(c["dc_para"], 3), 0)
        d14 = lambda D: np.concatenate([D, np.zeros(14 - len(D))])
        lenses = [(np.array([Kr[0, 0], Kr[1, 1], Kr[0, 2], Kr[1, 2]]), d14(D), np.ascontiguousarray(Rr).ravel().copy())
                  for Kr, D, Rr in ((K0, D0, R0), (K1, D1, R1))]
        equalize = True                                       # vo_tracking.cpp:257-263: need_equal_hist = true
        mk = lambda: f2f_ref.F2FTracking("stereo_unrect", c["w"], c["h"], K, c["feature_para"], c["vi_para"], c["dc_para"],
                                         T_i_c=T_i_c0, K1=Kr1, P0=P0, P1=P1, T_c1_c0=T10, lens0=(K0, D0, R0), lens1=(K1, D1, R1),
                                         equalize=True)
    else:
        K = c["K0"]
        P0 = np.array([[K[0], 0, K[2], 0], [0, K[1], K[3], 0], [0, 0, 1, 0.0]]); P1 = P0.copy(); P1[0, 3] = -c["bf"]
        T10 = OSE3.from7(seq.T_c0_c1.to7()).inverse()
        cfg = Cfg(1, c["w"], c["h"], _c(K, 4), _c(K, 4), 1000.0, _c(P0.ravel(), 12), _c(P1.ravel(), 12), _c(T10.to7(), 7),
                  _c(Ti, 7), _c(c["feature_para"], 6), _c(c["vi_para"], 6), _c(c["dc_para"], 3), 0)
        mk = lambda: f2f_ref.F2FTracking("stereo", c["w"], c["h"], K, c["feature_para"], c["vi_para"], c["dc_para"], K1=K, P0=P0, P1=P1,
                                         T_c1_c0=T10)
    return cfg, lenses, equalize, K, mk


class SingleTrackers:
    """N independent flv::F2FTracking handles behind the interface the harness drives."""

    def __init__(self, lib, seqs, hooks):
        _setup(lib)
        lib.flv_f2f_set_lens.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.flv_f2f_set_equalize_hist.argtypes = [C.c_void_p, C.c_int]
        lib.flv_f2f_imu_feed.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p]
        lib.flv_f2f_get_imu_bias.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        lib.flv_f2f_get_frame_ex.argtypes = [C.c_void_p] + [C.c_void_p] * 4 + [C.c_int]
        lib.flv_f2f_get_imu_states.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self.lib, self.h, self.refs, self.K = lib, [], [], None
        for seq in seqs:
            cfg, lenses, equalize, K, mk = make_cfg(seq)
            h = lib.flv_f2f_create(C.byref(cfg), 0)
            assert h and lib.flv_f2f_last_error(h) == b"", lib.flv_f2f_last_error(h)
            if lenses:
                for cam, (k4, d14, r9) in enumerate(lenses):
                    assert lib.flv_f2f_set_lens(h, cam, _vp(k4), _vp(d14), _vp(r9)) == 0
            if equalize:
                assert lib.flv_f2f_set_equalize_hist(h, 1) == 0
            if hooks:
                lib.flv_f2f_set_ransac_hooks(h, fmat_hook, pnp_hook, None)
            self.h.append(h); self.refs.append(mk()); self.K = K; self.mk = mk

    def imu_feed(self, s, t, a, g):
        assert self.lib.flv_f2f_imu_feed(self.h[s], t, _vp(a), _vp(g)) == 0

    def image_feed(self, ts, imgs0, imgs1):
        out = []
        for s, h in enumerate(self.h):
            kf = C.c_int(0); rs = C.c_int(0)
            rc = self.lib.flv_f2f_image_feed(h, float(ts[s]), _vp(np.ascontiguousarray(imgs0[s])), _vp(np.ascontiguousarray(imgs1[s])),
                                             C.byref(kf), C.byref(rs))
            assert rc == 0, self.lib.flv_f2f_last_error(h)
            out.append((bool(kf.value), bool(rs.value)))
        return out

    def state(self, s):
        return STATE[self.lib.flv_f2f_state(self.h[s])]

    def get_frame(self, s, cap=600):
        T = np.zeros(7); ids = np.zeros(cap, np.int64); pl = np.zeros((cap, 2)); un = np.zeros((cap, 2)); p3 = np.zeros((cap, 3))
        has = np.zeros(cap, np.uint8); inl = np.zeros(cap, np.uint8)
        n = self.lib.flv_f2f_get_frame(self.h[s], _vp(T), _vp(ids), _vp(pl), _vp(un), _vp(p3), _vp(has), _vp(inl), cap)
        return n, T, ids, pl, un, p3, has, inl

    def get_ex(self, s, n):
        cap = max(n, 1)
        p3c = np.zeros((cap, 3)); f2d = np.zeros((cap, 2)); fp = np.zeros((cap, 7)); Tkf = np.zeros(7)
        m = self.lib.flv_f2f_get_frame_ex(self.h[s], _vp(p3c), _vp(f2d), _vp(fp), _vp(Tkf), cap)
        return m, p3c, f2d, fp, Tkf

    def imu_states(self, s):
        st = np.zeros((400, 11))
        return st[:self.lib.flv_f2f_get_imu_states(self.h[s], _vp(st), 400)]

    def bias(self, s):
        ab = np.zeros(3); gb = np.zeros(3)
        return self.lib.flv_f2f_get_imu_bias(self.h[s], _vp(ab), _vp(gb)), ab, gb

    def counts(self, s):
        of = C.c_int(); fi = C.c_int(); pn = C.c_int()
        self.lib.flv_f2f_tracking_counts(self.h[s], C.byref(of), C.byref(fi), C.byref(pn))
        return of.value, fi.value, pn.value

    def close(self):
        for h in self.h:
            self.lib.flv_f2f_destroy(h)


class BatchTracker:
    """One flv_f2f_batch handle advancing len(seqs) streams of the same sensor together."""

    def __init__(self, lib, seqs, hooks, groups=1):
        _setup(lib)
        vp = C.c_void_p
        lib.flv_f2f_batch_create_grouped.restype = vp
        lib.flv_f2f_batch_create_grouped.argtypes = [C.POINTER(Cfg), C.c_int, C.c_int, C.c_int]
        lib.flv_f2f_batch_create.restype = vp
        lib.flv_f2f_batch_create.argtypes = [C.POINTER(Cfg), C.c_int, C.c_int]
        lib.flv_f2f_batch_destroy.argtypes = [vp]
        lib.flv_f2f_batch_last_error.restype = C.c_char_p
        lib.flv_f2f_batch_last_error.argtypes = [vp]
        lib.flv_f2f_batch_set_lens.argtypes = [vp, C.c_int, vp, vp, vp]
        lib.flv_f2f_batch_set_equalize_hist.argtypes = [vp, C.c_int]
        lib.flv_f2f_batch_set_ransac_hooks.argtypes = [vp, FMAT, PNP, vp]
        lib.flv_f2f_batch_imu_feed.argtypes = [vp, C.c_int, C.c_double, vp, vp]
        lib.flv_f2f_batch_image_feed.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
        lib.flv_f2f_batch_state.argtypes = [vp, C.c_int]
        lib.flv_f2f_batch_get_frame.argtypes = [vp, C.c_int] + [vp] * 7 + [C.c_int]
        lib.flv_f2f_batch_get_frame_ex.argtypes = [vp, C.c_int] + [vp] * 4 + [C.c_int]
        lib.flv_f2f_batch_get_imu_states.argtypes = [vp, C.c_int, vp, C.c_int]
        lib.flv_f2f_batch_get_imu_bias.argtypes = [vp, C.c_int, vp, vp]
        lib.flv_f2f_batch_tracking_counts.argtypes = [vp, C.c_int] + [C.POINTER(C.c_int)] * 3
        self.lib, self.S = lib, len(seqs)
        cfg, lenses, equalize, K, mk = make_cfg(seqs[0])
        self.K, self.mk = K, mk
        self.b = lib.flv_f2f_batch_create_grouped(C.byref(cfg), self.S, 0, groups)
        assert self.b and lib.flv_f2f_batch_last_error(self.b) == b"", lib.flv_f2f_batch_last_error(self.b)
        if lenses:
            for cam, (k4, d14, r9) in enumerate(lenses):
                assert lib.flv_f2f_batch_set_lens(self.b, cam, _vp(k4), _vp(d14), _vp(r9)) == 0
        if equalize:
            assert lib.flv_f2f_batch_set_equalize_hist(self.b, 1) == 0
        if hooks:
            lib.flv_f2f_batch_set_ransac_hooks(self.b, fmat_hook, pnp_hook, None)
        self.refs = [make_cfg(q)[4]() for q in seqs]

    def imu_feed(self, s, t, a, g):
        assert self.lib.flv_f2f_batch_imu_feed(self.b, s, t, _vp(a), _vp(g)) == 0

    def image_feed(self, ts, imgs0, imgs1):
        t = np.ascontiguousarray(ts, np.float64)
        i0 = np.ascontiguousarray(np.stack(imgs0)); i1 = np.ascontiguousarray(np.stack(imgs1))
        kf = np.zeros(self.S, np.int32); rs = np.zeros(self.S, np.int32)
        rc = self.lib.flv_f2f_batch_image_feed(self.b, _vp(t), _vp(i0), _vp(i1), 0, _vp(kf), _vp(rs))
        assert rc == 0, self.lib.flv_f2f_batch_last_error(self.b)
        return [(bool(kf[s]), bool(rs[s])) for s in range(self.S)]

    def state(self, s):
        return STATE[self.lib.flv_f2f_batch_state(self.b, s)]

    def get_frame(self, s, cap=600):
        T = np.zeros(7); ids = np.zeros(cap, np.int64); pl = np.zeros((cap, 2)); un = np.zeros((cap, 2)); p3 = np.zeros((cap, 3))
        has = np.zeros(cap, np.uint8); inl = np.zeros(cap, np.uint8)
        n = self.lib.flv_f2f_batch_get_frame(self.b, s, _vp(T), _vp(ids), _vp(pl), _vp(un), _vp(p3), _vp(has), _vp(inl), cap)
        return n, T, ids, pl, un, p3, has, inl

    def get_ex(self, s, n):
        cap = max(n, 1)
        p3c = np.zeros((cap, 3)); f2d = np.zeros((cap, 2)); fp = np.zeros((cap, 7)); Tkf = np.zeros(7)
        m = self.lib.flv_f2f_batch_get_frame_ex(self.b, s, _vp(p3c), _vp(f2d), _vp(fp), _vp(Tkf), cap)
        return m, p3c, f2d, fp, Tkf

    def imu_states(self, s):
        st = np.zeros((400, 11))
        return st[:self.lib.flv_f2f_batch_get_imu_states(self.b, s, _vp(st), 400)]

    def bias(self, s):
        ab = np.zeros(3); gb = np.zeros(3)
        return self.lib.flv_f2f_batch_get_imu_bias(self.b, s, _vp(ab), _vp(gb)), ab, gb

    def counts(self, s):
        of = C.c_int(); fi = C.c_int(); pn = C.c_int()
        self.lib.flv_f2f_batch_tracking_counts(self.b, s, C.byref(of), C.byref(fi), C.byref(pn))
        return of.value, fi.value, pn.value

    def close(self):
        self.lib.flv_f2f_batch_destroy(self.b)


def sync_oracle_from_tracker(trk, s, ref, n, T, pl, un, p3):
    """Teacher forcing: overwrite the oracle's continuous state (current frame's landmarks and pose, last-keyframe pose,
    IMU state queue and biases) with the C++ tracker's, bit for bit.  Both pipelines are chaotic in the last bits once the
    IMU pose guess feeds LK's float start positions (a 1e-10 pose difference from the fp64 GPU bundle adjustment flips a
    float rounding, then a FeatureDEM spacing test, then every later landmark id), so a free-running comparison can only
    be statistical; re-seeding after every frame keeps the per-frame comparison exact: each frame's complete transition
    (IMU guess -> LK -> RANSAC -> BA -> reprojection cull -> redetect -> depth innovation -> keyframe rule) is checked from
    identical inputs."""
    m, p3c, f2d, fp, Tkf = trk.get_ex(s, n)
    assert m == n == len(ref.curr.lms)
    raw = lambda t7: OSE3([t7[3], t7[0], t7[1], t7[2]], t7[4:7], normalize=False)
    for i, l in enumerate(ref.curr.lms):
        l.plane = pl[i].copy(); l.undist = un[i].copy(); l.p3d_w = p3[i].copy(); l.p3d_c = p3c[i].copy()
        l.first_2d = f2d[i].copy(); l.first_pose = raw(fp[i])
    ref.curr.T_c_w = raw(T)
    ref.T_kf = raw(Tkf)
    st = trk.imu_states(s)
    assert len(st) == len(ref.vim.states)
    ref.vim.states = [dict(t=float(r[0]), q=r[1:5].copy(), pos=r[5:8].copy(), vel=r[8:11].copy()) for r in st]
    _, ab, gb = trk.bias(s)
    ref.vim.acc_bias, ref.vim.gyro_bias = ab, gb


class LocalMapPair:
    """flv::LocalMap (C handle) next to oracle/localmap_ref.LocalMap, fed with the tracker's keyframes."""

    def __init__(self, window, K):
        from flvis_b200 import capi
        self.capi = capi
        self.ctx = capi.Context(1, 64, 64)
        cl = self.ctx.lib
        cl.flv_localmap_create.restype = C.c_void_p
        cl.flv_localmap_create.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
        cl.flv_localmap_destroy.argtypes = [C.c_void_p]
        cl.flv_localmap_add_keyframe.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 5 + \
            [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(capi.BAStats)]
        self.lm = cl.flv_localmap_create(self.ctx.h, window, *[float(v) for v in K])
        assert self.lm
        self.ref = localmap_ref.LocalMap(window, K)
        self.n_solved = 0
        self.max_dt = 0.0

    def add(self, frame_id, ids, und, p3, T, ref_kf, tol_t=1e-5, tol_lm=1e-4):
        cl = self.ctx.lib
        o = self.ref.frame_callback(ref_kf)
        fid = np.zeros(1, np.int64); oT = np.zeros(7); nlm = np.zeros(1, np.int32); olm = np.zeros(16384, np.int64)
        o3 = np.zeros((16384, 3)); nout = np.zeros(1, np.int32); oout = np.zeros(16384, np.int64)
        st = self.capi.BAStats()
        rc = cl.flv_localmap_add_keyframe(self.lm, int(frame_id), len(ids), _vp(ids), _vp(und), _vp(p3), _vp(T), _vp(fid), _vp(oT),
                                          _vp(nlm), _vp(olm), _vp(o3), 16384, _vp(nout), _vp(oout), 16384, C.byref(st))
        if o is None:
            assert rc == 0, cl.flv_last_error(self.ctx.h)
            return
        assert rc == 1, cl.flv_last_error(self.ctx.h)
        self.n_solved += 1
        assert fid[0] == o["frame_id"] and list(olm[:nlm[0]]) == o["lm_id"]
        assert sorted(oout[:nout[0]]) == sorted(o["outlier_id"])
        self.max_dt = max(self.max_dt, float(np.abs(oT[4:] - o["T_c_w"][4:]).max()))
        assert np.abs(oT[4:] - o["T_c_w"][4:]).max() <= tol_t
        if nlm[0]:
            assert np.abs(o3[:nlm[0]] - o["lm_3d"]).max() <= tol_lm * max(1.0, float(np.abs(o["lm_3d"]).max()))

    def close(self):
        self.ctx.lib.flv_localmap_destroy(self.lm)
        self.ctx.close()


def run_sequence(lib, seq, **kw):
    """One sequence through one flv::F2FTracking handle (see run)."""
    return run(lib, [seq], batch=False, **kw)[0]


def run(lib, seqs, batch=False, tol_pose=1e-6, tol_und=0.0, tol_p3=1e-6, window=None, hooks=True, lockstep=False, free_frames=0, groups=1):
    """Frame-by-frame comparison of len(seqs) sequences (same sensor, same frame count), each against its own oracle, driven
    through N single trackers or ONE batched tracker; returns one summary dict per sequence.
    lockstep: re-seed the oracle from the tracker after every frame (see sync_oracle_from_tracker);
    free_frames: additionally run a free-running oracle over the first `free_frames` frames for the ATE comparison;
    window: chain a local map of that size on every stream's keyframes."""
    trk = BatchTracker(lib, seqs, hooks, groups) if batch else SingleTrackers(lib, seqs, hooks)
    N = len(seqs)
    refs = trk.refs
    frees = [make_cfg(q)[4]() for q in seqs] if free_frames else [None] * N
    out_free = [[] for _ in range(N)]
    lmaps = [LocalMapPair(window, trk.K) if window else None for _ in range(N)]
    outs = [dict(states=[], kf=0, reset=0, guess_used=0, traj=[], traj_ref=[], traj_gt=[], traj_k=[], max_dpose=0.0, frames_tracked=0)
            for _ in range(N)]
    gens = [q.frames() for q in seqs]
    for k in range(seqs[0].n_frames):
        frames = [next(g) for g in gens]
        was_tracking = []
        for s, (t, img0, img1, imu) in enumerate(frames):
            for (ti, acc, gyro) in imu:
                trk.imu_feed(s, float(ti), np.ascontiguousarray(acc), np.ascontiguousarray(gyro))
                refs[s].imu_feed(float(ti), acc, gyro)
                if frees[s] is not None and k < free_frames:
                    frees[s].imu_feed(float(ti), acc, gyro)
            if frees[s] is not None and k < free_frames:
                frees[s].image_feed(float(t), img0, img1)
                if frees[s].state == "Tracking":
                    out_free[s].append((k, _cam_centre(frees[s].curr.T_c_w.to7())))
            was_tracking.append(refs[s].state == "Tracking")
            if was_tracking[s] and refs[s].has_imu and refs[s].vim.corr_frame_state(t) is not None:
                outs[s]["guess_used"] += 1
        flags = trk.image_feed([f[0] for f in frames], [f[1] for f in frames], [f[2] for f in frames])
        for s, (t, img0, img1, imu) in enumerate(frames):
            ref, out, seq = refs[s], outs[s], seqs[s]
            kf, rs = flags[s]
            rkf, rrs = ref.image_feed(float(t), img0, img1)
            state = trk.state(s)
            assert state == ref.state and kf == rkf and rs == rrs, (k, s, state, ref.state, kf, rkf, rs, rrs)
            out["states"].append(state); out["kf"] += int(rkf); out["reset"] += int(rrs)
            n, T, ids, pl, un, p3, has, inl = trk.get_frame(s)
            cur = ref.curr
            assert n == len(cur.lms), (k, s, n, len(cur.lms))
            assert list(ids[:n]) == [l.lm_id for l in cur.lms], (k, s)                 # landmark ids + order: bit-exact
            assert list(inl[:n].astype(bool)) == [bool(l.inlier) for l in cur.lms], (k, s)
            assert list(has[:n].astype(bool)) == [bool(l.has_3d) for l in cur.lms], (k, s)
            if n:
                assert np.array_equal(pl[:n], np.array([l.plane for l in cur.lms])), (k, s)      # LK pixel positions: bit-exact
                assert np.abs(un[:n] - np.array([l.undist for l in cur.lms])).max() <= tol_und, (k, s)
                r3 = np.array([l.p3d_w for l in cur.lms])
                assert np.abs(p3[:n] - r3).max() <= tol_p3 * max(1.0, float(np.abs(r3).max())), (k, s)
            rT = cur.T_c_w.to7()
            dpose = max(float(np.abs(T[4:] - rT[4:]).max()), 2 * float(np.arccos(min(1.0, abs(float(np.dot(T[:4], rT[:4])))))))
            out["max_dpose"] = max(out["max_dpose"], dpose)
            assert dpose <= tol_pose, (k, s, dpose)
            if state == "Tracking":
                out["frames_tracked"] += 1
                out["traj_k"].append(k)
                out["traj"].append(_cam_centre(T)); out["traj_ref"].append(_cam_centre(rT))
                out["traj_gt"].append(np.asarray(seq.T_w_c0(t).t, float))
                if was_tracking[s]:
                    assert trk.counts(s) == ref.counts, (k, s, trk.counts(s), ref.counts)
            if lmaps[s] is not None and rkf:
                sel = (has[:n] == 1) & (inl[:n] == 1)                                   # CameraFrame::getKeyFrameInf
                kids = np.ascontiguousarray(ids[:n][sel]); kuv = np.ascontiguousarray(un[:n][sel]); k3 = np.ascontiguousarray(p3[:n][sel])
                rsel = [l for l in cur.lms if l.has_3d and l.inlier]
                okf = {"frame_id": cur.frame_id, "lm_id": [l.lm_id for l in rsel], "lm_2d": np.array([l.undist for l in rsel]),
                       "lm_3d": np.array([l.p3d_w for l in rsel]), "T_c_w": rT}
                lmaps[s].add(cur.frame_id, kids, kuv, k3, T, okf)
            if lockstep:
                sync_oracle_from_tracker(trk, s, ref, n, T, pl, un, p3)
    for s in range(N):
        out, ref = outs[s], refs[s]
        if frees[s] is not None:
            # free-running reference path against the GPU path: absolute trajectory error over the common tracked frames
            kk = {k: c for k, c in out_free[s]}
            pairs = [(c, kk[k]) for k, c in zip(out["traj_k"], out["traj"]) if k in kk]
            a = np.array([p[0] for p in pairs]); b = np.array([p[1] for p in pairs])
            out["free_frames_compared"] = len(pairs)
            out["ate_vs_free_ref"] = float(np.sqrt(np.mean(np.sum((a - b) ** 2, axis=1))))
            out["free_path"] = float(np.sum(np.linalg.norm(np.diff(b, axis=0), axis=1)))
        out["has_imu"], out["acc_bias"], out["gyro_bias"] = trk.bias(s)
        out["ref_acc_bias"], out["ref_gyro_bias"] = ref.vim.acc_bias.copy(), ref.vim.gyro_bias.copy()
        out["final_state"] = ref.state
        if lmaps[s] is not None:
            out["n_solved"] = lmaps[s].n_solved; out["localmap_max_dt"] = lmaps[s].max_dt
            lmaps[s].close()
        traj, traj_ref, gt = np.array(out["traj"]), np.array(out["traj_ref"]), np.array(out["traj_gt"])
        out["ate_vs_ref"] = float(np.sqrt(np.mean(np.sum((traj - traj_ref) ** 2, axis=1))))
        out["path"] = float(np.sum(np.linalg.norm(np.diff(traj_ref, axis=0), axis=1)))
        # ATE of each path against the synthetic ground truth after removing the constant world-frame offset (the tracker's
        # world starts at its first pose): BASELINE's "ATE within 1 % of the reference"
        d = traj - gt; dr = traj_ref - gt
        out["ate_gt"] = float(np.sqrt(np.mean(np.sum((d - d.mean(0)) ** 2, axis=1))))
        out["ate_gt_ref"] = float(np.sqrt(np.mean(np.sum((dr - dr.mean(0)) ** 2, axis=1))))
    trk.close()
    return outs
