"""ctypes binding of the batched tracker (flv_f2f_batch) and the local-map worker (flv_localmap_batch):
include/flvis_b200_host.h.  Plumbing for bench.py, the multi-GPU driver and the tests; nothing here computes."""
import ctypes as C

import numpy as np

from . import capi

STATE = {0: "UnInit", 1: "Tracking", 2: "TrackingFail"}
STAGES = ["ingest+pyramids", "lk_f2f", "keep+fmat_ransac", "pnp_ransac", "pose_only_ba", "reprojection_cull", "redetect",
          "lk_left_right", "depth_innovation+finish"]


class F2FConfig(C.Structure):          # flv_f2f_config
    _fields_ = [("cam_type", C.c_int), ("img_w", C.c_int), ("img_h", C.c_int), ("cam0", C.c_double * 4),
                ("cam1", C.c_double * 4), ("depth_scale", C.c_double), ("P0", C.c_double * 12), ("P1", C.c_double * 12),
                ("T_cam1_cam0", C.c_double * 7), ("T_i_c0", C.c_double * 7), ("feature_para", C.c_double * 6),
                ("vi_para", C.c_double * 6), ("dc_para", C.c_double * 3), ("skip_first_n_imgs", C.c_int)]


class F2FStereoConfig(C.Structure):    # flv_f2f_stereo_config (DepthCamera::setSteroCamInfo's arguments)
    _fields_ = [("cam_type", C.c_int), ("img_w", C.c_int), ("img_h", C.c_int),
                ("K0", C.c_double * 9), ("D0", C.c_double * 14), ("R0", C.c_double * 9), ("P0", C.c_double * 12),
                ("K1", C.c_double * 9), ("D1", C.c_double * 14), ("R1", C.c_double * 9), ("P1", C.c_double * 12),
                ("T_c0_c1", C.c_double * 7), ("T_i_c0", C.c_double * 7), ("feature_para", C.c_double * 6),
                ("vi_para", C.c_double * 6), ("dc_para", C.c_double * 3), ("skip_first_n_imgs", C.c_int), ("need_equal_hist", C.c_int)]


def stereo_config_for(seq):
    """flv_f2f_stereo_config for a STEREO_UNRECT synthdata sequence: the matrices vo_tracking.cpp:222-262 passes to
    DepthCamera::setSteroCamInfo (cv::stereoRectify output included)."""
    _, _, equalize, _, m = config_for(seq)
    c = seq.cfg
    d = lambda vals, n: (C.c_double * n)(*([float(v) for v in np.asarray(vals).ravel()] + [0.0] * (n - np.asarray(vals).size)))
    return F2FStereoConfig(2, c["w"], c["h"], d(m["K0"], 9), d(m["D0"], 14), d(m["R0"], 9), d(m["P0"], 12), d(m["K1"], 9), d(m["D1"], 14),
                           d(m["R1"], 9), d(m["P1"], 12), d(seq.T_c0_c1.to7(), 7), d(seq.T_i_c0.to7(), 7), d(c["feature_para"], 6),
                           d(c["vi_para"], 6), d(c["dc_para"], 3), 0, 1 if equalize else 0)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _bind(lib):
    if getattr(lib, "_flv_batch_bound", False):
        return lib
    vp = C.c_void_p
    lib.flv_f2f_batch_create.restype = vp
    lib.flv_f2f_batch_create.argtypes = [C.POINTER(F2FConfig), C.c_int, C.c_int]
    lib.flv_f2f_batch_create_grouped.restype = vp
    lib.flv_f2f_batch_create_grouped.argtypes = [C.POINTER(F2FConfig), C.c_int, C.c_int, C.c_int]
    lib.flv_f2f_batch_groups.argtypes = [vp]
    lib.flv_f2f_batch_destroy.argtypes = [vp]
    lib.flv_f2f_batch_last_error.restype = C.c_char_p
    lib.flv_f2f_batch_last_error.argtypes = [vp]
    lib.flv_f2f_batch_context.restype = vp
    lib.flv_f2f_batch_context.argtypes = [vp]
    lib.flv_f2f_batch_set_lens.argtypes = [vp, C.c_int, vp, vp, vp]
    lib.flv_f2f_batch_set_equalize_hist.argtypes = [vp, C.c_int]
    lib.flv_f2f_batch_imu_feed_many.argtypes = [vp, C.c_int, vp, vp, vp, vp]
    lib.flv_f2f_batch_image_feed.argtypes = [vp, vp, vp, vp, C.c_int, vp, vp]
    lib.flv_f2f_batch_state.argtypes = [vp, C.c_int]
    lib.flv_f2f_batch_frame_async.argtypes = [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]
    lib.flv_f2f_batch_sync.argtypes = [vp, vp, vp]
    lib.flv_f2f_batch_get_frame.argtypes = [vp, C.c_int] + [vp] * 7 + [C.c_int]
    lib.flv_f2f_batch_launch_count.restype = C.c_longlong
    lib.flv_f2f_batch_launch_count.argtypes = [vp]
    lib.flv_f2f_batch_attach_localmap.argtypes = [vp, vp]
    lib.flv_f2f_batch_set_readback.argtypes = [vp, C.c_int]
    lib.flv_f2f_batch_set_result_log.argtypes = [vp, vp, C.c_int]
    lib.flv_f2f_batch_set_profile.argtypes = [vp, C.c_int]
    lib.flv_f2f_batch_get_profile.argtypes = [vp, vp, vp]
    lib.flv_f2f_batch_get_host_profile.argtypes = [vp, vp]
    lib.flv_f2f_batch_tracking_counts.argtypes = [vp, C.c_int] + [C.POINTER(C.c_int)] * 3
    lib.flv_localmap_batch_create.restype = vp
    lib.flv_localmap_batch_create.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
    lib.flv_localmap_batch_destroy.argtypes = [vp]
    lib.flv_localmap_batch_wait.argtypes = [vp]
    lib.flv_localmap_batch_stats.argtypes = [vp] * 5
    lib.flv_localmap_batch_problem_totals.argtypes = [vp, vp]
    lib.flv_localmap_batch_last_error.restype = C.c_char_p
    lib.flv_localmap_batch_last_error.argtypes = [vp]
    lib.flv_set_stream.argtypes = [vp, vp]
    lib._flv_batch_bound = True
    return lib


class LocalMapBatch:
    def __init__(self, n_streams, window, K, device=0, lib=None):
        self.lib = _bind(lib or capi.load_library())
        self.h = self.lib.flv_localmap_batch_create(device, n_streams, window, *[float(v) for v in K])
        if not self.h:
            raise capi.FlvError("flv_localmap_batch_create failed (window must be 3..100)")

    def wait(self):
        if self.lib.flv_localmap_batch_wait(self.h) != 0:
            raise capi.FlvError(self.lib.flv_localmap_batch_last_error(self.h).decode())

    def stats(self):
        nk = C.c_longlong(); ns = C.c_longlong(); nl = C.c_longlong(); ms = (C.c_double * 2)()
        self.lib.flv_localmap_batch_stats(self.h, C.byref(nk), C.byref(ns), C.byref(nl), ms)
        tot = (C.c_double * 4)()
        self.lib.flv_localmap_batch_problem_totals(self.h, tot)
        return dict(keyframes=nk.value, solves=ns.value, launches=nl.value, solve_ms=ms[0], host_ms=ms[1],
                    edges=tot[0], landmarks=tot[1], poses=tot[2], lm_iterations=tot[3])

    def close(self):
        if self.h:
            self.lib.flv_localmap_batch_destroy(self.h)
            self.h = None


class BatchTracker:
    """S sequences of one sensor model.  cfg: F2FConfig; lenses: [(K4, D14, R9), (K4, D14, R9)] for STEREO_UNRECT."""

    def __init__(self, cfg, n_streams, device=0, lenses=None, equalize=False, lib=None, groups=1):
        self.lib = _bind(lib or capi.load_library())
        self.S = n_streams
        self.h = self.lib.flv_f2f_batch_create_grouped(C.byref(cfg), n_streams, device, groups)
        self.groups = self.lib.flv_f2f_batch_groups(self.h) if self.h else 0
        err = self.lib.flv_f2f_batch_last_error(self.h) if self.h else b"allocation failed"
        if not self.h or err:
            raise capi.FlvError(f"flv_f2f_batch_create: {err.decode()}")
        if lenses:
            for cam, (k4, d14, r9) in enumerate(lenses):
                self._chk(self.lib.flv_f2f_batch_set_lens(self.h, cam, _p(np.ascontiguousarray(k4, np.float64)),
                                                          _p(np.ascontiguousarray(d14, np.float64)), _p(np.ascontiguousarray(r9, np.float64))))
        if equalize:
            self._chk(self.lib.flv_f2f_batch_set_equalize_hist(self.h, 1))
        self.kf = np.zeros(n_streams, np.int32)
        self.rs = np.zeros(n_streams, np.int32)

    def _chk(self, rc):
        if rc != 0:
            raise capi.FlvError(f"flv_f2f_batch ({rc}): {self.lib.flv_f2f_batch_last_error(self.h).decode()}")

    def set_stream(self, cuda_stream_ptr):
        """Single-group batches only: run on the caller's CUDA stream (grouped batches own one stream per group)."""
        if self.groups > 1:
            return
        self.lib.flv_set_stream(self.lib.flv_f2f_batch_context(self.h), C.c_void_p(cuda_stream_ptr))

    def set_readback(self, full):
        self._chk(self.lib.flv_f2f_batch_set_readback(self.h, 1 if full else 0))

    def set_result_log(self, ptr, block_frames):
        self._chk(self.lib.flv_f2f_batch_set_result_log(self.h, C.c_void_p(ptr) if ptr else None, block_frames))

    def attach_localmap(self, lm):
        self._chk(self.lib.flv_f2f_batch_attach_localmap(self.h, lm.h if lm else None))

    def imu_feed_many(self, streams, t, acc, gyro):
        n = len(streams)
        if n:
            self._chk(self.lib.flv_f2f_batch_imu_feed_many(self.h, n, _p(streams), _p(t), _p(acc), _p(gyro)))

    def image_feed(self, t, img0_ptr, img1_ptr, device_mem):
        """t: float64[S]; img pointers: raw addresses of S images back to back (host or device memory)."""
        self._chk(self.lib.flv_f2f_batch_image_feed(self.h, _p(t), C.c_void_p(img0_ptr), C.c_void_p(img1_ptr), 1 if device_mem else 0,
                                                    _p(self.kf), _p(self.rs)))
        return self.kf, self.rs

    def frame_async(self, t, img0_ptr, img1_ptr, device_mem, imu=None):
        """Pipelined frame: imu = (streams int32[n], t float64[n], acc float64[n,3], gyro float64[n,3]) or None."""
        if imu is None or len(imu[0]) == 0:
            self._chk(self.lib.flv_f2f_batch_frame_async(self.h, _p(t), C.c_void_p(img0_ptr), C.c_void_p(img1_ptr), 1 if device_mem else 0,
                                                         0, None, None, None, None))
        else:
            st, ti, acc, gyro = imu
            self._chk(self.lib.flv_f2f_batch_frame_async(self.h, _p(t), C.c_void_p(img0_ptr), C.c_void_p(img1_ptr), 1 if device_mem else 0,
                                                         len(st), _p(st), _p(ti), _p(acc), _p(gyro)))

    def sync(self):
        self._chk(self.lib.flv_f2f_batch_sync(self.h, _p(self.kf), _p(self.rs)))
        return self.kf, self.rs

    def state(self, s):
        return STATE[self.lib.flv_f2f_batch_state(self.h, s)]

    def pose(self, s):
        T = np.zeros(7)
        self.lib.flv_f2f_batch_get_frame(self.h, s, _p(T), None, None, None, None, None, None, 0)
        return T

    def n_landmarks(self, s):
        return self.lib.flv_f2f_batch_get_frame(self.h, s, None, None, None, None, None, None, None, 1 << 20)

    def launches(self):
        return int(self.lib.flv_f2f_batch_launch_count(self.h))

    def set_profile(self, on):
        self._chk(self.lib.flv_f2f_batch_set_profile(self.h, 1 if on else 0))

    def profile(self):
        ms = np.zeros(9); n = C.c_longlong()
        self.lib.flv_f2f_batch_get_profile(self.h, _p(ms), C.byref(n))
        return ms, n.value

    def host_profile(self):
        ms = np.zeros(4)
        self.lib.flv_f2f_batch_get_host_profile(self.h, _p(ms))
        return ms

    def close(self):
        if self.h:
            self.lib.flv_f2f_batch_destroy(self.h)
            self.h = None


def config_for(seq):
    """F2FConfig (+ lens models, equalize flag, rectified K) for a synthdata sequence, the way TrackingNodeletClass::onInit
    fills DepthCamera for that sensor (vo_tracking.cpp:140-306); STEREO_UNRECT needs cv2.stereoRectify (the reference calls
    cv::stereoRectify there -- node initialisation, outside the hot path)."""
    c = seq.cfg
    d = lambda vals, n: (C.c_double * n)(*[float(v) for v in vals])
    Ti = seq.T_i_c0.to7()
    lenses, equalize = None, False
    if seq.cam_type == "depth":
        K = c["K0"]
        cfg = F2FConfig(0, c["w"], c["h"], d(K, 4), d(K, 4), c["depth_factor"], (C.c_double * 12)(), (C.c_double * 12)(),
                        d([0, 0, 0, 1, 0, 0, 0], 7), d(Ti, 7), d(c["feature_para"], 6), d(c["vi_para"], 6), d(c["dc_para"], 3), c["skip"])
    elif seq.cam_type == "stereo_unrect":
        import cv2
        from synthdata.se3 import q2R
        T10 = seq.T_c0_c1.inverse()
        K0 = np.array([[c["K0"][0], 0, c["K0"][2]], [0, c["K0"][1], c["K0"][3]], [0, 0, 1.0]])
        K1 = np.array([[c["K1"][0], 0, c["K1"][2]], [0, c["K1"][1], c["K1"][3]], [0, 0, 1.0]])
        D0, D1 = np.array(c["D0"]), np.array(c["D1"])
        size = (c["w"], c["h"])
        R0, R1, P0, P1, _, _, _ = cv2.stereoRectify(K0, D0, K1, D1, size, q2R(T10.q), T10.t.reshape(3, 1), flags=cv2.CALIB_ZERO_DISPARITY,
                                                    alpha=0, newImageSize=size)
        K = (P0[0, 0], P0[1, 1], P0[0, 2], P0[1, 2]); Kr1 = (P1[0, 0], P1[1, 1], P1[0, 2], P1[1, 2])
        cfg = F2FConfig(2, c["w"], c["h"], d(K, 4), d(Kr1, 4), 1000.0, d(P0.ravel(), 12), d(P1.ravel(), 12), d(T10.to7(), 7), d(Ti, 7),
                        d(c["feature_para"], 6), d(c["vi_para"], 6), d(c["dc_para"], 3), 0)
        d14 = lambda D: np.concatenate([D, np.zeros(14 - len(D))])
        lenses = [(np.array([Kr[0, 0], Kr[1, 1], Kr[0, 2], Kr[1, 2]]), d14(D), np.ascontiguousarray(Rr).ravel().copy())
                  for Kr, D, Rr in ((K0, D0, R0), (K1, D1, R1))]
        equalize = True
        return cfg, lenses, equalize, K, dict(K0=K0, D0=D0, R0=R0, P0=P0, K1=K1, D1=D1, R1=R1, P1=P1)
    else:
        K = c["K0"]
        P0 = np.array([[K[0], 0, K[2], 0], [0, K[1], K[3], 0], [0, 0, 1, 0.0]]); P1 = P0.copy(); P1[0, 3] = -c["bf"]
        T10 = seq.T_c0_c1.inverse()
        cfg = F2FConfig(1, c["w"], c["h"], d(K, 4), d(K, 4), 1000.0, d(P0.ravel(), 12), d(P1.ravel(), 12), d(T10.to7(), 7), d(Ti, 7),
                        d(c["feature_para"], 6), d(c["vi_para"], 6), d(c["dc_para"], 3), 0)
    return cfg, lenses, equalize, K, None
