"""ctypes binding of include/flvis_b200.h (tests / bench / multi-GPU driver only).

Fails loudly when libflvis_b200.so is missing: there is no CPU or PyTorch fallback.
"""
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libflvis_b200.so")

FLV_OK = 0
MEM_HOST, MEM_DEVICE = 0, 1
NUM_SLOTS = 4

SYMBOLS = [
    "flv_create", "flv_destroy", "flv_set_stream", "flv_sync", "flv_last_error", "flv_version",
    "flv_launch_count", "flv_level_info", "flv_num_levels", "flv_upload_images", "flv_build_pyramid",
    "flv_download_level", "flv_lk_track", "flv_select_tracked", "flv_gftt", "flv_download_eig", "flv_gftt_capacity",
    "flv_feature_detect", "flv_feature_redetect", "flv_ba_reserve", "flv_ba_optimize", "flv_gftt_keep_response",
    "flv_ba_profile", "flv_set_ba_stream", "flv_set_ba_cluster", "flv_feature_prepare", "flv_set_equalize_hist", "flv_fundamental_ransac", "flv_pnp_ransac", "flv_upload_color_images", "flv_depth_innovation", "flv_reprojection_inliers",
    "flv_ba_trace", "flv_ba_debug_edges", "flv_ba_big_emulate_host", "flv_get_flags",
]


class LKParams(C.Structure):
    _fields_ = [("win", C.c_int), ("max_level", C.c_int), ("max_iter", C.c_int), ("eps", C.c_double),
                ("min_eig_threshold", C.c_double)]


class RansacParams(C.Structure):
    _fields_ = [("threshold_px", C.c_double), ("confidence", C.c_double), ("max_iterations", C.c_int)]


class FeatureParams(C.Structure):
    _fields_ = [("max_region_feature_num", C.c_int), ("min_region_feature_num", C.c_int),
                ("boundary_dis", C.c_int), ("gftt_num", C.c_int), ("gftt_ql", C.c_double),
                ("gftt_dis", C.c_int)]


class Camera(C.Structure):
    _fields_ = [("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double),
                ("P0", C.c_double * 12), ("P1", C.c_double * 12), ("cam_type", C.c_int), ("depth_scale", C.c_double)]


class DepthParams(C.Structure):
    _fields_ = [("iir_ratio", C.c_float), ("range", C.c_float), ("dummy_depth", C.c_int)]


class BAProblem(C.Structure):
    _fields_ = [("n_poses", C.c_int), ("n_landmarks", C.c_int), ("n_edges", C.c_int),
                ("fixed_pose", C.c_int), ("fix_landmarks", C.c_int),
                ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


class BAParams(C.Structure):
    _fields_ = [("iters1", C.c_int), ("iters2", C.c_int), ("huber_delta", C.c_double),
                ("cull_chi2", C.c_double), ("min_edges_after_cull", C.c_int), ("ws_slot0", C.c_int)]


class BAStats(C.Structure):
    _fields_ = [("iterations_run", C.c_int), ("n_culled", C.c_int), ("ok", C.c_int), ("reserved", C.c_int),
                ("chi2_initial", C.c_double), ("chi2_after1", C.c_double), ("chi2_final", C.c_double),
                ("lambda_final", C.c_double)]


def load_library(path=LIB_PATH):
    if not os.path.exists(path):
        raise ImportError(
            f"{path} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C flvis_b200/csrc`.  flvis_b200 has no CPU fallback.")
    lib = C.CDLL(path)
    lib.flv_last_error.restype = C.c_char_p
    lib.flv_version.restype = C.c_char_p
    lib.flv_launch_count.restype = C.c_longlong
    lib.flv_destroy.restype = None
    vp = C.c_void_p
    lib.flv_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int]
    lib.flv_destroy.argtypes = [vp]
    lib.flv_set_stream.argtypes = [vp, vp]
    lib.flv_sync.argtypes = [vp]
    lib.flv_last_error.argtypes = [vp]
    lib.flv_launch_count.argtypes = [vp]
    lib.flv_num_levels.argtypes = [vp]
    lib.flv_gftt_capacity.argtypes = [vp]
    lib.flv_gftt_keep_response.argtypes = [vp, C.c_int]
    lib.flv_ba_profile.argtypes = [vp, C.c_int, vp]
    lib.flv_ba_trace.argtypes = [vp, C.c_int, vp, C.c_int, C.c_int]
    lib.flv_ba_debug_edges.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.flv_ba_reserve.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.flv_set_ba_stream.argtypes = [vp, vp, C.c_int]
    lib.flv_set_ba_cluster.argtypes = [vp, C.c_int, C.c_int]
    lib.flv_level_info.argtypes = [vp, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                   C.POINTER(C.c_size_t)]
    lib.flv_upload_images.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_int]
    lib.flv_build_pyramid.argtypes = [vp, C.c_int, C.c_int]
    lib.flv_download_level.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, C.c_int]
    lib.flv_lk_track.argtypes = [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, vp, vp, vp,
                                 C.POINTER(LKParams), C.c_int]
    lib.flv_select_tracked.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, vp, vp]
    lib.flv_gftt.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, vp, vp, C.c_int, C.c_int]
    lib.flv_download_eig.argtypes = [vp, C.c_int, vp, C.c_int]
    lib.flv_set_equalize_hist.argtypes = [vp, C.c_int]
    lib.flv_upload_color_images.argtypes = [vp, C.c_int, C.c_int, vp, C.c_size_t, C.c_size_t, C.c_int, C.c_int, C.c_int]
    lib.flv_fundamental_ransac.argtypes = [vp, C.c_int, vp, vp, vp, C.POINTER(RansacParams), vp, vp, vp, C.c_int]
    lib.flv_pnp_ransac.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.POINTER(RansacParams), vp, vp, vp, C.c_int]
    lib.flv_feature_prepare.argtypes = [vp, C.c_int, C.c_int, C.POINTER(FeatureParams), C.c_int]
    lib.flv_feature_detect.argtypes = [vp, C.c_int, C.c_int, C.POINTER(FeatureParams), vp, vp, C.c_int]
    lib.flv_feature_redetect.argtypes = [vp, C.c_int, C.c_int, C.POINTER(FeatureParams), vp, vp, vp, vp, C.c_int]
    lib.flv_depth_innovation.argtypes = [vp, C.c_int, vp, C.POINTER(Camera), C.POINTER(DepthParams)] + [vp] * 13 + [C.c_int]
    lib.flv_reprojection_inliers.argtypes = [vp, C.c_int, vp, C.POINTER(Camera), vp, vp, vp, C.c_double, vp, vp, C.c_int]
    lib.flv_ba_reserve.argtypes = [vp, C.c_int, C.c_int, C.c_int]
    lib.flv_ba_optimize.argtypes = [vp, C.c_int, C.POINTER(BAProblem), C.POINTER(BAParams), vp, vp, vp, vp, vp,
                                    vp, C.POINTER(BAStats), C.c_int]
    return lib


class FlvError(RuntimeError):
    pass


def _ptr(a):
    """numpy array -> void* (host); int -> device pointer passthrough."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    assert a.flags["C_CONTIGUOUS"], "array must be C-contiguous"
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Thin OO wrapper over the C ABI.  Host-memory calls take/return numpy arrays; *_dev calls take
    raw device pointers (e.g. torch.Tensor.data_ptr()) and only enqueue work."""

    def __init__(self, max_streams, width, height, max_pts=512, device=0, lib=None):
        self.lib = lib or load_library()
        self.h = C.c_void_p()
        self.S, self.w, self.hh, self.max_pts = max_streams, width, height, max_pts
        rc = self.lib.flv_create(C.byref(self.h), device, max_streams, width, height, max_pts)
        if rc != FLV_OK:
            msg = self.lib.flv_last_error(self.h).decode() if self.h else "allocation failed"
            raise FlvError(f"flv_create failed ({rc}): {msg}")

    def close(self):
        if self.h:
            self.lib.flv_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _chk(self, rc):
        if rc != FLV_OK:
            raise FlvError(f"flvis_b200 error {rc}: {self.lib.flv_last_error(self.h).decode()}")

    # -- plumbing
    def set_stream(self, cuda_stream_ptr):
        self._chk(self.lib.flv_set_stream(self.h, C.c_void_p(cuda_stream_ptr)))

    def sync(self):
        self._chk(self.lib.flv_sync(self.h))

    @property
    def launches(self):
        return int(self.lib.flv_launch_count(self.h))

    @property
    def num_levels(self):
        return int(self.lib.flv_num_levels(self.h))

    def level_info(self, level):
        w, h, p, o = C.c_int(), C.c_int(), C.c_int(), C.c_size_t()
        self._chk(self.lib.flv_level_info(self.h, level, C.byref(w), C.byref(h), C.byref(p), C.byref(o)))
        return w.value, h.value, p.value, o.value

    # -- images / pyramid
    def upload(self, slot, imgs):
        """imgs: (S,h,w) u8 numpy (host)."""
        imgs = np.ascontiguousarray(imgs, np.uint8)
        if imgs.ndim == 2:
            imgs = imgs[None]
        n = imgs.shape[0]
        assert imgs.shape[1:] == (self.hh, self.w)
        self._chk(self.lib.flv_upload_images(self.h, slot, n, _ptr(imgs), self.w, self.w * self.hh, MEM_HOST))
        self.sync()

    def upload_dev(self, slot, n_streams, dev_ptr, row_stride=None, img_stride=None):
        row_stride = row_stride or self.w
        img_stride = img_stride or row_stride * self.hh
        self._chk(self.lib.flv_upload_images(self.h, slot, n_streams, C.c_void_p(dev_ptr), row_stride,
                                             img_stride, MEM_DEVICE))

    def upload_host_async(self, slot, n_streams, host_ptr, row_stride=None, img_stride=None):
        """Pinned host pointer; the copy is enqueued on the ctx stream without synchronising."""
        row_stride = row_stride or self.w
        img_stride = img_stride or row_stride * self.hh
        self._chk(self.lib.flv_upload_images(self.h, slot, n_streams, C.c_void_p(host_ptr), row_stride,
                                             img_stride, MEM_HOST))

    def build_pyramid(self, slot, n_streams):
        self._chk(self.lib.flv_build_pyramid(self.h, slot, n_streams))

    def download_level(self, slot, stream, level):
        w, h, _, _ = self.level_info(level)
        out = np.empty((h, w), np.uint8)
        self._chk(self.lib.flv_download_level(self.h, slot, stream, level, _ptr(out), MEM_HOST))
        return out

    # -- LK
    def lk_track(self, src_slot, dst_slot, prev_xy, init_xy, n_pts=None, max_level=10, max_iter=30,
                 eps=1e-3, min_eig=1e-4, win=31):
        """prev_xy/init_xy: (S,max_pts,2) f32 host arrays (or (n,2) for a single stream)."""
        prev_xy = np.asarray(prev_xy, np.float32)
        init_xy = np.asarray(init_xy, np.float32)
        single = prev_xy.ndim == 2
        if single:
            n = len(prev_xy)
            p = np.zeros((1, self.max_pts, 2), np.float32); p[0, :n] = prev_xy
            q = np.zeros((1, self.max_pts, 2), np.float32); q[0, :n] = init_xy
            prev_xy, init_xy, n_pts = p, q, np.array([n], np.int32)
        S = prev_xy.shape[0]
        prev_xy = np.ascontiguousarray(prev_xy, np.float32)
        init_xy = np.ascontiguousarray(init_xy, np.float32)
        n_pts = np.ascontiguousarray(n_pts, np.int32)
        nxt = np.empty_like(init_xy)
        st = np.empty((S, self.max_pts), np.uint8)
        er = np.empty((S, self.max_pts), np.float32)
        prm = LKParams(win, max_level, max_iter, eps, min_eig)
        self._chk(self.lib.flv_lk_track(self.h, src_slot, dst_slot, S, _ptr(n_pts), _ptr(prev_xy), _ptr(init_xy),
                                        _ptr(nxt), _ptr(st), _ptr(er), C.byref(prm), MEM_HOST))
        if single:
            n = int(n_pts[0])
            return nxt[0, :n], st[0, :n], er[0, :n]
        return nxt, st, er

    def lk_track_dev(self, src_slot, dst_slot, n_streams, d_npts, d_prev, d_init, d_next, d_status, d_err,
                     max_level=10, max_iter=30, eps=1e-3, min_eig=1e-4, win=31):
        prm = LKParams(win, max_level, max_iter, eps, min_eig)
        self._chk(self.lib.flv_lk_track(self.h, src_slot, dst_slot, n_streams, C.c_void_p(d_npts),
                                        C.c_void_p(d_prev), C.c_void_p(d_init), C.c_void_p(d_next),
                                        C.c_void_p(d_status), C.c_void_p(d_err), C.byref(prm), MEM_DEVICE))

    def select_tracked_dev(self, n_streams, d_npts, d_prev, d_next, d_status, d_keep, d_out, d_out64):
        self._chk(self.lib.flv_select_tracked(self.h, n_streams, C.c_void_p(d_npts), C.c_void_p(d_prev),
                                              C.c_void_p(d_next), C.c_void_p(d_status), C.c_void_p(d_keep),
                                              C.c_void_p(d_out), C.c_void_p(d_out64)))

    # -- GFTT / FeatureDEM
    def gftt(self, slot, n_streams, max_corners, quality, min_distance):
        xy = np.zeros((n_streams, max_corners, 2), np.float32)
        n = np.zeros(n_streams, np.int32)
        self._chk(self.lib.flv_gftt(self.h, slot, n_streams, max_corners, quality, min_distance, _ptr(xy), _ptr(n),
                                    max_corners, MEM_HOST))
        return [xy[s, :n[s]].copy() for s in range(n_streams)]

    def gftt_dev(self, slot, n_streams, max_corners, quality, min_distance, d_xy, d_n, stride_pts):
        self._chk(self.lib.flv_gftt(self.h, slot, n_streams, max_corners, quality, min_distance, C.c_void_p(d_xy),
                                    C.c_void_p(d_n), stride_pts, MEM_DEVICE))

    def keep_response(self, enable=True):
        self._chk(self.lib.flv_gftt_keep_response(self.h, 1 if enable else 0))

    def set_ba_stream(self, cuda_stream_ptr, enable=True):
        self._chk(self.lib.flv_set_ba_stream(self.h, C.c_void_p(cuda_stream_ptr), 1 if enable else 0))

    def set_ba_cluster(self, host_mode=4, device_mode=1):
        self._chk(self.lib.flv_set_ba_cluster(self.h, host_mode, device_mode))

    def ba_profile(self, stream):
        out = np.zeros(16, np.int64)
        self._chk(self.lib.flv_ba_profile(self.h, stream, _ptr(out)))
        return out

    def ba_trace(self, stream, iterations_run):
        """Per-iteration LM trace of the last solve: rows of (chi2, lambda, rho, trials)."""
        out = np.zeros((32, 4))
        n = self.lib.flv_ba_trace(self.h, stream, _ptr(out), 32, int(iterations_run))
        if n < 0:
            self._chk(n)
        return out[:n]

    def ba_debug_edges(self, poses7, pts3, uv2, K4):
        """(r, A, B) exactly as ba_kernel evaluates them for n independent edges."""
        poses7 = np.ascontiguousarray(poses7, np.float64).reshape(-1, 7); n = len(poses7)
        pts3 = np.ascontiguousarray(pts3, np.float64).reshape(n, 3); uv2 = np.ascontiguousarray(uv2, np.float64).reshape(n, 2)
        K4 = np.ascontiguousarray(K4, np.float64)
        r = np.zeros((n, 2)); A = np.zeros((n, 2, 3)); B = np.zeros((n, 2, 6))
        self._chk(self.lib.flv_ba_debug_edges(self.h, n, _ptr(poses7), _ptr(pts3), _ptr(uv2), _ptr(K4), _ptr(r), _ptr(A), _ptr(B)))
        return r, A, B

    def download_eig(self, stream):
        out = np.empty((self.hh, self.w), np.float32)
        self._chk(self.lib.flv_download_eig(self.h, stream, _ptr(out), MEM_HOST))
        return out

    def feature_detect_dev(self, slot, n_streams, fp, d_xy, d_n):
        self._chk(self.lib.flv_feature_detect(self.h, slot, n_streams, C.byref(fp), C.c_void_p(d_xy), C.c_void_p(d_n),
                                              MEM_DEVICE))

    def feature_detect(self, slot, n_streams, fp):
        xy = np.zeros((n_streams, self.max_pts, 2), np.float32)
        n = np.zeros(n_streams, np.int32)
        self._chk(self.lib.flv_feature_detect(self.h, slot, n_streams, C.byref(fp), _ptr(xy), _ptr(n), MEM_HOST))
        return [xy[s, :n[s]].copy() for s in range(n_streams)]

    def feature_redetect(self, slot, n_streams, fp, existing):
        """existing: list of (k,2) float64 arrays, one per stream."""
        ex = np.zeros((n_streams, self.max_pts, 2), np.float64)
        ne = np.zeros(n_streams, np.int32)
        for s, e in enumerate(existing):
            ex[s, :len(e)] = e
            ne[s] = len(e)
        xy = np.zeros((n_streams, self.max_pts, 2), np.float32)
        n = np.zeros(n_streams, np.int32)
        self._chk(self.lib.flv_feature_redetect(self.h, slot, n_streams, C.byref(fp), _ptr(ex), _ptr(ne), _ptr(xy),
                                                _ptr(n), MEM_HOST))
        return [xy[s, :n[s]].copy() for s in range(n_streams)]

    def fundamental_ransac(self, from_xy, to_xy, n_pts, thr=5.0):
        """from_xy/to_xy: (S,max_pts,2) f32 host arrays. Returns (mask (S,max_pts) u8, F (S,3,3), n_inliers (S,))."""
        S = from_xy.shape[0]
        a = np.ascontiguousarray(from_xy, np.float32); b = np.ascontiguousarray(to_xy, np.float32)
        n = np.ascontiguousarray(n_pts, np.int32)
        mask = np.zeros((S, self.max_pts), np.uint8); F = np.zeros((S, 9)); ni = np.zeros(S, np.int32)
        prm = RansacParams(thr, 0.99, 1000)
        self._chk(self.lib.flv_fundamental_ransac(self.h, S, _ptr(n), _ptr(a), _ptr(b), C.byref(prm), _ptr(mask), _ptr(F), _ptr(ni), MEM_HOST))
        return mask, F.reshape(S, 3, 3), ni

    def pnp_ransac(self, p3d, p2d, n_pts, K4, T_in, thr=3.0):
        S = p3d.shape[0]
        x = np.ascontiguousarray(p3d, np.float32); u = np.ascontiguousarray(p2d, np.float32)
        n = np.ascontiguousarray(n_pts, np.int32); K = np.ascontiguousarray(K4, np.float64); Ti = np.ascontiguousarray(T_in, np.float64)
        mask = np.zeros((S, self.max_pts), np.uint8); To = np.zeros((S, 7)); ni = np.zeros(S, np.int32)
        prm = RansacParams(thr, 0.99, 100)
        self._chk(self.lib.flv_pnp_ransac(self.h, S, _ptr(n), _ptr(x), _ptr(u), _ptr(K), _ptr(Ti), C.byref(prm), _ptr(To), _ptr(mask), _ptr(ni), MEM_HOST))
        return To, mask, ni

    def upload_color(self, slot, imgs, is_rgb=False):
        """imgs: (S,h,w,3|4) u8 host array, interleaved BGR(A) (or RGB(A) with is_rgb)."""
        a = np.ascontiguousarray(imgs, np.uint8)
        S, h, w, ch = a.shape
        self._chk(self.lib.flv_upload_color_images(self.h, slot, S, _ptr(a), w * ch, h * w * ch, ch, 1 if is_rgb else 0, MEM_HOST))

    def set_equalize_hist(self, enable):
        self._chk(self.lib.flv_set_equalize_hist(self.h, 1 if enable else 0))

    def feature_prepare(self, slot, n_streams, fp, redetect=True):
        self._chk(self.lib.flv_feature_prepare(self.h, slot, n_streams, C.byref(fp), 1 if redetect else 0))

    def feature_redetect_dev(self, slot, n_streams, fp, d_exist, d_nexist, d_xy, d_n):
        self._chk(self.lib.flv_feature_redetect(self.h, slot, n_streams, C.byref(fp), C.c_void_p(d_exist),
                                                C.c_void_p(d_nexist), C.c_void_p(d_xy), C.c_void_p(d_n), MEM_DEVICE))
