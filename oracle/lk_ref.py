"""Scalar restatement of OpenCV's pyramidal Lucas-Kanade as FLVIS calls it.

TEST INFRASTRUCTURE (oracle).  Follows:
  * call sites  /root/reference/src/processing/lkorb_tracking.cpp:64-73  (winSize 31x31,
    maxLevel 10, COUNT+EPS 30/1e-3, OPTFLOW_USE_INITIAL_FLOW) and
    /root/reference/src/processing/camera_frame.cpp:124-128 (maxLevel 5);
  * the algorithm is OpenCV's `calcOpticalFlowPyrLK` (external dependency, not in
    /root/reference; pinned to opencv 4.13.0 through tests/golden/lk_*.npz), restated in
    SURVEY.md Appendix A.1.

Arithmetic contract shared with the CUDA kernel (flvis_b200/csrc/lk.cu):
  * all window sums (A11,A12,A22,b1,b2,err) are EXACT integers, converted to f32 once;
  * every f32 operation is an individually rounded IEEE op (no FMA contraction).
  => kernel output must equal this oracle bit-for-bit; oracle vs cv2 is <= 1e-3 px
     (cv2 sums in 4 f32 SIMD lanes, an unspecified order we do not imitate).
"""
import math
import numpy as np

f32 = np.float32
WB = 14                      # bilinear weight bits
FLT_SCALE = f32(1.0 / (1 << 20))
FLT_EPSILON = float(np.finfo(np.float32).eps)


def reflect101(i, n):
    """BORDER_REFLECT_101 index map (valid for -n < i < 2n-1)."""
    i = np.asarray(i)
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def pyr_down(img):
    """cv2.pyrDown for u8: separable [1 4 6 4 1], (sum+128)>>8, REFLECT_101, out=((w+1)/2,(h+1)/2)."""
    h, w = img.shape
    oh, ow = (h + 1) // 2, (w + 1) // 2
    k = np.array([1, 4, 6, 4, 1], np.int32)
    src = img.astype(np.int32)
    xs = reflect101(2 * np.arange(ow)[:, None] + np.arange(-2, 3)[None, :], w)   # (ow,5)
    rows = (src[:, xs] * k[None, None, :]).sum(axis=2)                            # (h,ow)
    ys = reflect101(2 * np.arange(oh)[:, None] + np.arange(-2, 3)[None, :], h)   # (oh,5)
    out = (rows[ys, :] * k[None, :, None]).sum(axis=1)                            # (oh,ow)
    return ((out + 128) >> 8).astype(np.uint8)


def build_pyramid(img, win=31, max_level=10):
    """Levels OpenCV's buildOpticalFlowPyramid keeps: stop before a level with w<=win or h<=win."""
    pyr = [np.ascontiguousarray(img)]
    lvl = 0
    while lvl < max_level:
        nxt = pyr_down(pyr[-1])
        if nxt.shape[1] <= win or nxt.shape[0] <= win:
            break
        pyr.append(nxt)
        lvl += 1
    return pyr


def scharr(img):
    """Unscaled int Scharr derivatives with REFLECT_101 at the image edge (calcSharrDeriv)."""
    h, w = img.shape
    p = img.astype(np.int64)
    ym, yp = reflect101(np.arange(h) - 1, h), reflect101(np.arange(h) + 1, h)
    xm, xp = reflect101(np.arange(w) - 1, w), reflect101(np.arange(w) + 1, w)
    t0 = (p[ym] + p[yp]) * 3 + p * 10          # vertical smooth
    t1 = p[yp] - p[ym]                         # vertical diff
    dx = t0[:, xp] - t0[:, xm]
    dy = (t1[:, xp] + t1[:, xm]) * 3 + t1 * 10
    return dx, dy


def _cvround(x):
    return int(np.rint(x))      # round-half-even == lrint == cvRound


def _weights(a, b):
    one = f32(1)
    s = f32(1 << WB)
    iw00 = _cvround(f32(f32(one - a) * f32(one - b)) * s)
    iw01 = _cvround(f32(a * f32(one - b)) * s)
    iw10 = _cvround(f32(f32(one - a) * b) * s)
    iw11 = (1 << WB) - iw00 - iw01 - iw10
    return iw00, iw01, iw10, iw11


def _descale(x, n):
    return (x + (1 << (n - 1))) >> n


class _Level:
    def __init__(self, I, J, win):
        self.h, self.w = I.shape
        B = win + 1                      # one extra ring so the (win+1)-wide tap window never leaves the array
        self.B = B
        idx_y = reflect101(np.arange(-B, self.h + B), self.h)
        idx_x = reflect101(np.arange(-B, self.w + B), self.w)
        self.I = I.astype(np.int64)[idx_y][:, idx_x]
        self.J = J.astype(np.int64)[idx_y][:, idx_x]
        dx, dy = scharr(I)
        self.DX = np.pad(dx, B)
        self.DY = np.pad(dy, B)


def _interp(A, x0, y0, win, w4, n):
    p = A[y0:y0 + win + 1, x0:x0 + win + 1]
    return _descale(p[:-1, :-1] * w4[0] + p[:-1, 1:] * w4[1] + p[1:, :-1] * w4[2] + p[1:, 1:] * w4[3], n)


def calc_optical_flow_pyr_lk(prev_img, next_img, prev_pts, init_pts, win=31, max_level=10,
                             max_iter=30, eps=1e-3, min_eig_thr=1e-4, pyramids=None,
                             return_iters=False):
    """Returns next_pts (N,2) f32, status (N,) u8, err (N,) f32.  USE_INITIAL_FLOW semantics."""
    prev_pts = np.asarray(prev_pts, f32).reshape(-1, 2)
    nxt = np.array(init_pts, f32).reshape(-1, 2).copy()
    n = len(prev_pts)
    max_iter = min(max(max_iter, 0), 100)
    eps = min(max(eps, 0.0), 10.0)
    eps2 = eps * eps
    if pyramids is None:
        pI = build_pyramid(prev_img, win, max_level)
        pJ = build_pyramid(next_img, win, max_level)
    else:
        pI, pJ = pyramids
    nlev = len(pI)
    levels = [_Level(pI[l], pJ[l], win) for l in range(nlev)]
    status = np.ones(n, np.uint8)
    errs = np.zeros(n, f32)
    iters = np.zeros((n, nlev), np.int32)
    half = f32((win - 1) * 0.5)
    for level in range(nlev - 1, -1, -1):
        L = levels[level]
        h, w, B = L.h, L.w, L.B
        sc = f32(1.0 / (1 << level))
        for i in range(n):
            px = f32(prev_pts[i, 0] * sc); py = f32(prev_pts[i, 1] * sc)
            if level == nlev - 1:
                nx = f32(nxt[i, 0] * sc); ny = f32(nxt[i, 1] * sc)
            else:
                nx = f32(nxt[i, 0] * f32(2)); ny = f32(nxt[i, 1] * f32(2))
            nxt[i] = (nx, ny)
            px = f32(px - half); py = f32(py - half)
            ipx = int(math.floor(px)); ipy = int(math.floor(py))
            if ipx < -win or ipx >= w or ipy < -win or ipy >= h:
                if level == 0:
                    status[i] = 0; errs[i] = 0
                continue
            a = f32(px - f32(ipx)); b = f32(py - f32(ipy))
            w4 = _weights(a, b)
            Iw = _interp(L.I, ipx + B, ipy + B, win, w4, WB - 5)
            Ix = _interp(L.DX, ipx + B, ipy + B, win, w4, WB)
            Iy = _interp(L.DY, ipx + B, ipy + B, win, w4, WB)
            A11 = f32(f32(float(int((Ix * Ix).sum()))) * FLT_SCALE)
            A12 = f32(f32(float(int((Ix * Iy).sum()))) * FLT_SCALE)
            A22 = f32(f32(float(int((Iy * Iy).sum()))) * FLT_SCALE)
            D = f32(f32(A11 * A22) - f32(A12 * A12))
            dA = f32(A11 - A22)
            disc = f32(f32(dA * dA) + f32(f32(f32(4) * A12) * A12))
            min_eig = f32(f32(f32(A22 + A11) - f32(np.sqrt(disc))) / f32(2 * win * win))
            if float(min_eig) < min_eig_thr or float(D) < FLT_EPSILON:
                if level == 0:
                    status[i] = 0
                continue
            D = f32(f32(1) / D)
            nx = f32(nx - half); ny = f32(ny - half)
            pdx = pdy = f32(0)
            for j in range(max_iter):
                inx = int(math.floor(nx)); iny = int(math.floor(ny))
                if inx < -win or inx >= w or iny < -win or iny >= h:
                    if level == 0:
                        status[i] = 0
                    break
                iters[i, level] += 1
                a = f32(nx - f32(inx)); b = f32(ny - f32(iny))
                w4 = _weights(a, b)
                diff = _interp(L.J, inx + B, iny + B, win, w4, WB - 5) - Iw
                b1 = f32(f32(float(int((diff * Ix).sum()))) * FLT_SCALE)
                b2 = f32(f32(float(int((diff * Iy).sum()))) * FLT_SCALE)
                dx = f32(f32(f32(A12 * b2) - f32(A22 * b1)) * D)
                dy = f32(f32(f32(A12 * b1) - f32(A11 * b2)) * D)
                nx = f32(nx + dx); ny = f32(ny + dy)
                nxt[i] = (f32(nx + half), f32(ny + half))
                if float(dx) * float(dx) + float(dy) * float(dy) <= eps2:
                    break
                if j > 0 and abs(float(f32(dx + pdx))) < 0.01 and abs(float(f32(dy + pdy))) < 0.01:
                    nxt[i, 0] = f32(nxt[i, 0] - f32(dx * f32(0.5)))
                    nxt[i, 1] = f32(nxt[i, 1] - f32(dy * f32(0.5)))
                    break
                pdx, pdy = dx, dy
            if level == 0 and status[i]:
                qx = f32(nxt[i, 0] - half); qy = f32(nxt[i, 1] - half)
                inx = int(math.floor(qx)); iny = int(math.floor(qy))
                if inx < -win or inx >= w or iny < -win or iny >= h:
                    status[i] = 0
                    continue
                a = f32(qx - f32(inx)); b = f32(qy - f32(iny))
                w4 = _weights(a, b)
                diff = _interp(L.J, inx + B, iny + B, win, w4, WB - 5) - Iw
                errs[i] = f32(f32(float(int(np.abs(diff).sum()))) * f32(1.0 / (32 * win * win)))
    if return_iters:
        return nxt, status, errs, iters
    return nxt, status, errs


def calc_optical_flow_pyr_lk_c(prev_img, next_img, prev_pts, init_pts, win=31, max_level=10, max_iter=30, eps=1e-3,
                               min_eig_thr=1e-4):
    """The same restatement compiled from oracle/lk_ref.c (bit-identical outputs, ~1000x faster): used where a test has
    to run hundreds of frames live.  tests/test_oracle_cpu.py pins it to calc_optical_flow_pyr_lk above."""
    import ctypes as C
    from .feature_dem_ref import helpers_lib
    lib = helpers_lib()
    lib.lk_ref_track.restype = C.c_int
    lib.lk_ref_track.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                 C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double]
    I = np.ascontiguousarray(prev_img, np.uint8); J = np.ascontiguousarray(next_img, np.uint8)
    assert I.shape == J.shape and I.ndim == 2
    p = np.ascontiguousarray(prev_pts, f32).reshape(-1, 2); q = np.ascontiguousarray(init_pts, f32).reshape(-1, 2)
    n = len(p)
    out = np.zeros((n, 2), f32); st = np.zeros(n, np.uint8); err = np.zeros(n, f32)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.lk_ref_track(vp(I), vp(J), I.shape[1], I.shape[0], n, vp(p), vp(q), vp(out), vp(st), vp(err), win, max_level,
                     max_iter, float(eps), float(min_eig_thr))
    return out, st, err
