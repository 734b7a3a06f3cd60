"""Restatement of cv::goodFeaturesToTrack as FLVIS calls it -- TEST INFRASTRUCTURE (oracle).

Call sites: /root/reference/src/processing/feature_dem.cpp:160 (N, q, d, mask = all 255) and
:221 (2N, q, d).  The algorithm is OpenCV's (external dependency, not in /root/reference;
pinned to opencv 4.13.0 by tests/golden/gftt_*.npz); SURVEY.md Appendix A.2 states it.

Float contract shared with the CUDA kernel (flvis_b200/csrc/gftt.cu), found by probing
cv2 4.13.0 (x86-64 wheel, AVX2/AVX-512 dispatch) bit-for-bit:
  scale = f32(1/(4*3*255));  k1 = scale, k0 = 2*scale
  Dx  = fma(k1, (r[y-1]+r[y+1]), k0*r[y])         r = I[x+1]-I[x-1]  (exact ints)
  s   = fma(k1, I[x+1], fma(k0, I[x], k1*I[x-1]))  for x <  w - w%32   (vector body, FMA)
        (k1*I[x-1] + k0*I[x]) + k1*I[x+1]           for x >= w - w%32   (scalar tail, no FMA)
  Dy  = s[y+1] - s[y-1]
  cov = (Dx*Dx, Dx*Dy, Dy*Dy) each rounded to f32; 3x3 box sums are exact (double), rounded once
  a = 0.5*Sxx, b = Sxy, c = 0.5*Syy;  eig = (a+c) - sqrt((a-c)*(a-c) + b*b)   (no FMA)
  all borders REFLECT_101.
"""
import numpy as np

f32 = np.float32
f64 = np.float64


def _reflect101(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * n - 2 - i, i)


def _fma(a, b, c):
    # a*b is exact in f64 (24x24 bits); the f64 add then f32 round double-rounds only on
    # measure-zero ties that cannot occur here (operands are small-integer multiples of k1).
    return (f64(a) * f64(b) + f64(c)).astype(f32)


def sobel_dx_dy(img):
    h, w = img.shape
    scale = f32(1.0 / (4 * 3 * 255.0))
    k1 = scale
    k0 = f32(2) * scale
    p = img.astype(np.int32)
    xm = _reflect101(np.arange(w) - 1, w); xp = _reflect101(np.arange(w) + 1, w)
    ym = _reflect101(np.arange(h) - 1, h); yp = _reflect101(np.arange(h) + 1, h)
    r = (p[:, xp] - p[:, xm]).astype(f32)
    dx = _fma(k1, (r[ym] + r[yp]).astype(f32), (k0 * r).astype(f32))
    S0 = p[:, xm].astype(f32); S1 = p.astype(f32); S2 = p[:, xp].astype(f32)
    s_fma = _fma(k1, S2, _fma(k0, S1, (k1 * S0).astype(f32)))
    s_no = (((k1 * S0).astype(f32) + (k0 * S1).astype(f32)).astype(f32) + (k1 * S2).astype(f32)).astype(f32)
    body = w - (w % 32)
    s = np.where(np.arange(w)[None, :] < body, s_fma, s_no)
    dy = (s[yp] - s[ym]).astype(f32)
    return dx, dy


def _box3(a):
    h, w = a.shape
    ys = _reflect101(np.arange(-1, h + 1), h); xs = _reflect101(np.arange(-1, w + 1), w)
    q = a[ys][:, xs].astype(f64)
    s = q[0:h, 0:w].copy()
    for i in range(3):
        for j in range(3):
            if i or j:
                s += q[i:i + h, j:j + w]
    return s.astype(f32)


def corner_min_eigen_val(img):
    """cv2.cornerMinEigenVal(img, blockSize=3, ksize=3) for u8 input."""
    dx, dy = sobel_dx_dy(img)
    a = (_box3((dx * dx).astype(f32)) * f32(0.5)).astype(f32)
    b = _box3((dx * dy).astype(f32))
    c = (_box3((dy * dy).astype(f32)) * f32(0.5)).astype(f32)
    t = (a - c).astype(f32)
    return ((a + c).astype(f32) - np.sqrt(((t * t).astype(f32) + (b * b).astype(f32)).astype(f32))).astype(f32)


def good_features_to_track(img, max_corners, quality, min_dist, eig=None, return_all=False):
    """Returns (n,2) f32 integer-valued corner coordinates, response-descending, greedy min-dist."""
    h, w = img.shape
    if eig is None:
        eig = corner_min_eigen_val(img)
    max_val = eig.max()
    thr = f32(f64(max_val) * quality)            # cv::threshold(eig, eig, maxVal*quality, 0, THRESH_TOZERO), f32 thresh
    eig_t = np.where(eig > thr, eig, f32(0))
    # 3x3 dilate (max filter, border = -inf)
    pad = np.full((h + 2, w + 2), -np.inf, f32); pad[1:-1, 1:-1] = eig_t
    dil = pad[0:h, 0:w].copy()
    for i in range(3):
        for j in range(3):
            dil = np.maximum(dil, pad[i:i + h, j:j + w])
    m = (eig_t != 0) & (eig_t == dil)
    m[0, :] = False; m[-1, :] = False; m[:, 0] = False; m[:, -1] = False
    ys, xs = np.nonzero(m)
    vals = eig_t[ys, xs]
    addr = ys.astype(np.int64) * w + xs
    order = np.lexsort((-addr, -vals.astype(f64)))     # value desc, ties: higher address first
    ys, xs = ys[order], xs[order]
    if return_all:
        return np.stack([xs, ys], 1).astype(f32), vals[order]
    out = []
    if min_dist >= 1:
        cell = int(round(min_dist))
        gw = (w + cell - 1) // cell; gh = (h + cell - 1) // cell
        grid = [[[] for _ in range(gw)] for _ in range(gh)]
        md2 = min_dist * min_dist
        for y, x in zip(ys.tolist(), xs.tolist()):
            xc, yc = x // cell, y // cell
            good = True
            for yy in range(max(0, yc - 1), min(gh - 1, yc + 1) + 1):
                for xx in range(max(0, xc - 1), min(gw - 1, xc + 1) + 1):
                    for (px, py) in grid[yy][xx]:
                        if (x - px) ** 2 + (y - py) ** 2 < md2:
                            good = False
                            break
                    if not good:
                        break
                if not good:
                    break
            if good:
                grid[yc][xc].append((x, y))
                out.append((x, y))
                if 0 < max_corners <= len(out):
                    break
    else:
        for y, x in zip(ys.tolist(), xs.tolist()):
            out.append((x, y))
            if 0 < max_corners <= len(out):
                break
    return np.array(out, f32).reshape(-1, 2)
