"""Seeded synthetic inputs (pure numpy, bit-reproducible on any machine) -- TEST INFRASTRUCTURE.

SURVEY.md 8(d): textures are blurred uint8 noise, frames are warps of a larger canvas so the ground-truth
flow is analytic.  Nothing here calls cv2, so tests/golden fixtures only need to store seeds + expected outputs.
"""
import numpy as np


def _box_blur_int(a, r):
    """Separable box blur with radius r on an int64 array (edge-replicated), exact integer arithmetic."""
    if r <= 0:
        return a
    k = 2 * r + 1
    for axis in (0, 1):
        pad = [(0, 0), (0, 0)]
        pad[axis] = (r + 1, r)
        p = np.pad(a, pad, mode="edge")
        c = np.cumsum(p, axis=axis)
        if axis == 0:
            a = c[k:, :] - c[:-k, :]
        else:
            a = c[:, k:] - c[:, :-k]
        a = a // k
    return a


def texture(seed, h, w, blur=2, passes=3):
    """uint8 texture: uniform noise, `passes` box blurs of radius `blur` (~Gaussian), stretched to 0..255."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 256, (h, w), dtype=np.int64) * 256
    for _ in range(passes):
        a = _box_blur_int(a, blur)
    lo, hi = int(a.min()), int(a.max())
    return ((a - lo) * 255 // max(hi - lo, 1)).astype(np.uint8)


def texture_multiscale(seed, h, w, blurs=(2, 6, 16), weights=(2, 3, 4)):
    """uint8 texture with structure at several scales (sum of blurred noise fields): coarse pyramid levels keep
    gradients, so large displacements (KITTI-sized disparities) are trackable."""
    rng = np.random.default_rng(seed)
    acc = np.zeros((h, w), np.int64)
    for b, wt in zip(blurs, weights):
        a = rng.integers(0, 256, (h, w), dtype=np.int64) * 256
        for _ in range(3):
            a = _box_blur_int(a, b)
        lo, hi = int(a.min()), int(a.max())
        acc += wt * ((a - lo) * 1024 // max(hi - lo, 1))
    lo, hi = int(acc.min()), int(acc.max())
    return ((acc - lo) * 255 // max(hi - lo, 1)).astype(np.uint8)


def warp_affine(canvas, A, out_h, out_w):
    """out(y,x) = bilinear(canvas at A @ [x,y,1]) rounded to uint8; A is 2x3 (dst -> src), float64."""
    ys, xs = np.mgrid[0:out_h, 0:out_w].astype(np.float64)
    sx = A[0, 0] * xs + A[0, 1] * ys + A[0, 2]
    sy = A[1, 0] * xs + A[1, 1] * ys + A[1, 2]
    x0 = np.floor(sx).astype(np.int64); y0 = np.floor(sy).astype(np.int64)
    fx = sx - x0; fy = sy - y0
    H, W = canvas.shape
    x0c = np.clip(x0, 0, W - 1); x1c = np.clip(x0 + 1, 0, W - 1)
    y0c = np.clip(y0, 0, H - 1); y1c = np.clip(y0 + 1, 0, H - 1)
    c = canvas.astype(np.float64)
    v = (c[y0c, x0c] * (1 - fx) * (1 - fy) + c[y0c, x1c] * fx * (1 - fy) +
         c[y1c, x0c] * (1 - fx) * fy + c[y1c, x1c] * fx * fy)
    return np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)


def frame_pair(seed, h, w, shift=(5.0, 3.0), rot_deg=0.0, scale=1.0, blur=2, margin=48):
    """Two frames of one textured plane: I = canvas crop, J = canvas seen after a small similarity motion.
    Returns I, J and flow(x,y) -> (x',y') mapping I pixel coordinates to J pixel coordinates."""
    canvas = texture(seed, h + 2 * margin, w + 2 * margin, blur)
    A0 = np.array([[1.0, 0.0, margin], [0.0, 1.0, margin]])
    I = warp_affine(canvas, A0, h, w)
    th = np.deg2rad(rot_deg)
    cx, cy = w / 2.0, h / 2.0
    # J(x') = canvas(M^-1 ...): define forward map p_J = R*s*(p_I - c) + c + shift, so src = inverse
    R = scale * np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    Rinv = np.linalg.inv(R)
    t = np.array([cx, cy]) + np.array(shift)
    # p_I = Rinv (p_J - t) + c ; canvas coords = p_I + margin
    A1 = np.zeros((2, 3)); A1[:, :2] = Rinv; A1[:, 2] = -Rinv @ t + np.array([cx, cy]) + margin
    J = warp_affine(canvas, A1, h, w)

    def flow(pts):
        p = np.asarray(pts, np.float64)
        return (p - np.array([cx, cy])) @ R.T + t
    return I, J, flow


def grid_points(h, w, n, seed, border=20):
    """n pseudo-random f32 sub-pixel points inside the image."""
    rng = np.random.default_rng(seed)
    x = rng.uniform(border, w - 1 - border, n)
    y = rng.uniform(border, h - 1 - border, n)
    return np.stack([x, y], 1).astype(np.float32)
