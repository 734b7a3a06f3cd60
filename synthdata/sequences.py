"""BASELINE-shaped synthetic sequences -- TEST INFRASTRUCTURE.

A textured plane (synthdata/textures.py) stands in front of the rig; every camera image is a ray-plane rendering of it
through the camera's own lens model, the IMU samples are the analytic derivatives of the same rig trajectory in FLVIS's
internal convention (`acc = R^T (a_w + (0,0,-9.81))`, since `v' = R acc - g` with g = (0,0,-9.81):
/root/reference/src/processing/vi_motion.cpp:25,197).  Per-pixel arithmetic uses only + - * / (IEEE-exact), rotations come
from `math` (libm), so the images are bit-reproducible on any machine with this image's Python.

Configurations (values copied from the reference's launch files):
  c0  launch/d435i/sn943222072828_depth.yaml   640x480 depth + IMU 200 Hz, 30 Hz images, skip_first_n_imgs = 50
  c1  launch/EuRoC_MAV/euroc.yaml              752x480 STEREO_UNRECT (radtan lenses, cv::stereoRectify), equalizeHist,
                                               IMU 200 Hz, 20 Hz images, local-map window 10
  c3  launch/KITTI/KITTI.yaml                  1241x376 rectified stereo, no IMU, 10 Hz, window_size overridden to 20
"""
import math

import numpy as np

from . import textures as synth
from .se3 import SE3, R2q, q2R, rpy2R

G = 9.81


# ---- rig trajectory (body = IMU frame "i", world z up) ----------------------------------------------------------------
class Trajectory:
    """p(t), rpy(t) as sums of sinusoids that start from rest after `t_rest` seconds (smooth ramp)."""

    def __init__(self, pos_amp, pos_w, rpy_amp, rpy_w, t_rest=0.6, vel=(0.0, 0.0, 0.0), phase=0.0):
        self.pa, self.pw = np.array(pos_amp, float), np.array(pos_w, float)
        self.ra, self.rw = np.array(rpy_amp, float), np.array(rpy_w, float)
        self.t_rest = t_rest
        self.vel = np.array(vel, float)
        self.phase = float(phase)            # phase of the sinusoids (the pose at s = 0 stays the identity)

    def _s(self, t):
        """motion clock: 0 before t_rest, then a C2 ramp into s = t - t_rest - 0.5."""
        u = t - self.t_rest
        if u <= 0:
            return 0.0, 0.0, 0.0
        if u < 1.0:                       # s = u^3 - u^4/2 : s'(0)=s''(0)=0, s'(1)=1, s''(1)=0
            return u ** 3 - 0.5 * u ** 4, 3 * u ** 2 - 2 * u ** 3, 6 * u - 6 * u ** 2
        return u - 0.5, 1.0, 0.0

    def eval(self, t):
        """-> p, v, a (world), rpy, rpy', (ignored rpy'') at time t."""
        s, sd, sdd = self._s(t)
        ph = self.phase
        sin_p, cos_p = np.sin(self.pw * s + ph), np.cos(self.pw * s + ph)
        p = self.pa * (math.cos(ph) - cos_p) + self.vel * s
        dp = self.pa * self.pw * sin_p + self.vel
        ddp = self.pa * self.pw ** 2 * cos_p
        v = dp * sd
        a = ddp * sd * sd + dp * sdd
        sin_r, cos_r = np.sin(self.rw * s + ph), np.cos(self.rw * s + ph)
        rpy = self.ra * (sin_r - math.sin(ph))
        drpy = self.ra * self.rw * cos_r * sd
        return p, v, a, rpy, drpy

    def T_w_i(self, t):
        p, _, _, rpy, _ = self.eval(t)
        return SE3(R2q(rpy2R(rpy)), p)

    def imu(self, t):
        """(acc, gyro) in the body frame, FLVIS internal convention."""
        p, v, a, rpy, d = self.eval(t)
        R = rpy2R(rpy)
        r, pt = rpy[0], rpy[1]
        gyro = np.array([d[0] - d[2] * math.sin(pt),
                         d[1] * math.cos(r) + d[2] * math.sin(r) * math.cos(pt),
                         -d[1] * math.sin(r) + d[2] * math.cos(r) * math.cos(pt)])
        acc = R.T @ (a + np.array([0.0, 0.0, -G]))
        return acc, gyro


# ---- cameras -----------------------------------------------------------------------------------------------------------
def _undistort_normalized(xd, yd, D, iters=25):
    """inverse of the radtan model on normalized coordinates (fixed-point iteration, arithmetic ops only)."""
    k1, k2, p1, p2 = [float(v) for v in D[:4]]
    x, y = xd.copy(), yd.copy()
    for _ in range(iters):
        r2 = x * x + y * y
        icd = 1.0 / (1.0 + (k2 * r2 + k1) * r2)
        dx = 2 * p1 * x * y + p2 * (r2 + 2 * x * x)
        dy = p1 * (r2 + 2 * y * y) + 2 * p2 * x * y
        x = (xd - dx) * icd
        y = (yd - dy) * icd
    return x, y


class PlaneRenderer:
    """Renders the plane x_w = X0 (textured on its (-y, -z) axes) as seen by a camera with intrinsics K4 = fx fy cx cy and
    radtan distortion D; canvas pixel (cu, cv) = (cu0 - y*ppm, cv0 - z*ppm)."""

    def __init__(self, canvas, X0, ppm, w, h, K4, D=None, origin=(0.0, 0.0)):
        self.canvas = canvas.astype(np.float64)
        self.X0, self.ppm, self.w, self.h = float(X0), float(ppm), w, h
        ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
        xn = (xs - K4[2]) / K4[0]; yn = (ys - K4[3]) / K4[1]
        if D is not None and np.any(np.asarray(D) != 0):
            xn, yn = _undistort_normalized(xn, yn, D)
        self.xn, self.yn = xn, yn
        # world (y, z) = origin lands on the canvas centre
        self.cu0 = canvas.shape[1] / 2.0 + origin[0] * self.ppm; self.cv0 = canvas.shape[0] / 2.0 + origin[1] * self.ppm

    def render(self, T_w_c, want_depth=False):
        R = q2R(T_w_c.q); c = T_w_c.t
        # ray direction in the world for the camera ray (xn, yn, 1)
        dx = R[0, 0] * self.xn + R[0, 1] * self.yn + R[0, 2]
        dy = R[1, 0] * self.xn + R[1, 1] * self.yn + R[1, 2]
        dz = R[2, 0] * self.xn + R[2, 1] * self.yn + R[2, 2]
        s = (self.X0 - c[0]) / dx                       # = depth along the optical axis (ray has z_c = 1)
        py = c[1] + s * dy; pz = c[2] + s * dz
        cu = self.cu0 - py * self.ppm; cv = self.cv0 - pz * self.ppm
        H, W = self.canvas.shape
        x0 = np.floor(cu); y0 = np.floor(cv)
        fx = cu - x0; fy = cv - y0
        x0 = x0.astype(np.int64); y0 = y0.astype(np.int64)
        x0c = np.clip(x0, 0, W - 1); x1c = np.clip(x0 + 1, 0, W - 1)
        y0c = np.clip(y0, 0, H - 1); y1c = np.clip(y0 + 1, 0, H - 1)
        cvs = self.canvas
        v = (cvs[y0c, x0c] * (1 - fx) * (1 - fy) + cvs[y0c, x1c] * fx * (1 - fy) + cvs[y1c, x0c] * (1 - fx) * fy + cvs[y1c, x1c] * fx * fy)
        img = np.clip(np.floor(v + 0.5), 0, 255).astype(np.uint8)
        if want_depth:
            return img, s
        return img


# ---- sequence configurations ------------------------------------------------------------------------------------------
def _mat44_to_se3(m):
    m = np.array(m, float).reshape(4, 4)
    return SE3(R2q(m[:3, :3]), m[:3, 3])


EUROC = dict(
    w=752, h=480,
    K0=(458.654, 457.296, 367.215, 248.375), D0=(-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05),
    K1=(457.587, 456.134, 379.999, 255.238), D1=(-0.28368365, 0.07451284, -0.00010473, -3.55590700e-05),
    T_imu_mavimu=[0.0, 0.0, 1.0, 0.0, 0.0, -1.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0],
    T_mavimu_cam0=[0.0148655429818, -0.999880929698, 0.00414029679422, -0.0216401454975,
                   0.999557249008, 0.0149672133247, 0.025715529948, -0.064676986768,
                   -0.0257744366974, 0.00375618835797, 0.999660727178, 0.00981073058949, 0.0, 0.0, 0.0, 1.0],
    T_mavimu_cam1=[0.0125552670891, -0.999755099723, 0.0182237714554, -0.0198435579556,
                   0.999598781151, 0.0130119051815, 0.0251588363115, 0.0453689425024,
                   -0.0253898008918, 0.0179005838253, 0.999517347078, 0.00786212447038, 0.0, 0.0, 0.0, 1.0],
    feature_para=[30, 20, 5, 1000, 0.01, 10], vi_para=[0.1, 0.01, 0.001, 0.001, 0.3, 0.1], dc_para=[0.90, 50.0, 1.0],
    window=10, img_hz=20.0)
D435 = dict(
    w=640, h=480, K0=(384.16455078125, 384.16455078125, 320.2144470214844, 238.94403076171875), depth_factor=1000.0,
    T_imu_cam0=[0.0, 0.0, 1.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, -1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 1.0],
    feature_para=[30, 15, 5, 500, 0.01, 15], vi_para=[0.1, 0.01, 0.001, 0.001, 0.1, 0.1], dc_para=[0.98, 40.0, 1.0],
    window=8, img_hz=30.0, skip=50)
KITTI = dict(
    w=1241, h=376, K0=(718.856, 718.856, 607.1928, 185.2157), bf=386.1448,
    feature_para=[30, 15, 10, 2000, 0.0001, 10], vi_para=[0.1, 0.03, 0.003, 0.01, 0.5, 0.1], dc_para=[0.8, 1000.0, 0.0],
    window=20, img_hz=10.0)


class Sequence:
    """frames(): yields (t, img0, img1, imu_samples_before_this_frame) ; imu samples are (t, acc, gyro) tuples."""

    def __init__(self, name, cam_type, n_frames, img_hz, traj, T_i_c0, render0, render1=None, T_c0_c1=None, imu_hz=200.0,
                 blank_frames=(), depth_factor=1000.0, seed=0):
        self.name, self.cam_type, self.n_frames, self.img_hz, self.traj = name, cam_type, n_frames, img_hz, traj
        self.T_i_c0, self.r0, self.r1, self.T_c0_c1 = T_i_c0, render0, render1, T_c0_c1
        self.imu_hz, self.blank, self.depth_factor = imu_hz, set(blank_frames), depth_factor
        self.rng = np.random.default_rng(seed + 77)
        self.imu_noise = (0.02, 0.002)

    def T_w_c0(self, t):
        return self.traj.T_w_i(t) * self.T_i_c0

    def frames(self):
        k_imu = 0
        for k in range(self.n_frames):
            t = k / self.img_hz
            imu = []
            if self.imu_hz:
                while k_imu / self.imu_hz <= t + 1e-9:
                    ti = k_imu / self.imu_hz
                    acc, gyro = self.traj.imu(ti)
                    acc = acc + self.rng.normal(0, self.imu_noise[0], 3); gyro = gyro + self.rng.normal(0, self.imu_noise[1], 3)
                    imu.append((ti, acc, gyro))
                    k_imu += 1
            T0 = self.T_w_c0(t)
            if self.cam_type == "depth":
                img0, z = self.r0.render(T0, want_depth=True)
                img1 = np.clip(np.floor(z * self.depth_factor + 0.5), 0, 65535).astype(np.uint16)
            else:
                img0 = self.r0.render(T0)
                img1 = self.r1.render(T0 * self.T_c0_c1)
            if k in self.blank:                          # an unrelated view (the image turned by 180 degrees): tracking must fail
                img0 = np.ascontiguousarray(img0[::-1, ::-1])
                if self.cam_type != "depth":
                    img1 = np.ascontiguousarray(img1[::-1, ::-1])
            yield t, img0, img1, imu


def make_c0(n_frames=150, seed=0, blank_frames=(), skip=None, t_rest=1.9, imu=True):
    """C0: 640x480 D435i depth + IMU; the first 50 images are skipped by the tracker (vo_tracking.cpp:171).
    skip / t_rest / imu: overrides for short test sequences (fewer skipped images, earlier motion, no IMU samples)."""
    c = D435 if skip is None else dict(D435, skip=skip)
    X0 = 3.0; ppm = c["K0"][0] / X0
    canvas = synth.texture(1000 * 0 + seed, int(5.2 * ppm), int(7.0 * ppm), blur=2)
    traj = Trajectory(pos_amp=(0.10, 0.35, 0.20), pos_w=(0.9, 0.7, 0.8), rpy_amp=(0.06, 0.05, 0.08), rpy_w=(0.8, 0.6, 0.5), t_rest=t_rest)
    T_i_c0 = _mat44_to_se3(c["T_imu_cam0"])
    r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"])
    seq = Sequence("c0", "depth", n_frames, c["img_hz"], traj, T_i_c0, r0, blank_frames=blank_frames, depth_factor=c["depth_factor"], seed=seed,
                   imu_hz=200.0 if imu else 0)
    seq.cfg = c
    return seq


def euroc_rig():
    """T_i_c0, T_c0_c1 exactly as vo_tracking.cpp:222-234 composes them."""
    c = EUROC
    T_mavi_c0 = _mat44_to_se3(c["T_mavimu_cam0"]); T_mavi_c1 = _mat44_to_se3(c["T_mavimu_cam1"])
    T_i_mavi = _mat44_to_se3(c["T_imu_mavimu"])
    T_c0_c1 = T_mavi_c0.inverse() * T_mavi_c1
    return T_i_mavi * T_mavi_c0, T_c0_c1


def make_c1(n_frames=200, seed=0, blank_frames=(), t_rest=0.5):
    """C1: EuRoC-shaped raw stereo (radtan distortion => STEREO_UNRECT), 20 Hz images, 200 Hz IMU."""
    c = EUROC
    X0 = 3.0; ppm = c["K0"][0] / X0
    canvas = synth.texture(1000 * 1 + seed, int(6.0 * ppm), int(8.5 * ppm), blur=2)
    traj = Trajectory(pos_amp=(0.12, 0.45, 0.25), pos_w=(0.5, 0.45, 0.55), rpy_amp=(0.05, 0.04, 0.07), rpy_w=(0.5, 0.45, 0.35), t_rest=t_rest)
    T_i_c0, T_c0_c1 = euroc_rig()
    r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"], c["D0"])
    r1 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K1"], c["D1"])
    seq = Sequence("c1", "stereo_unrect", n_frames, c["img_hz"], traj, T_i_c0, r0, r1, T_c0_c1, blank_frames=blank_frames, seed=seed)
    seq.cfg = c
    return seq


def make_c3(n_frames=34, seed=0):
    """C3: KITTI-shaped rectified stereo, no IMU, 10 Hz; sideways + forward motion of ~0.35 m per frame so that every
    frame is a keyframe (f2f_tracking.cpp:345-354) and the 20-KF window fills."""
    c = KITTI
    X0 = 12.0; ppm = c["K0"][0] / X0
    canvas = synth.texture_multiscale(1000 * 3 + seed, int(11.0 * ppm), int(40.0 * ppm))
    traj = Trajectory(pos_amp=(0.4, 0.0, 0.15), pos_w=(0.25, 0.0, 0.3), rpy_amp=(0.0, 0.004, 0.01), rpy_w=(0.0, 0.5, 0.3), t_rest=0.0,
                      vel=(0.0, 3.2, 0.0))
    T_i_c0 = _mat44_to_se3(D435["T_imu_cam0"])            # no IMU: any level body->camera rotation
    b = c["bf"] / c["K0"][0]
    T_c0_c1 = SE3([1.0, 0, 0, 0], [b, 0, 0])
    r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"], origin=(0.5 * 0.32 * n_frames, 0.0))
    seq = Sequence("c3", "stereo", n_frames, c["img_hz"], traj, T_i_c0, r0, r0, T_c0_c1, imu_hz=0, seed=seed)
    seq.cfg = c
    return seq


# ---- periodic workloads for bench.py --------------------------------------------------------------------------------------
def make_bench(workload, stream_id, period_frames=40, render=True):
    """An endless sequence of the named BASELINE shape for throughput runs: after the start-up ramp the rig trajectory is
    periodic with period `period_frames` images, so a pool of one period of rendered frames (plus the start-up frames) serves
    any number of steps.  Motion is sized so that a keyframe falls on every 3rd-5th frame (f2f_tracking.cpp:339-355).
    The sequence carries n_startup / period: frames [0, n_startup) are played once, then frames [n_startup, n_startup + period)
    repeat.  render=False builds only the trajectory and rig (ground-truth poses), not the textures."""
    seed = 7000 + stream_id
    rng = np.random.default_rng(seed)
    jitter = 1.0 + 0.15 * (rng.uniform(size=6) - 0.5)
    phase = 2 * math.pi * float(rng.uniform())     # cameras are not synchronised: keyframes of different streams fall on different frames
    if workload == "euroc":
        c = EUROC; hz = c["img_hz"]
    elif workload == "d435":
        c = D435; hz = c["img_hz"]
    else:
        c = KITTI; hz = c["img_hz"]
    Tp = period_frames / hz
    w0 = 2 * math.pi / Tp
    t_rest = 0.4
    n_startup = int(math.ceil((t_rest + 1.0) * hz)) + 1            # rest + ramp: afterwards s = t - t_rest - 0.5 is linear in t
    if workload == "euroc":
        X0 = 3.0; ppm = c["K0"][0] / X0
        canvas = synth.texture(seed, int(5.0 * ppm), int(7.0 * ppm), blur=2) if render else np.zeros((8, 8), np.uint8)
        traj = Trajectory(pos_amp=(0.02 * jitter[0], 0.06 * jitter[1], 0.04 * jitter[2]), pos_w=(w0, w0, 2 * w0),
                          rpy_amp=(0.03 * jitter[3], 0.02 * jitter[4], 0.04 * jitter[5]), rpy_w=(w0, 2 * w0, w0), t_rest=t_rest, phase=phase)
        T_i_c0, T_c0_c1 = euroc_rig()
        r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"], c["D0"]) if render else None
        r1 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K1"], c["D1"]) if render else None
        seq = Sequence("c1", "stereo_unrect", n_startup + period_frames, hz, traj, T_i_c0, r0, r1, T_c0_c1, seed=seed)
    elif workload == "d435":
        X0 = 3.0; ppm = c["K0"][0] / X0
        canvas = synth.texture(seed, int(4.6 * ppm), int(6.2 * ppm), blur=2) if render else np.zeros((8, 8), np.uint8)
        traj = Trajectory(pos_amp=(0.02 * jitter[0], 0.05 * jitter[1], 0.035 * jitter[2]), pos_w=(w0, w0, 2 * w0),
                          rpy_amp=(0.03 * jitter[3], 0.02 * jitter[4], 0.04 * jitter[5]), rpy_w=(w0, 2 * w0, w0), t_rest=t_rest, phase=phase)
        c = dict(c, skip=0)
        r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"]) if render else None
        seq = Sequence("c0", "depth", n_startup + period_frames, hz, traj, _mat44_to_se3(c["T_imu_cam0"]), r0, depth_factor=c["depth_factor"], seed=seed)
    else:
        X0 = 12.0; ppm = c["K0"][0] / X0
        canvas = synth.texture_multiscale(seed, int(9.0 * ppm), int(26.0 * ppm)) if render else np.zeros((8, 8), np.uint8)
        traj = Trajectory(pos_amp=(0.3 * jitter[0], 1.2 * jitter[1], 0.2 * jitter[2]), pos_w=(w0, w0, 2 * w0),
                          rpy_amp=(0.0, 0.004, 0.01), rpy_w=(w0, w0, w0), t_rest=0.0, phase=phase)
        n_startup = int(math.ceil(1.0 * hz)) + 1
        b = c["bf"] / c["K0"][0]
        r0 = PlaneRenderer(canvas, X0, ppm, c["w"], c["h"], c["K0"]) if render else None
        seq = Sequence("c3", "stereo", n_startup + period_frames, hz, traj, _mat44_to_se3(D435["T_imu_cam0"]), r0, r0,
                       SE3([1.0, 0, 0, 0], [b, 0, 0]), imu_hz=0, seed=seed)
    seq.cfg = c
    seq.n_startup, seq.period = n_startup, period_frames
    return seq


def render_pool(args):
    """(workload, stream_id, period) -> (img0 [n,h,w] u8, img1 [n,h,w] u8|u16, imu list per frame); picklable entry point for a
    process pool."""
    workload, stream_id, period = args
    seq = make_bench(workload, stream_id, period)
    f0, f1, imu = [], [], []
    for t, a, b, samples in seq.frames():
        f0.append(a); f1.append(b)
        imu.append(np.array([[ti, *acc, *gyro] for ti, acc, gyro in samples], np.float64).reshape(-1, 7))
    return np.stack(f0), np.stack(f1), imu
