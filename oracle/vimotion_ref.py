"""Restatement of VIMOTION (IMU attitude filter / propagation / vision bias feedback) -- TEST INFRASTRUCTURE.

Follows /root/reference/src/processing/vi_motion.cpp:3-464 and /root/reference/src/utils/include/kinetic_math.h:17-141
line by line, with the Sophus/Eigen conventions of 3rdPartLib/Sophus/sophus/{so3,se3}.cpp (products and
constructors normalise, rotation = Eigen _transformVector).  Quaternions are (w,x,y,z).  Kept quirks:
`s *= s.norm()`, float-cast scalar in scalar_multi_q, gyro clamp testing ba_est_norm, (1-para_3) on the gyro
decay.  No golden vectors exist in the reference for this path (no tests, no published numbers).
"""
import math
import numpy as np


def qn(q):
    return q / math.sqrt(float(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]))


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def qrot(q, v):
    w, x, y, z = q
    ux = y * v[2] - z * v[1]; uy = z * v[0] - x * v[2]; uz = x * v[1] - y * v[0]
    ux += ux; uy += uy; uz += uz
    return np.array([v[0] + w * ux + (y * uz - z * uy), v[1] + w * uy + (z * ux - x * uz), v[2] + w * uz + (x * uy - y * ux)])


def q2R(q):
    w, x, y, z = q
    tx, ty, tz = 2 * x, 2 * y, 2 * z
    twx, twy, twz = tx * w, ty * w, tz * w
    txx, txy, txz = tx * x, ty * x, tz * x
    tyy, tyz, tzz = ty * y, tz * y, tz * z
    return np.array([[1 - (tyy + tzz), txy - twz, txz + twy], [txy + twz, 1 - (txx + tzz), tyz - twx],
                     [txz - twy, tyz + twx, 1 - (txx + tyy)]])


def R2q(m):
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)          # w,x,y,z
    if t > 0:
        t = math.sqrt(t + 1.0); q[0] = 0.5 * t; t = 0.5 / t
        q[1] = (m[2, 1] - m[1, 2]) * t; q[2] = (m[0, 2] - m[2, 0]) * t; q[3] = (m[1, 0] - m[0, 1]) * t
    else:
        i = 0
        if m[1, 1] > m[0, 0]: i = 1
        if m[2, 2] > m[i, i]: i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        t = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
        q[1 + i] = 0.5 * t; t = 0.5 / t
        q[0] = (m[k, j] - m[j, k]) * t; q[1 + j] = (m[j, i] + m[i, j]) * t; q[1 + k] = (m[k, i] + m[i, k]) * t
    return q


def rpy2R(rpy):
    r, p, y = rpy
    cy, sy, cp, sp, cr, sr = math.cos(y), math.sin(y), math.cos(p), math.sin(p), math.cos(r), math.sin(r)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def R2rpy(R):
    return np.array([math.atan2(R[2, 1], R[2, 2]), math.atan2(-R[2, 0], math.sqrt(R[2, 1] * R[2, 1] + R[2, 2] * R[2, 2])),
                     math.atan2(R[1, 0], R[0, 0])])


def rpy2Q(rpy):
    return R2q(rpy2R(rpy))


def Q2rpy(q):
    return R2rpy(q2R(q))


class SE3:
    def __init__(self, q=None, t=None, normalize=True):
        self.q = np.array([1.0, 0, 0, 0]) if q is None else (qn(np.array(q, float)) if normalize else np.array(q, float))
        self.t = np.zeros(3) if t is None else np.array(t, float)

    def __mul__(self, o):
        r = SE3()
        r.t = self.t + qrot(self.q, o.t)
        r.q = qn(qmul(self.q, o.q))
        return r

    def inverse(self):
        r = SE3()
        r.q = qn(self.q * np.array([1, -1, -1, -1.0]))
        r.t = qrot(r.q, -self.t)
        return r

    @staticmethod
    def from7(p):
        return SE3([p[3], p[0], p[1], p[2]], p[4:7])

    def to7(self):
        return np.array([self.q[1], self.q[2], self.q[3], self.q[0], *self.t])


def dot(a, b):
    """left-to-right sum of products, the order the C++ expressions use (np.dot / `@` go through BLAS kernels whose
    summation order and FMA use differ in the last bit)."""
    r = a[0] * b[0]
    for i in range(1, len(a)):
        r = r + a[i] * b[i]
    return float(r)


def mat3_vec(R, v):
    return np.array([R[0, 0] * v[0] + R[0, 1] * v[1] + R[0, 2] * v[2], R[1, 0] * v[0] + R[1, 1] * v[1] + R[1, 2] * v[2],
                     R[2, 0] * v[0] + R[2, 1] * v[1] + R[2, 2] * v[2]])


def q1_multi_q2(q1, q2):
    return np.array([q2[0] * q1[0] - q2[1] * q1[1] - q2[2] * q1[2] - q2[3] * q1[3],
                     q2[1] * q1[0] + q2[0] * q1[1] + q2[3] * q1[2] - q2[2] * q1[3],
                     q2[2] * q1[0] - q2[3] * q1[1] + q2[0] * q1[2] + q2[1] * q1[3],
                     q2[3] * q1[0] + q2[2] * q1[1] - q2[1] * q1[2] + q2[0] * q1[3]])


def scalar_multi_q(a, b):
    return float(np.float32(a)) * b           # `const float a` in kinetic_math.h:123


class VIMOTION:
    QUEUE = 400

    def __init__(self, T_i_c, g=9.81, p1=0.1, p2=0.05, p3=0.01, p4=0.01, p5=0.5, p6=0.1):
        self.T_i_c = T_i_c; self.T_c_i = T_i_c.inverse()
        self.acc_bias = np.zeros(3); self.gyro_bias = np.zeros(3)
        self.states = []                      # dicts: pos, vel, q, t
        self.imu_initialized = False; self.is_first = True
        self.g = g; self.gravity = np.array([0, 0, -g])
        self.p1, self.p2, self.p3, self.p4, self.ba_sat, self.bw_sat = p1, p2, p3, p4, p5, p6

    def _qdot(self, q_prev, acc, gyro, gain):
        qdot = scalar_multi_q(0.5, q1_multi_q2(q_prev, np.array([0, gyro[0], gyro[1], gyro[2]])))
        an = math.sqrt(dot(acc, acc))
        if (an - self.g) < 0.3:
            ax, ay, az = acc[0] / an, acc[1] / an, acc[2] / an
            qw, qx, qy, qz = q_prev
            s = np.array([
                2 * qx * (ay + 2 * qw * qx + 2 * qy * qz) - 2 * qy * (ax - 2 * qw * qy + 2 * qx * qz),
                2 * qw * (ay + 2 * qw * qx + 2 * qy * qz) + 2 * qz * (ax - 2 * qw * qy + 2 * qx * qz) - 4 * qx * (-2 * qx * qx - 2 * qy * qy + az + 1),
                2 * qz * (ay + 2 * qw * qx + 2 * qy * qz) - 2 * qw * (ax - 2 * qw * qy + 2 * qx * qz) - 4 * qy * (-2 * qx * qx - 2 * qy * qy + az + 1),
                2 * qx * (ax - 2 * qw * qy + 2 * qx * qz) + 2 * qy * (ay + 2 * qw * qx + 2 * qy * qz)])
            s = s * math.sqrt(dot(s, s))
            qdot = qdot - gain * s
        return qdot

    def imu_feed(self, t, acc, gyro):
        """F2FTracking::imu_feed: init until imu_initialized, then propagate.  Returns (q, pos, vel)."""
        acc = np.asarray(acc, float) - self.acc_bias; gyro = np.asarray(gyro, float) - self.gyro_bias
        if not self.imu_initialized:
            q_out = np.array([1.0, 0, 0, 0]); z = np.zeros(3)
            if self.is_first:
                if (math.sqrt(dot(acc, acc)) - self.g) < 0.3:
                    rpy = np.array([math.atan2(-acc[1], -acc[2]), math.atan2(acc[0], -acc[2]), 0.0])
                    q = rpy2Q(rpy)
                    self._push(dict(pos=z.copy(), vel=z.copy(), q=q, t=t))
                    self.is_first = False
                    q_out = q.copy()
            else:
                dt = t - self.states[-1]["t"]
                qp = self.states[-1]["q"]
                qd = self._qdot(qp, acc, gyro, 10 * self.p1)
                qnew = qn(qp + scalar_multi_q(dt, qd))
                self._push(dict(pos=z.copy(), vel=z.copy(), q=qnew, t=t))
                if len(self.states) > 30:
                    self.imu_initialized = True
            return q_out, z, z
        sp = self.states[-1]
        dt = t - sp["t"]
        R = q2R(sp["q"])
        qd = self._qdot(sp["q"], acc, gyro, self.p1)
        qnew = qn(sp["q"] + scalar_multi_q(dt, qd))
        pos = sp["pos"] + sp["vel"] * dt
        vel = sp["vel"] + (mat3_vec(R, acc) - self.gravity) * dt
        self._push(dict(pos=pos, vel=vel, q=qnew, t=t))
        return qnew, pos, vel

    def _push(self, s):
        self.states.append(s)
        if len(self.states) >= self.QUEUE:
            self.states.pop(0)

    def vision_trigger(self):
        s = dict(self.states[-1]); s["pos"] = np.zeros(3); s["vel"] = np.zeros(3)
        rpy = Q2rpy(s["q"]); rpy[2] = 0
        s["q"] = qn(rpy2Q(rpy))
        self.states = [s]
        return s["q"]

    def find_state_idx(self, time):
        idx = 9999
        for i in range(len(self.states) - 1, -1, -1):
            idx = i
            if not (self.states[i]["t"] - time) > 0:
                break
        return idx if (idx > 0 and idx != 9999) else None

    def correction_from_vision(self, t_curr, Tcw_curr, t_last, Tcw_last):
        il = self.find_state_idx(t_last)
        if il is None: return
        ic = self.find_state_idx(t_curr)
        if ic is None or il == ic: return
        dt = t_curr - t_last
        im = il + int(math.floor((ic - il) / 2)) if False else il + (ic - il) // 2
        st = self.states
        T_w_iA = Tcw_last.inverse() * self.T_c_i; T_w_iB = Tcw_curr.inverse() * self.T_c_i
        T_w_ia = SE3(st[il]["q"], st[il]["pos"]); T_w_ib = SE3(st[ic]["q"], st[ic]["pos"]); T_w_im = SE3(st[im]["q"], st[im]["pos"])
        T_iB_iA = T_w_iB.inverse() * T_w_iA; T_ib_ia = T_w_ib.inverse() * T_w_ia
        qb = T_ib_ia.q; n2 = dot(qb, qb)
        qbi = np.array([qb[0] / n2, -qb[1] / n2, -qb[2] / n2, -qb[3] / n2])
        QBb = qmul(T_iB_iA.q, qbi)
        gyro_est = np.array([QBb[1] / dt, QBb[2] / dt, QBb[3] / dt])
        cnt = ic - il + 1
        vel_imu = np.zeros(3)
        for i in range(il, ic + 1):
            vel_imu = vel_imu + st[i]["vel"]
        vel_imu = vel_imu * (1.0 / cnt)
        vel_vis = (T_w_iB.t - T_w_iA.t) / dt
        dvw = vel_vis - vel_imu
        qm = T_w_im.q; m2 = dot(qm, qm)
        Rm = q2R(np.array([qm[0] / m2, -qm[1] / m2, -qm[2] / m2, -qm[3] / m2]))
        acc_est = -mat3_vec(Rm, dvw) / dt
        T_diff = T_w_iB * T_w_ib.inverse()
        for i in range(ic, len(st)):
            nT = T_diff * SE3(st[i]["q"], st[i]["pos"])
            st[i]["q"] = nT.q; st[i]["pos"] = nT.t; st[i]["vel"] = st[i]["vel"] + dvw
        if math.isnan(acc_est[0]): acc_est = np.zeros(3)
        if math.isnan(gyro_est[0]): gyro_est = np.zeros(3)
        ban = math.sqrt(dot(acc_est, acc_est))
        if ban > self.ba_sat: acc_est = acc_est * (self.ba_sat / ban)
        bwn = math.sqrt(dot(gyro_est, gyro_est))
        if ban > self.bw_sat: gyro_est = gyro_est * (self.bw_sat / bwn)
        if dt < 0.1:
            self.acc_bias = (1 - self.p3) * self.acc_bias + self.p3 * acc_est
            self.gyro_bias = (1 - self.p3) * self.gyro_bias + self.p4 * gyro_est

    def corr_frame_state(self, time):
        i = self.find_state_idx(time)
        if i is None: return None
        T_w_i = SE3(self.states[i]["q"], self.states[i]["pos"])
        return (T_w_i * self.T_i_c).inverse()

    def rp_compensation(self, time, T_c_w):
        T_w_i_before = T_c_w.inverse() * self.T_c_i
        rpy_b = Q2rpy(T_w_i_before.q)
        i = self.find_state_idx(time)
        if i is None: return T_c_w
        rpy_i = Q2rpy(SE3(self.states[i]["q"], self.states[i]["pos"]).q)
        rpy_v = np.array([rpy_i[0], rpy_i[1], rpy_b[2]])
        rpy_a = rpy_b * (1 - self.p2) + rpy_v * self.p2
        return (SE3(rpy2Q(rpy_a), T_w_i_before.t) * self.T_i_c).inverse()


def synth_imu(n, seed=0, rate=200.0, g=9.81):
    """Smooth synthetic IMU in FLVIS's internal convention (acc = R^T (a_w - gravity), gravity=(0,0,-g))."""
    rng = np.random.default_rng(seed)
    t = np.arange(n) / rate
    rpy = np.stack([0.05 * np.sin(0.7 * t), 0.04 * np.cos(0.5 * t), 0.1 * np.sin(0.2 * t)], 1)
    acc = np.zeros((n, 3)); gyro = np.zeros((n, 3))
    for i in range(n):
        R = rpy2R(rpy[i])
        a_w = np.array([0.3 * np.sin(0.9 * t[i]), 0.2 * np.cos(0.6 * t[i]), 0.1 * np.sin(0.4 * t[i])])
        acc[i] = R.T @ (a_w + np.array([0, 0, -g])) + rng.normal(0, 0.02, 3)
        gyro[i] = np.array([0.035 * np.cos(0.7 * t[i]), -0.02 * np.sin(0.5 * t[i]), 0.02 * np.cos(0.2 * t[i])]) + rng.normal(0, 0.002, 3)
    return t + 100.0, acc, gyro
