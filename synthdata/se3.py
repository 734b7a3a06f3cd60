"""Minimal rigid-body helpers for the synthetic rigs (quaternions are (w, x, y, z); rpy = ZYX Euler angles)."""
import math

import numpy as np


def qn(q):
    return q / math.sqrt(float(q @ q))


def qmul(a, b):
    return np.array([a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                     a[0] * b[2] + a[2] * b[0] + a[3] * b[1] - a[1] * b[3], a[0] * b[3] + a[3] * b[0] + a[1] * b[2] - a[2] * b[1]])


def q2R(q):
    w, x, y, z = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R2q(m):
    t = m[0, 0] + m[1, 1] + m[2, 2]
    q = np.zeros(4)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q[:] = (0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s)
    else:
        i = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]])); j = (i + 1) % 3; k = (i + 2) % 3
        s = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0) * 2
        q[1 + i] = 0.25 * s; q[0] = (m[k, j] - m[j, k]) / s; q[1 + j] = (m[j, i] + m[i, j]) / s; q[1 + k] = (m[k, i] + m[i, k]) / s
    return qn(q)


def rpy2R(rpy):
    r, p, y = rpy
    cy, sy, cp, sp, cr, sr = math.cos(y), math.sin(y), math.cos(p), math.sin(p), math.cos(r), math.sin(r)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr], [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


class SE3:
    def __init__(self, q=None, t=None):
        self.q = np.array([1.0, 0, 0, 0]) if q is None else qn(np.array(q, float))
        self.t = np.zeros(3) if t is None else np.array(t, float)

    def __mul__(self, o):
        return SE3(qmul(self.q, o.q), self.t + q2R(self.q) @ o.t)

    def inverse(self):
        qi = self.q * np.array([1, -1, -1, -1.0])
        return SE3(qi, -(q2R(qi) @ self.t))

    def to7(self):
        """[qx qy qz qw tx ty tz] (the C ABI's pose layout)."""
        return np.array([self.q[1], self.q[2], self.q[3], self.q[0], *self.t])
