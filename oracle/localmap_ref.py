"""Restatement of LocalMapNodeletClass::frame_callback + PoseLMBag -- TEST INFRASTRUCTURE (oracle).

Follows /root/reference/src/backend/vo_localmap.cpp:87-380 and /root/reference/src/backend/poselmbag.cpp:5-208
line by line (state machine, graph edits, the off-by-one keyframe of :226-232, getMultiViewLMs(4)); the g2o
solve is oracle/ba_ref.c.  PARITY UNPINNED for the solve (see ba_ref.c); the bookkeeping is plain integer logic.
"""
import numpy as np

from . import ba_ref


def _g2o_pose(p):
    """g2o::SE3Quat(R(q), t): quaternion -> rotation matrix -> quaternion, w >= 0, unit norm."""
    x, y, z, w = p[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    t = np.trace(R)
    q = np.zeros(4)
    if t > 0:
        s = np.sqrt(t + 1.0); q[3] = 0.5 * s; s = 0.5 / s
        q[0] = (R[2, 1] - R[1, 2]) * s; q[1] = (R[0, 2] - R[2, 0]) * s; q[2] = (R[1, 0] - R[0, 1]) * s
    else:
        i = 0
        if R[1, 1] > R[0, 0]: i = 1
        if R[2, 2] > R[i, i]: i = 2
        j = (i + 1) % 3; k = (j + 1) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0); q[i] = 0.5 * s; s = 0.5 / s
        q[3] = (R[k, j] - R[j, k]) * s; q[j] = (R[j, i] + R[i, j]) * s; q[k] = (R[k, i] + R[i, k]) * s
    if q[3] < 0: q = -q
    q /= np.linalg.norm(q)
    return np.concatenate([q, p[4:7]])


class PoseLMBag:
    def __init__(self, n):
        self.n = n
        self.reset()

    def reset(self):
        self.lms = []                      # [id, count, p3d]
        self.poses = [[0, 0, None] for _ in range(self.n)]   # frame_id, pose_id, pose
        self.wp_init = 0; self.initialized = False; self.newest = 0; self.oldest = 0

    def _find(self, i):
        for k, it in enumerate(self.lms):
            if it[0] == i:
                return k
        return -1

    def add_lm_observation(self, i, p):
        k = self._find(i)
        if k >= 0:
            cnt = self.lms[k][1]
            q = float(cnt) * self.lms[k][2] + p
            cnt += 1
            self.lms[k][1] = cnt; self.lms[k][2] = (1.0 / float(cnt)) * q
            return False
        self.lms.append([i, 1, np.array(p, float)])
        return True

    def add_lm_observation_sliding(self, i, p):
        k = self._find(i)
        if k >= 0:
            self.lms[k][1] += 1
            return False
        self.lms.append([i, 1, np.array(p, float)])
        return True

    def remove_lm_observation(self, i):
        k = self._find(i)
        if k >= 0:
            self.lms[k][1] -= 1
            if self.lms[k][1] == 0:
                del self.lms[k]
                return True
        return False

    def add_pose(self, fid, pose):
        if self.initialized:
            self.newest = self.oldest
            self.poses[self.newest][0] = fid; self.poses[self.newest][2] = pose
            self.oldest += 1
            if self.oldest == self.n: self.oldest = 0
        else:
            self.poses[self.wp_init] = [fid, self.wp_init, pose]
            self.wp_init += 1
            if self.wp_init == self.n:
                self.initialized = True; self.oldest = 0; self.newest = self.n - 1

    def pose_id_by_frame(self, fid):
        for i in range(self.n):
            if self.poses[i][0] == fid:
                return i
        return -1


class LocalMap:
    def __init__(self, window, K):
        self.W, self.K = window, K
        self.bag = PoseLMBag(window)
        self.reset()

    def reset(self):
        self.state = "UN_INITIALIZED"
        self.bag.reset(); self.kfs = []
        self.pose_est = [None] * self.W; self.fixed = -1
        self.lm_est = {}; self.edges = []     # edges: [lm_id, slot, uv]

    def frame_callback(self, kf):
        """kf = dict(frame_id, lm_id (n,), lm_2d (n,2), lm_3d (n,3), T_c_w (7,)). Returns CorrectionInf dict or None."""
        self.kfs.append(kf)
        bag = self.bag
        if self.state == "UN_INITIALIZED":
            if len(self.kfs) >= self.W:
                for f in range(self.W):
                    bag.add_pose(self.kfs[f]["frame_id"], self.kfs[f]["T_c_w"])
                    for i, p in zip(self.kfs[f]["lm_id"], self.kfs[f]["lm_3d"]):
                        bag.add_lm_observation(int(i), np.array(p, float))
                for fid, pid, pose in bag.poses:
                    self.pose_est[pid] = _g2o_pose(np.asarray(pose, float))
                self.fixed = bag.oldest
                for i, c, p in bag.lms:
                    self.lm_est[i] = p.copy()
                self.edges = []
                for f in range(self.W):
                    slot = bag.pose_id_by_frame(self.kfs[f]["frame_id"])
                    for i, uv in zip(self.kfs[f]["lm_id"], self.kfs[f]["lm_2d"]):
                        self.edges.append([int(i), slot, np.array(uv, float)])
                self.state = "OPTIMIZING"
            else:
                return None
        elif self.state == "SLIDING_WINDOW":
            old = bag.oldest
            self.edges = [e for e in self.edges if e[1] != old]
            for i in self.kfs[0]["lm_id"]:
                if bag.remove_lm_observation(int(i)):
                    self.lm_est.pop(int(i), None)
                    self.edges = [e for e in self.edges if e[0] != int(i)]
            kb = self.kfs[-1]
            bag.add_pose(kb["frame_id"], kb["T_c_w"])
            self.pose_est[bag.newest] = _g2o_pose(np.asarray(kb["T_c_w"], float))
            self.fixed = bag.oldest
            for i, p in zip(kb["lm_id"], kb["lm_3d"]):
                if bag.add_lm_observation_sliding(int(i), np.array(p, float)):
                    self.lm_est[int(i)] = np.array(p, float)
            for i, uv in zip(kb["lm_id"], kb["lm_2d"]):
                self.edges.append([int(i), bag.newest, np.array(uv, float)])
            self.state = "OPTIMIZING"
        out = self._solve()
        self.state = "SLIDING_WINDOW"
        self.kfs.pop(0)
        return out

    def _solve(self):
        ids = sorted(self.lm_est)
        idx = {i: k for k, i in enumerate(ids)}
        d = ba_ref.BAData(np.array(self.pose_est), np.array([self.lm_est[i] for i in ids]).reshape(-1, 3),
                          [e[1] for e in self.edges], [idx[e[0]] for e in self.edges],
                          np.array([e[2] for e in self.edges]).reshape(-1, 2), self.K, fixed_pose=self.fixed)
        st = ba_ref.optimize(d, 12, 8)
        for p in range(self.W):
            self.pose_est[p] = d.poses[p].copy()
        for k, i in enumerate(ids):
            self.lm_est[i] = d.lms[k].copy()
        outliers = [self.edges[e][0] for e in range(len(self.edges) - 1, -1, -1) if not d.active[e]]
        self.edges = [e for k, e in enumerate(self.edges) if d.active[k]]
        mv = [it for it in self.bag.lms if it[1] >= 4]
        return {"frame_id": self.kfs[-1]["frame_id"], "T_c_w": self.pose_est[self.bag.newest].copy(),
                "lm_id": [it[0] for it in mv], "lm_3d": np.array([self.lm_est[it[0]] for it in mv]).reshape(-1, 3),
                "outlier_id": outliers, "stats": st}


def make_keyframe_sequence(n_kf, seed=0, K=(458.654, 457.296, 367.215, 248.375), w=752, h=480, n_per_kf=160,
                           noise_px=0.5, outlier_frac=0.03):
    """Synthetic KeyFrame messages: a camera translating along +x over a landmark cloud; landmark ids start at
    100 (landmark.cpp:3) and persist across the keyframes that see them; lm_3d is a noisy world position."""
    rng = np.random.default_rng(seed)
    fx, fy, cx, cy = K
    kfs = []
    pool = {}           # id -> world point
    next_id = 100
    for k in range(n_kf):
        t = np.array([-0.15 * k, 0.02 * np.sin(0.7 * k), 0.01 * k])
        aa = np.array([0.004 * k, -0.003 * k, 0.002 * k])
        q = np.concatenate([0.5 * aa, [1.0]]); q /= np.linalg.norm(q)
        pose = np.concatenate([q, t])
        R = np.array(ba_ref.edge(pose, np.zeros(3), np.zeros(2), K)[1])  # unused; keep ba_ref import exercised
        # visible existing landmarks
        ids, uv, p3 = [], [], []
        for i, X in pool.items():
            r, _, _ = ba_ref.edge(pose, X, np.zeros(2), K, jac=False)
            u = -r            # r = 0 - proj
            if 20 < u[0] < w - 20 and 20 < u[1] < h - 20 and len(ids) < n_per_kf - 30:
                ids.append(i); uv.append(u); p3.append(X)
        while len(ids) < n_per_kf:
            z = rng.uniform(4, 12)
            uu = np.array([rng.uniform(30, w - 30), rng.uniform(30, h - 30)])
            Xc = np.array([(uu[0] - cx) / fx * z, (uu[1] - cy) / fy * z, z])
            # world point: X = R^T (Xc - t); use the oracle's rotation through a tiny helper
            x, y, zq, wq = q
            Rm = np.array([[1 - 2 * (y * y + zq * zq), 2 * (x * y - zq * wq), 2 * (x * zq + y * wq)],
                           [2 * (x * y + zq * wq), 1 - 2 * (x * x + zq * zq), 2 * (y * zq - x * wq)],
                           [2 * (x * zq - y * wq), 2 * (y * zq + x * wq), 1 - 2 * (x * x + y * y)]])
            Xw = Rm.T @ (Xc - t)
            pool[next_id] = Xw
            ids.append(next_id); uv.append(uu); p3.append(Xw); next_id += 1
        uv = np.array(uv) + rng.normal(0, noise_px, (len(ids), 2))
        bad = rng.uniform(size=len(ids)) < outlier_frac
        uv[bad] += rng.uniform(-25, 25, (int(bad.sum()), 2))
        p3 = np.array(p3) + rng.normal(0, 0.03, (len(ids), 3))
        noisy_pose = pose.copy(); noisy_pose[4:] += rng.normal(0, 0.01, 3)
        kfs.append({"frame_id": 1000 + 3 * k, "lm_id": np.array(ids, np.int64), "lm_2d": uv, "lm_3d": p3,
                    "T_c_w": noisy_pose})
    return kfs
