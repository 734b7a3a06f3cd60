"""Independent numpy/scipy fp64 restatement of the reference's local bundle adjustment -- TEST INFRASTRUCTURE (oracle).

Purpose: a second opinion on oracle/ba_ref.c AND on the CUDA kernel that shares NO code or structure with either
(SURVEY.md section 4 item 1, 8(c)): the Jacobians are built by rotation-matrix algebra from the definition of the
residual (not g2o's hand-expanded 2x6 table), the whole Hessian is formed densely as J^T W J over ALL variables and
solved in one scipy Cholesky -- no Schur complement, no block structure.  Eliminating landmarks by Schur complement is
exact algebra, so the increments agree with g2o's BlockSolver_6_3 to rounding (~1e-9 relative) and the per-iteration
chi2 / lambda / rho traces can be compared line by line.

Follows (for WHAT is computed, not how):
  residual        /root/reference/3rdPartLib/g2o/g2o/types/sba/types_six_dof_expmap.h:209-214, .cpp:427-433
  pose update     types_six_dof_expmap.h:98-101  (pose <- exp(delta) * pose, delta = (omega, upsilon)), se3quat.h:218-260
  robust kernel   core/robust_kernel_impl.cpp:65-78 (Huber, delta = 1): rho(e) = e | 2 d sqrt(e) - d^2, rho' = 1 | d/sqrt(e)
  quadratic form  core/base_binary_edge.hpp:62-134: H += rho' J^T J, b += -rho' J^T r  (second-order term dropped)
  LM control      core/optimization_algorithm_levenberg.cpp:58-175
  driver          /root/reference/src/backend/vo_localmap.cpp:292-319 (12 it, cull chi2 > 3, 8 it)
PARITY UNPINNED against g2o itself (it cannot be built here: no Eigen3 / CHOLMOD in the image).
"""
import math

import numpy as np
import scipy.linalg


def quat_to_R(q):
    """q = (x, y, z, w) unit quaternion -> rotation matrix."""
    x, y, z, w = q
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def R_to_quat(R):
    """rotation matrix -> (x, y, z, w), w >= 0 branch selection by the largest diagonal term (any valid root works)."""
    t = np.trace(R)
    if t > 0:
        s = math.sqrt(t + 1.0) * 2
        q = np.array([(R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s, 0.25 * s])
    else:
        i = int(np.argmax(np.diag(R))); j = (i + 1) % 3; k = (i + 2) % 3
        s = math.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        q = np.zeros(4)
        q[i] = 0.25 * s; q[3] = (R[k, j] - R[j, k]) / s; q[j] = (R[j, i] + R[i, j]) / s; q[k] = (R[k, i] + R[i, k]) / s
    return q / np.linalg.norm(q)


def skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0.0]])


def se3_exp(d):
    """(omega, upsilon) -> (R, t) by the closed-form Rodrigues series."""
    om, up = np.asarray(d[:3], float), np.asarray(d[3:], float)
    th = np.linalg.norm(om)
    W = skew(om); W2 = W @ W
    if th < 1e-5:
        R = np.eye(3) + W + 0.5 * W2
        V = np.eye(3) + 0.5 * W + W2 / 6.0
    else:
        R = np.eye(3) + math.sin(th) / th * W + (1 - math.cos(th)) / th ** 2 * W2
        V = np.eye(3) + (1 - math.cos(th)) / th ** 2 * W + (th - math.sin(th)) / th ** 3 * W2
    return R, V @ up


def pose_oplus(pose, d):
    """pose7 = (qx qy qz qw tx ty tz) of T_c_w;  T <- exp(d) * T."""
    Rd, td = se3_exp(d)
    R = quat_to_R(pose[:4]); t = pose[4:]
    return np.concatenate([R_to_quat(Rd @ R), Rd @ t + td])


def project(pose, X, K):
    fx, fy, cx, cy = K
    Xc = quat_to_R(pose[:4]) @ X + pose[4:]
    return np.array([fx * Xc[0] / Xc[2] + cx, fy * Xc[1] / Xc[2] + cy]), Xc


def edge_jacobians(pose, X, uv, K):
    """r = uv - proj;  d r / d X (2x3) and d r / d delta (2x6) for T <- exp(delta) T at delta = 0, by the chain rule:
    Xc(delta) = exp(delta)(Xc) ~ Xc + omega x Xc + upsilon  =>  d Xc / d delta = [ -skew(Xc) | I ]."""
    fx, fy, _, _ = K
    uvp, Xc = project(pose, X, K)
    x, y, z = Xc
    dproj = np.array([[fx / z, 0, -fx * x / (z * z)], [0, fy / z, -fy * y / (z * z)]])      # d proj / d Xc
    R = quat_to_R(pose[:4])
    A = -dproj @ R                                       # d r / d X
    B = -dproj @ np.hstack([-skew(Xc), np.eye(3)])       # d r / d (omega, upsilon)
    return uv - uvp, A, B


class DenseBA:
    """poses (P,7), lms (L,3), ep/el (E,), uv (E,2), active (E,) u8 -- same array form as oracle.ba_ref.BAData."""

    def __init__(self, poses, lms, ep, el, uv, K, fixed_pose=0, fix_landmarks=0, active=None, huber_delta=1.0):
        self.poses = np.array(poses, float).reshape(-1, 7); self.lms = np.array(lms, float).reshape(-1, 3)
        self.ep = np.asarray(ep, int); self.el = np.asarray(el, int); self.uv = np.array(uv, float).reshape(-1, 2)
        self.K = tuple(float(v) for v in K)
        self.fixed_pose, self.fix_landmarks, self.delta = int(fixed_pose), int(fix_landmarks), float(huber_delta)
        self.active = np.ones(len(self.ep), np.uint8) if active is None else np.array(active, np.uint8)
        self.trace = []                          # per LM iteration: (chi2, lambda, rho, trials)

    # -- variable layout: free poses that have an active edge (by index), then landmarks with an active edge (by index)
    def _index(self):
        act = self.active.astype(bool)
        pset = sorted(set(self.ep[act].tolist()) - {self.fixed_pose})
        lset = [] if self.fix_landmarks else sorted(set(self.el[act].tolist()))
        self.pcol = {p: 6 * i for i, p in enumerate(pset)}
        self.lcol = {l: 6 * len(pset) + 3 * i for i, l in enumerate(lset)}
        self.nvar = 6 * len(pset) + 3 * len(lset)

    def chi2(self):
        """activeRobustChi2: sum of rho(e) over the active edges."""
        s = 0.0
        for e in np.nonzero(self.active)[0]:
            uvp, _ = project(self.poses[self.ep[e]], self.lms[self.el[e]], self.K)
            r = self.uv[e] - uvp
            c = float(r @ r)
            s += c if c <= self.delta ** 2 else 2 * self.delta * math.sqrt(c) - self.delta ** 2
        return s

    def _normal_equations(self):
        H = np.zeros((self.nvar, self.nvar)); b = np.zeros(self.nvar)
        for e in np.nonzero(self.active)[0]:
            p, l = int(self.ep[e]), int(self.el[e])
            r, A, B = edge_jacobians(self.poses[p], self.lms[l], self.uv[e], self.K)
            c = float(r @ r)
            w = 1.0 if c <= self.delta ** 2 else self.delta / math.sqrt(c)
            cols, J = [], []
            if p in self.pcol:
                cols += list(range(self.pcol[p], self.pcol[p] + 6)); J.append(B)
            if l in self.lcol:
                cols += list(range(self.lcol[l], self.lcol[l] + 3)); J.append(A)
            if not cols:
                continue
            J = np.hstack(J)
            H[np.ix_(cols, cols)] += w * J.T @ J
            b[cols] += -w * (J.T @ r)
        return H, b

    def _apply(self, x):
        for p, c in self.pcol.items():
            self.poses[p] = pose_oplus(self.poses[p], x[c:c + 6])
        for l, c in self.lcol.items():
            self.lms[l] = self.lms[l] + x[c:c + 3]

    def lm(self, iters):
        """OptimizationAlgorithmLevenberg over `iters` iterations; returns iterations run."""
        self._index()
        lam, ni, done = 0.0, 2.0, 0
        for it in range(iters):
            cur = self.chi2()
            H, b = self._normal_equations()
            if it == 0:
                lam = 1e-5 * float(np.abs(np.diag(H)).max()); ni = 2.0
            rho, q = 0.0, 0
            while True:
                bk_p, bk_l = self.poses.copy(), self.lms.copy()
                ok = True
                try:
                    x = scipy.linalg.cho_solve(scipy.linalg.cho_factor(H + lam * np.eye(self.nvar)), b)
                    self._apply(x)
                    tmp = self.chi2()
                except np.linalg.LinAlgError:
                    ok = False; tmp = float("inf"); x = np.zeros(self.nvar)
                scale = float(x @ (lam * x + b)) + 1e-3 if ok else 1e-3
                rho = (cur - tmp) / scale
                if rho > 0 and math.isfinite(tmp):
                    lam *= max(1.0 / 3.0, min(1.0 - (2 * rho - 1) ** 3, 2.0 / 3.0)); ni = 2.0; cur = tmp
                else:
                    lam *= ni; ni *= 2
                    self.poses, self.lms = bk_p, bk_l
                    if not math.isfinite(lam):
                        break
                q += 1
                if not (rho < 0 and q < 10):
                    break
            done += 1
            self.trace.append((cur, lam, rho, q))
            if q == 10 or rho == 0 or not math.isfinite(lam):
                break
        self.lam = lam
        return done

    def optimize(self, iters1=12, iters2=8, cull_chi2=3.0, min_edges_after_cull=0):
        """-> dict(iterations_run, n_culled, ok, chi2_initial, chi2_after1, chi2_final)."""
        out = dict(ok=1, n_culled=0, chi2_initial=self.chi2())
        out["iterations_run"] = self.lm(iters1)
        out["chi2_after1"] = self.chi2()
        for e in np.nonzero(self.active)[0]:
            uvp, _ = project(self.poses[self.ep[e]], self.lms[self.el[e]], self.K)
            r = self.uv[e] - uvp
            if float(r @ r) > cull_chi2:
                self.active[e] = 0; out["n_culled"] += 1
        if int(self.active.sum()) < min_edges_after_cull:
            out["ok"] = 0
        else:
            out["iterations_run"] += self.lm(iters2)
        out["chi2_final"] = self.chi2()
        return out
