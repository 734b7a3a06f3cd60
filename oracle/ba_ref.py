"""ctypes wrapper of oracle/ba_ref.c (C restatement of the g2o LM/Schur path) -- TEST INFRASTRUCTURE.

PARITY UNPINNED (see ba_ref.c header): validated by known-answer tests only.
"""
import ctypes as C
import numpy as np

from .feature_dem_ref import helpers_lib


class Problem(C.Structure):
    _fields_ = [("n_poses", C.c_int), ("n_landmarks", C.c_int), ("n_edges", C.c_int), ("fixed_pose", C.c_int),
                ("fix_landmarks", C.c_int), ("fx", C.c_double), ("fy", C.c_double), ("cx", C.c_double), ("cy", C.c_double)]


class Stats(C.Structure):
    _fields_ = [("iterations_run", C.c_int), ("n_culled", C.c_int), ("ok", C.c_int), ("reserved", C.c_int),
                ("chi2_initial", C.c_double), ("chi2_after1", C.c_double), ("chi2_final", C.c_double),
                ("lambda_final", C.c_double)]


class BAData:
    """One BA problem in array form (same layout the C ABI takes for one stream)."""

    def __init__(self, poses, lms, ep, el, uv, K, fixed_pose=0, fix_landmarks=0, active=None):
        self.poses = np.ascontiguousarray(poses, np.float64).reshape(-1, 7)
        self.lms = np.ascontiguousarray(lms, np.float64).reshape(-1, 3)
        self.ep = np.ascontiguousarray(ep, np.int32)
        self.el = np.ascontiguousarray(el, np.int32)
        self.uv = np.ascontiguousarray(uv, np.float64).reshape(-1, 2)
        self.K = tuple(float(v) for v in K)
        self.fixed_pose, self.fix_landmarks = int(fixed_pose), int(fix_landmarks)
        self.active = np.ones(len(self.ep), np.uint8) if active is None else np.ascontiguousarray(active, np.uint8)

    def copy(self):
        return BAData(self.poses.copy(), self.lms.copy(), self.ep, self.el, self.uv, self.K, self.fixed_pose,
                      self.fix_landmarks, self.active.copy())

    def c_problem(self):
        return Problem(len(self.poses), len(self.lms), len(self.ep), self.fixed_pose, self.fix_landmarks, *self.K)


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def optimize(d, iters1=12, iters2=8, huber_delta=1.0, cull_chi2=3.0, min_edges_after_cull=0, trace=None):
    """In-place local BA on BAData `d` (12 it, cull chi2>3, 8 it by default). Returns Stats.
    trace: optional list that receives one (chi2, lambda, rho, trials) tuple per LM iteration (not thread-safe)."""
    lib = helpers_lib()
    tbuf = None
    if trace is not None:
        tbuf = np.zeros((iters1 + iters2 + 2, 4))
        lib.oracle_ba_set_trace.argtypes = [C.c_void_p, C.c_int]
        lib.oracle_ba_set_trace(_p(tbuf), len(tbuf))
    lib.oracle_ba_optimize.argtypes = [C.POINTER(Problem), C.c_int, C.c_int, C.c_double, C.c_double, C.c_int,
                                       C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.POINTER(Stats)]
    pb = d.c_problem()
    st = Stats()
    lib.oracle_ba_optimize(C.byref(pb), iters1, iters2, huber_delta, cull_chi2, min_edges_after_cull, _p(d.poses),
                           _p(d.lms), _p(d.ep), _p(d.el), _p(d.uv), _p(d.active), C.byref(st))
    if trace is not None:
        n = lib.oracle_ba_trace_rows()
        lib.oracle_ba_set_trace(None, 0)
        trace.extend(tuple(r) for r in tbuf[:n])
    return st


def edge(pose, X, uv, K, jac=True):
    lib = helpers_lib()
    lib.oracle_ba_edge.argtypes = [C.c_void_p] * 3 + [C.c_double] * 4 + [C.c_void_p] * 3
    pose = np.ascontiguousarray(pose, np.float64); X = np.ascontiguousarray(X, np.float64)
    uv = np.ascontiguousarray(uv, np.float64)
    r = np.zeros(2); A = np.zeros((2, 3)); B = np.zeros((2, 6))
    lib.oracle_ba_edge(_p(pose), _p(X), _p(uv), *K, _p(r), _p(A) if jac else None, _p(B) if jac else None)
    return r, A, B


def se3_oplus(pose, u):
    lib = helpers_lib()
    lib.oracle_se3_oplus.argtypes = [C.c_void_p, C.c_void_p]
    p = np.array(pose, np.float64); u = np.ascontiguousarray(u, np.float64)
    lib.oracle_se3_oplus(_p(p), _p(u))
    return p
