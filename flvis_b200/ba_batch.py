"""Batched array form of BA problems for flv_ba_optimize (the layout the C ABI takes): bench / test support
around the C ABI.  The problems themselves come from synthdata/ba_problems.py; nothing here touches oracle/."""
import ctypes as C

import numpy as np

from . import capi


class Batch:
    """S problems padded to common strides (the layout flv_ba_optimize takes)."""

    def __init__(self, problems):
        self.problems = problems
        S = len(problems)
        self.MP = max(len(p.poses) for p in problems)
        self.ML = max(len(p.lms) for p in problems)
        self.ME = max(len(p.ep) for p in problems)
        self.poses = np.zeros((S, self.MP, 7)); self.poses[:, :, 3] = 1.0
        self.lms = np.zeros((S, self.ML, 3))
        self.ep = np.zeros((S, self.ME), np.int32); self.el = np.zeros((S, self.ME), np.int32)
        self.uv = np.zeros((S, self.ME, 2)); self.active = np.zeros((S, self.ME), np.uint8)
        self.cprob = (capi.BAProblem * S)()
        for s, p in enumerate(problems):
            self.poses[s, :len(p.poses)] = p.poses; self.lms[s, :len(p.lms)] = p.lms
            E = len(p.ep)
            self.ep[s, :E] = p.ep; self.el[s, :E] = p.el; self.uv[s, :E] = p.uv; self.active[s, :E] = 1
            self.cprob[s] = capi.BAProblem(len(p.poses), len(p.lms), E, p.fixed_pose, p.fix_landmarks, *p.K)


def make_batch(n_streams, window, n_landmarks, obs_per_frame, seed):
    from synthdata import ba_problems
    return Batch([ba_problems.make_problem(window, n_landmarks, obs_per_frame, seed=1000 * seed + s) for s in range(n_streams)])


def solve_batch_host(ctx, batch, prm=None):
    """Run flv_ba_optimize with HOST arrays (the drop-in call). Returns (poses, lms, active, stats list)."""
    prm = prm or capi.BAParams(12, 8, 1.0, 3.0, 0)
    S = len(batch.problems)
    ctx._chk(ctx.lib.flv_ba_reserve(ctx.h, batch.MP, batch.ML, batch.ME))
    poses, lms, active = batch.poses.copy(), batch.lms.copy(), batch.active.copy()
    stats = (capi.BAStats * S)()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    ctx._chk(ctx.lib.flv_ba_optimize(ctx.h, S, batch.cprob, C.byref(prm), vp(poses), vp(lms), vp(batch.ep),
                                     vp(batch.el), vp(batch.uv), vp(active), stats, capi.MEM_HOST))
    return poses, lms, active, list(stats)
