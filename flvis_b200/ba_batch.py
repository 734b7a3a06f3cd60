"""Batched array form of BA problems for flv_ba_optimize (host and device-resident variants): bench / test support
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


class DeviceBatch:
    """Device-resident copy of a Batch + pinned host mirrors; runs the BA share of a bench step.

    Every step solves the windows of the streams whose keyframe falls on it: stream s is a keyframe on step i
    iff (i + s) % kf_every == 0.  The selected windows are reset to their initial (perturbed) state first so
    every solve does the full 12 + cull + 8 iterations' worth of work."""

    def __init__(self, ctx, batch, dev):
        import torch
        self.torch, self.ctx, self.b, self.dev = torch, ctx, batch, dev
        self.S = len(batch.problems)
        ctx._chk(ctx.lib.flv_ba_reserve(ctx.h, batch.MP, batch.ML, batch.ME))
        self.h0 = {k: torch.from_numpy(getattr(batch, k)).pin_memory() for k in ("poses", "lms", "ep", "el", "uv", "active")}
        self.d0 = {k: v.to(dev) for k, v in self.h0.items()}            # pristine inputs (HBM-resident)
        self.prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
        self.groups = None
        self.streams = None          # one CUDA stream per keyframe phase: the local-map thread analogue

    def _groups(self, kf_every):
        if self.groups is None:
            torch = self.torch
            self.groups = []
            nstat = C.sizeof(capi.BAStats)
            slot0 = 0
            for r in range(kf_every):
                sel = [s for s in range(self.S) if (r + s) % kf_every == 0]
                if not sel:
                    self.groups.append(None)
                    continue
                cp = (capi.BAProblem * len(sel))(*[self.b.cprob[s] for s in sel])
                h_prob = torch.frombuffer(bytearray(bytes(cp)), dtype=torch.uint8).pin_memory()
                # one packed blob per group: [poses | lms | stats | uv | ep | el | active] so that a (re)load of the windows
                # is ONE copy and the read-back of the results ONE copy (they sit in front of / behind every solve)
                isel = torch.tensor(sel)
                parts = {k: self.h0[k][isel].contiguous() for k in ("poses", "lms", "uv", "ep", "el", "active")}
                order = ["poses", "lms", "stats", "uv", "ep", "el", "active"]
                sizes = {k: v.numel() * v.element_size() for k, v in parts.items()}
                sizes["stats"] = len(sel) * nstat
                off, o = {}, 0
                for k in order:
                    off[k] = o
                    o = (o + sizes[k] + 63) & ~63
                total = o
                h_blob = torch.zeros(total, dtype=torch.uint8).pin_memory()
                for k, v in parts.items():
                    h_blob[off[k]:off[k] + sizes[k]] = v.reshape(-1).view(torch.uint8)
                d_src = h_blob.to(self.dev)                       # pristine inputs, HBM-resident
                d_blob = torch.zeros(total, dtype=torch.uint8, device=self.dev)
                h_res = torch.zeros(off["uv"], dtype=torch.uint8).pin_memory()      # poses | lms | stats
                ptr = {k: d_blob.data_ptr() + off[k] for k in order}
                g = {"sel": sel, "d_prob": h_prob.to(self.dev), "prm": capi.BAParams(12, 8, 1.0, 3.0, 0, slot0),
                     "h_blob": h_blob, "d_src": d_src, "d_blob": d_blob, "h_res": h_res, "ptr": ptr, "res_bytes": off["uv"]}
                slot0 += len(sel)      # concurrent launches of different phases use disjoint workspace slots
                self.groups.append(g)
        return self.groups

    def _n_per_step(self, kf_every):
        return -(-self.S // kf_every)

    def h2d_bytes_per_step(self, kf_every):
        return self._n_per_step(kf_every) * (self.b.MP * 56 + self.b.ML * 24 + self.b.ME * (16 + 8 + 1))

    def d2h_bytes_per_step(self, kf_every):
        return self._n_per_step(kf_every) * (self.b.MP * 56 + self.b.ML * 24 + C.sizeof(capi.BAStats))

    def step(self, i, mode, kf_every):
        g = self._groups(kf_every)[i % kf_every]
        if g is None:
            return
        torch = self.torch
        if self.streams is None:
            self.streams = [torch.cuda.Stream(self.dev) for _ in range(kf_every)]
        st = self.streams[i % kf_every]
        # like FLVIS's local-map nodelet, the BA of a keyframe runs concurrently with the tracking of the following
        # frames: its own stream, nothing in the frame loop waits for it (the reference's feedback path is dead code)
        with torch.cuda.stream(st):
            self.ctx.set_ba_stream(st.cuda_stream, True)
            self._solve(g, mode)
            self.ctx.set_ba_stream(0, False)

    def join(self, main_stream):
        """Make `main_stream` wait for all outstanding BA work (end of a timed region)."""
        if self.streams:
            for st in self.streams:
                main_stream.wait_stream(st)

    def _solve(self, g, mode):
        g["d_blob"].copy_(g["h_blob"] if mode == "host" else g["d_src"], non_blocking=True)   # (re)load: H2D in e2e mode, D2D otherwise
        n = len(g["sel"])
        p = g["ptr"]
        self.ctx._chk(self.ctx.lib.flv_ba_optimize(
            self.ctx.h, n, C.cast(C.c_void_p(g["d_prob"].data_ptr()), C.POINTER(capi.BAProblem)), C.byref(g["prm"]),
            C.c_void_p(p["poses"]), C.c_void_p(p["lms"]), C.c_void_p(p["ep"]), C.c_void_p(p["el"]), C.c_void_p(p["uv"]),
            C.c_void_p(p["active"]), C.cast(C.c_void_p(p["stats"]), C.POINTER(capi.BAStats)), capi.MEM_DEVICE))
        if mode == "host":
            g["h_res"].copy_(g["d_blob"][:g["res_bytes"]], non_blocking=True)                 # poses | landmarks | stats
