"""Batched device-resident frame loop used by bench.py and the multi-GPU driver.

Order of stages per stereo frame = the order F2FTracking::image_feed runs them
(/root/reference/src/frontend/f2f_tracking.cpp:187-355): LK frame->frame (:227) -> redetect (:291) ->
depthInnovation's left->right LK (:326 -> camera_frame.cpp:124-128); every KF_EVERY-th frame of a stream
is a keyframe and triggers that stream's local BA (vo_localmap.cpp:292-319).  All S streams share one
launch per stage.  torch is plumbing here: device buffers, pinned host buffers, events.
"""
import ctypes as C

import numpy as np

from . import capi


def make_ba_batch(n_streams, window, seed=0, n_landmarks=1500, obs_per_frame=480):
    """Synthetic local-BA windows (ba_demo pattern, SURVEY.md 8(d) C1: W=10, E=4800, L~1500)."""
    from . import ba_batch
    return ba_batch.make_batch(n_streams, window, n_landmarks, obs_per_frame, seed)


class FrontendBench:
    def __init__(self, n_streams, w, h, max_pts, npts, feature_para, device_index, ba_window=10, kf_every=5, seed=0,
                 stereo=True, ba_landmarks=1500):
        import torch
        self.torch = torch
        self.S, self.w, self.h, self.max_pts, self.npts = n_streams, w, h, max_pts, npts
        self.stereo = stereo            # False: depth camera (DEPTH_D435): no right image, no left->right LK
        self.dev = torch.device("cuda", device_index)
        self.ctx = capi.Context(n_streams, w, h, max_pts, device=device_index)
        # a dedicated (non-default) torch stream: the library launches on it and all events are recorded on it
        self.stream = torch.cuda.Stream(self.dev)
        self.side = torch.cuda.Stream(self.dev)
        self.res_stream = torch.cuda.Stream(self.dev)     # result read-back (e2e mode) off the compute stream
        self.ev_step = torch.cuda.Event()
        self.ev_sel, self.ev_red = torch.cuda.Event(), torch.cuda.Event()
        self.ctx.set_stream(self.stream.cuda_stream)
        self.fp = capi.FeatureParams(int(feature_para[0]), int(feature_para[1]), int(feature_para[2] // 2),
                                     int(feature_para[3]), float(feature_para[4]), int(feature_para[5]))
        S, M = n_streams, max_pts
        z = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype, device=self.dev)
        self.d_npts = z(S, dtype=torch.int32)
        self.d_pts = z(S, M, 2, dtype=torch.float32)        # current feature positions in prev0
        self.d_next = z(S, M, 2, dtype=torch.float32)
        self.d_status = z(S, M, dtype=torch.uint8)
        self.d_err = z(S, M, dtype=torch.float32)
        self.d_keep = z(S, M, dtype=torch.uint8)
        self.d_cur = z(S, M, 2, dtype=torch.float32)        # positions in cur0 after the keep rule
        self.d_cur64 = z(S, M, 2, dtype=torch.float64)
        self.d_right = z(S, M, 2, dtype=torch.float32)
        self.d_rstatus = z(S, M, dtype=torch.uint8)
        self.d_rerr = z(S, M, dtype=torch.float32)
        self.d_new = z(S, M, 2, dtype=torch.float32)
        self.d_nnew = z(S, dtype=torch.int32)
        # pinned result buffers for the e2e mode
        pin = lambda *shape, dtype: torch.zeros(*shape, dtype=dtype).pin_memory()
        # (double-buffered: the host reads the results of frame k-1 while frame k computes)
        self.h_out = [dict(cur=pin(S, M, 2, dtype=torch.float32), keep=pin(S, M, dtype=torch.uint8),
                           right=pin(S, M, 2, dtype=torch.float32), rstatus=pin(S, M, dtype=torch.uint8),
                           new=pin(S, M, 2, dtype=torch.float32), nnew=pin(S, dtype=torch.int32)) for _ in range(2)]
        self.done = [torch.cuda.Event(), torch.cuda.Event()]
        self.prefetched = -1            # frame-pool index whose host upload is already in flight
        self.pending = None             # index of the result buffer the host has not consumed yet
        self.slots = [0, 1, 2]          # prev0, cur0, cur1
        self.kf_every = kf_every
        self.collect_lk = False
        self.lk_events = []
        self.lk_ms = 0.0
        self.has_ba = False
        self.ba = None
        try:
            import os
            if os.environ.get("FLV_BENCH_NO_BA"):          # diagnostic only: frontend alone under bench conditions
                raise ImportError
            from . import ba_batch
            self.ba = ba_batch.DeviceBatch(self.ctx, make_ba_batch(n_streams, ba_window, seed=seed, n_landmarks=ba_landmarks), self.dev)
            self.has_ba = True
        except ImportError:
            self.ba = None
        self.h2d_bytes_per_step = (2 if stereo else 1) * S * w * h + (self.ba.h2d_bytes_per_step(kf_every) if self.has_ba else 0)
        self.d2h_bytes_per_step = (S * M * (8 + 1) * 2 + S * M * 8 + S * 4 +
                                   (self.ba.d2h_bytes_per_step(kf_every) if self.has_ba else 0))

    # frame pool -------------------------------------------------------------------------------
    def load_pool(self, f0, f1):
        torch = self.torch
        self.n_pool = f0.shape[0]
        self.h_pool0 = torch.from_numpy(f0).pin_memory()
        self.h_pool1 = torch.from_numpy(f1).pin_memory()
        self.d_pool0 = self.h_pool0.to(self.dev)
        self.d_pool1 = self.h_pool1.to(self.dev)

    def reset(self):
        """Frame 0: upload, pyramid, FeatureDEM::detect => the tracked set (init_frame, f2f_tracking.cpp:402-453)."""
        ctx, S = self.ctx, self.S
        self.torch.cuda.synchronize()
        self.prefetched, self.pending = -1, None
        prev0 = self.slots[0]
        ctx.upload_dev(prev0, S, self.d_pool0[0].data_ptr())
        ctx.build_pyramid(prev0, S)
        ctx.feature_detect_dev(prev0, S, self.fp, self.d_pts.data_ptr(), self.d_npts.data_ptr())
        self.torch.cuda.synchronize()

    def _lk(self, src, dst, d_prev, d_init, d_out, d_st, d_err, max_level):
        if self.collect_lk:
            e0 = self.torch.cuda.Event(enable_timing=True); e1 = self.torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
        self.ctx.lk_track_dev(src, dst, self.S, self.d_npts.data_ptr(), d_prev.data_ptr(), d_init.data_ptr(),
                              d_out.data_ptr(), d_st.data_ptr(), d_err.data_ptr(), max_level=max_level)
        if self.collect_lk:
            e1.record(self.stream)
            self.lk_events.append((e0, e1))

    def finish_lk_timing(self):
        self.torch.cuda.synchronize()
        ms = sum(a.elapsed_time(b) for a, b in self.lk_events)
        n = len(self.lk_events)
        self.lk_events = []
        self.collect_lk = False
        return ms, n

    def join(self):
        if self.has_ba:
            self.ba.join(self.stream)
        if self.pending is not None:                       # last frame's results
            self.done[self.pending].synchronize()
            self.pending = None

    def step(self, i, mode):
        with self.torch.cuda.stream(self.stream):
            self._step(i, mode)

    def _step(self, i, mode):
        ctx, S = self.ctx, self.S
        prev0, cur0, cur1 = self.slots
        k = (i + 1) % self.n_pool
        if mode == "host":
            if self.prefetched != k:                       # first host-mode step: nothing in flight yet
                ctx.upload_host_async(cur0, S, self.h_pool0[k].data_ptr())
                if self.stereo:
                    ctx.upload_host_async(cur1, S, self.h_pool1[k].data_ptr())
        else:
            self.prefetched = -1
            ctx.upload_dev(cur0, S, self.d_pool0[k].data_ptr())
            if self.stereo:
                ctx.upload_dev(cur1, S, self.d_pool1[k].data_ptr())
        ctx.build_pyramid(cur0, S)
        # Shi-Tomasi of the new left image only needs the image: start it now on the library's auxiliary stream, it
        # overlaps the right pyramid, the frame->frame LK and the keep rule (results identical, see flv_feature_prepare)
        ctx.feature_prepare(cur0, S, self.fp, redetect=True)
        if self.stereo:
            ctx.build_pyramid(cur1, S)
        # frame -> frame LK (lkorb_tracking.cpp:64-73) + keep rule (:98-119)
        self._lk(prev0, cur0, self.d_pts, self.d_pts, self.d_next, self.d_status, self.d_err, 10)
        if mode == "host" and self.pending is not None:
            self.stream.wait_event(self.done[self.pending])      # the previous frame's read-back precedes the overwrite
        ctx.select_tracked_dev(S, self.d_npts.data_ptr(), self.d_pts.data_ptr(), self.d_next.data_ptr(),
                               self.d_status.data_ptr(), self.d_keep.data_ptr(), self.d_cur.data_ptr(),
                               self.d_cur64.data_ptr())
        # redetect on cur0 against the surviving features (f2f_tracking.cpp:291 -> feature_dem.cpp:124): the region
        # selection is a one-CTA-per-stream kernel, so it runs on a side stream next to the left -> right LK of the
        # surviving features (camera_frame.cpp:124-128, maxLevel 5, initial flow = cam0 position); both only read cur0
        self.ev_sel.record(self.stream)
        self.side.wait_event(self.ev_sel)
        ctx.set_stream(self.side.cuda_stream)
        ctx.feature_redetect_dev(cur0, S, self.fp, self.d_cur64.data_ptr(), self.d_npts.data_ptr(),
                                 self.d_new.data_ptr(), self.d_nnew.data_ptr())
        self.ev_red.record(self.side)
        ctx.set_stream(self.stream.cuda_stream)
        if self.stereo:
            self._lk(cur0, cur1, self.d_cur, self.d_cur, self.d_right, self.d_rstatus, self.d_rerr, 5)
        self.stream.wait_event(self.ev_red)
        if self.has_ba:
            self.ba.step(i, mode, self.kf_every)
        if mode == "host":
            # read the frame's results back on their own stream: the compute stream goes straight on to the next frame
            # (its kernels that overwrite these buffers wait for this copy, see the wait before the keep rule)
            o = self.h_out[i & 1]
            self.ev_step.record(self.stream)
            self.res_stream.wait_event(self.ev_step)
            with self.torch.cuda.stream(self.res_stream):
                o["cur"].copy_(self.d_cur, non_blocking=True); o["keep"].copy_(self.d_keep, non_blocking=True)
                o["right"].copy_(self.d_right, non_blocking=True); o["rstatus"].copy_(self.d_rstatus, non_blocking=True)
                o["new"].copy_(self.d_new, non_blocking=True); o["nnew"].copy_(self.d_nnew, non_blocking=True)
                self.done[i & 1].record(self.res_stream)
            # submit the NEXT frame's images now (library copy stream: the H2D overlaps this frame's kernels), then
            # consume the PREVIOUS frame's results -- every frame's inputs and outputs cross PCIe, one frame of latency
            k2 = (i + 2) % self.n_pool
            ctx.upload_host_async(prev0, S, self.h_pool0[k2].data_ptr())      # next step's cur0 slot
            if self.stereo:
                ctx.upload_host_async(cur1, S, self.h_pool1[k2].data_ptr())
            self.prefetched = k2
            if self.pending is not None:
                self.done[self.pending].synchronize()
            self.pending = i & 1
        # the tracked set of the next frame lives in cur0
        self.d_pts, self.d_cur = self.d_cur, self.d_pts
        self.slots = [cur0, prev0, cur1]
