#!/usr/bin/env python
"""bench.py -- frames/sec of the FLVIS hot path on EuRoC-shaped 752x480 stereo streams (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--streams S] [--impl reference]

A "step" is one stereo frame for every one of the S concurrent streams on this GPU:
    pyramid(cur0), pyramid(cur1) -> LK frame->frame (prev0 -> cur0, 480 pts/stream)
    -> FeatureDEM redetect on cur0 (Shi-Tomasi + region select) -> LK left->right (cur0 -> cur1)
    -> local BA (10-KF window, 12 + cull + 8 LM iterations) for the streams whose keyframe falls on this step
       (one keyframe every KF_EVERY frames, phases staggered so every step does the same amount of work).
`value`  : device-timed, inputs already resident in HBM (a device-side frame pool), CUDA events.
`e2e`    : the same step driven through the C ABI with pinned HOST buffers: the two images per stream are
           copied H2D and the tracked points / status / new corners / BA poses are copied D2H every step.
Multi-GPU: streams are independent; every rank runs its own S streams ("weak" scaling),
no data-path collective; NCCL is used for the barrier and the max-over-ranks reduction of the timing only.
`--impl reference`: the reference's CPU path (cv2 = the OpenCV the reference links, + the C port of its g2o
path) on the host cores, bounded sample, same metric/config.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from flvis_b200 import sharding  # noqa: E402

W, H = 752, 480
NPTS = 480
MAX_PTS = 512
KF_EVERY = 5
BA_WINDOW = 10
BA_LANDMARKS = 1500
STEREO = True
WORKLOAD = "euroc"
WORKLOADS = {   # BASELINE.json configs: image size, stereo, local-BA window / landmarks
    "euroc": dict(w=752, h=480, stereo=True, window=10, landmarks=1500, name="EuRoC-shaped 752x480 stereo (configs[1]/[2])"),
    "kitti": dict(w=1241, h=376, stereo=True, window=20, landmarks=2000, name="KITTI-shaped 1241x376 stereo, 20-KF / 2k-landmark window (configs[3])"),
    "d435": dict(w=640, h=480, stereo=False, window=10, landmarks=1500, name="640x480 D435i depth (configs[4]); depth lookups are per-point reads, no right image"),
}
FEATURE_PARA = [30, 20, 5, 1000, 0.01, 10]       # launch/EuRoC_MAV/euroc.yaml:57-67
P_PYR = 479400                                    # pyramid pixels of 752x480
LK_BYTES_PER_CALL = 6 * P_PYR + 29 * NPTS         # algorithmic bytes of one LK call, one stream: both u8 pyramids + the
                                                  # 4 B/px Scharr pyramid of the first image + 29 B per point
LK_NCU_TRAFFIC = 95.5e6                           # dram read+write of one lk_track_kernel_v4 launch, 32 streams
                                                  # (profiles/r01_lk_v4_ncu_full.csv: 91.8 MB + 3.8 MB)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
# synthetic EuRoC-shaped stereo streams: per-stream textured canvas, integer-pixel camera motion,
# right image = left shifted by an integer disparity.  Frames are crops, so generation is cheap.
MOTION = [(3, 1), (2, -2), (-3, 2), (-2, -1)]     # cumulative motion is periodic => points stay in view


def make_streams(n_streams, seed0, n_frames):
    """Streams with global ids seed0 .. seed0+n_streams-1 (flvis_b200.sharding.stream_seed gives the texture seed)."""
    from synthdata import textures as synth
    frames0 = np.empty((n_frames, n_streams, H, W), np.uint8)
    frames1 = np.empty((n_frames, n_streams, H, W), np.uint8)
    for s in range(n_streams):
        canvas = synth.texture(sharding.stream_seed(seed0 + s), H + 64, W + 128, blur=2)
        ox, oy = 48, 32
        disp = 20 + (s % 7)
        for t in range(n_frames):
            frames0[t, s] = canvas[oy:oy + H, ox:ox + W]
            frames1[t, s] = canvas[oy:oy + H, ox - disp + 32:ox - disp + 32 + W]
            dx, dy = MOTION[t % len(MOTION)]
            ox += dx; oy += dy
    return frames0, frames1


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.stop_flag = threading.Event()
        self.rows = []

    def run(self):
        # NVML (pynvml) polls every few milliseconds -- the timed region is only tens of milliseconds long; nvidia-smi
        # (one process spawn per sample) is the fallback
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = [("hw_slowdown", pynvml.nvmlClocksThrottleReasonHwSlowdown),
                    ("hw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", pynvml.nvmlClocksThrottleReasonSwThermalSlowdown),
                    ("sw_power_cap", pynvml.nvmlClocksThrottleReasonSwPowerCap)]
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for _, b in bits])
                self.stop_flag.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                      str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows)
        reasons = []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(r[2 + i].lower().startswith("active") for r in self.rows):
                reasons.append(n)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(self.rows)}


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from flvis_b200 import capi
    from flvis_b200.pipeline import FrontendBench

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line (NCCL prints its version)
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    S = args.streams
    n_pool = 8                                        # distinct frames per stream in the pool (cycled)
    f0, f1 = make_streams(S, sharding.stream_ids(rank, world, S)[0], n_pool)     # rank r owns streams [r*S, (r+1)*S)
    bench = FrontendBench(S, W, H, MAX_PTS, NPTS, FEATURE_PARA, local_rank, ba_window=BA_WINDOW,
                          kf_every=KF_EVERY, seed=rank, stereo=STEREO, ba_landmarks=BA_LANDMARKS)
    bench.load_pool(f0, f1)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    host_ms = {}

    def timed(mode, steps, warmup, bench=bench):
        bench.reset()
        for i in range(warmup):
            bench.step(i, mode)
        bench.join()
        barrier()
        launches0 = bench.ctx.launches
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        bench.lk_ms = 0.0
        bench.collect_lk = True
        ev0.record(bench.stream)
        t_host = time.perf_counter()
        for i in range(steps):
            bench.step(warmup + i, mode)
        host_ms[mode] = (time.perf_counter() - t_host) * 1e3 / steps      # host time to SUBMIT one step (no sync in device mode)
        bench.join()                       # the timed region ends when the asynchronous BA streams have drained too
        ev1.record(bench.stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        lk_ms, lk_calls = bench.finish_lk_timing()
        ms = sharding.max_over_ranks(ms, dist if world > 1 else None, dev)        # job time = slowest rank
        return ms, bench.ctx.launches - launches0, lk_ms, lk_calls

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    ms_dev, launches, lk_ms, lk_calls = timed("device", args.steps, args.warmup)
    ms_e2e, _, _, _ = timed("host", args.steps, args.warmup)
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=2)
    # BASELINE configs[1]: ONE stream on the GPU (latency-oriented), reported next to the batched headline
    single = None
    if world == 1 and S > 1:
        b1 = FrontendBench(1, W, H, MAX_PTS, NPTS, FEATURE_PARA, local_rank, ba_window=BA_WINDOW, kf_every=KF_EVERY, seed=rank,
                           stereo=STEREO, ba_landmarks=BA_LANDMARKS)
        b1.load_pool(f0[:, :1].copy(), f1[:, :1].copy())
        k1 = max(args.steps, 50)
        ms1, _, _, _ = timed("device", k1, args.warmup, bench=b1)
        ms1h, _, _, _ = timed("host", k1, args.warmup, bench=b1)
        single = {"workload": "1 EuRoC-shaped 752x480 stereo stream, LK frontend + 10-KF local BA (BASELINE configs[1])",
                  "steps": k1, "value": k1 / (ms1 * 1e-3), "e2e": k1 / (ms1h * 1e-3), "unit": "frames/s",
                  "ms_per_frame": ms1 / k1}

    frames = args.steps * S * world
    peak, peak_src = load_peaks()
    out = None
    if rank == 0:
        lk_us = 1e3 * lk_ms / max(lk_calls, 1)
        achieved = LK_BYTES_PER_CALL * S / (lk_us * 1e-6) / 1e9 if lk_us > 0 else 0.0
        # bounded sample of the same workload on every host core: all S streams, 6 frames each
        cpu = cpu_baseline(min(S, os.cpu_count() or 1), 6, args) if world == 1 and not args.no_cpu else None
        if single is not None and not args.no_cpu:
            single["cpu_value"] = cpu_baseline(1, 10, args)["value"]      # same single stream on the host (cv2 uses all cores)
        out = {
            "metric": "frames/sec (device-timed), EuRoC-shaped 752x480 stereo, LK frontend + 10-KF local BA" if WORKLOAD == "euroc"
                      else f"frames/sec (device-timed), {WORKLOADS[WORKLOAD]['name']}",
            "value": frames / (ms_dev * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8/i32 fixed-point + f32 (frontend), f64 (BA)",
            "data": "synthetic",
            "config": {"workload": (f"{S} concurrent EuRoC-shaped 752x480 stereo streams per GPU (BASELINE configs[2]); " if WORKLOAD == "euroc"
                                    else f"{S} concurrent streams per GPU, {WORKLOADS[WORKLOAD]['name']}; ") +
                                   f"480 pts/stream, LK 31x31 4 levels x2, GFTT N=1000 + FeatureDEM redetect, "
                                   f"local BA W={BA_WINDOW} every {KF_EVERY}th frame" + ("" if bench.has_ba else " [BA NOT YET IN STEP]"),
                       "streams_per_gpu": S, "image": [W, H], "points_per_stream": NPTS,
                       "l2_note": "no explicit L2 flush: every step ingests two fresh S*361 KB image sets from a rotating "
                                  "frame pool and rewrites all pyramids, so no step reuses another step's cached inputs",
                       "e2e_pipeline": "H2D of frame k+1 (library copy stream) and the host's read of frame k-1's results "
                                       "overlap the kernels of frame k; every frame's inputs and outputs cross PCIe "
                                       "inside the timed region (pinned host buffers, one frame of result latency)",
                       "parallelism": f"streams sharded {S}/GPU x {world} GPU, no data-path collective",
                       "host_submit_ms_per_step": {k: round(v, 3) for k, v in host_ms.items()}},
            "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": "frames/s",
                    "h2d_bytes_per_step": bench.h2d_bytes_per_step, "d2h_bytes_per_step": bench.d2h_bytes_per_step,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches,
            "single_stream": single,
            "roofline": {"kernel": "lk_track_kernel_v4 (frame->frame + left->right)", "bound": "hbm",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": LK_NCU_TRAFFIC * S / 32 if WORKLOAD == "euroc" else None, "peak_source": peak_src,
                         "us_per_launch": lk_us, "algorithmic_bytes_per_launch": LK_BYTES_PER_CALL * S,
                         "note": "LK is instruction-issue bound, not HBM bound (ncu: DRAM < 3 % busy, traffic == algorithmic "
                                 "bytes, i.e. no re-reads); us_per_launch is measured inside the step, where other streams' "
                                 "kernels share the SMs (269 / 295 us alone); see DESIGN.md section 5"},
            "clocks": sampler.summary() if sampler else None,
        }
        if cpu:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return out


# ------------------------------------------------------------------------------------------------
def cpu_frame(cv2, fd_para, prev0, cur0, cur1, pts, crit):
    """One stereo frame of the reference's OpenCV stages for one stream (call sites in the module docstring)."""
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(prev0, cur0, pts, pts.copy(), winSize=(31, 31), maxLevel=10, criteria=crit,
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW)            # lkorb_tracking.cpp:64-73
    mask = np.full(cur0.shape, 255, np.uint8)
    cv2.goodFeaturesToTrack(cur0, fd_para[3], fd_para[4], fd_para[5], mask=mask)        # feature_dem.cpp:160
    if STEREO:
        cv2.calcOpticalFlowPyrLK(cur0, cur1, nxt, nxt.copy(), winSize=(31, 31), maxLevel=5, criteria=crit,
                                 flags=cv2.OPTFLOW_USE_INITIAL_FLOW)                     # camera_frame.cpp:124-128
    ok = st.ravel() == 1
    nxt[~ok] = pts[~ok]
    return nxt


def cpu_baseline(n_streams, n_frames, args):
    """Reference CPU path on this box's host cores, bounded sample: n_streams streams x n_frames frames."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    cores = os.cpu_count() or 1
    workers = min(n_streams, cores)
    cv2.setNumThreads(max(1, cores // workers))
    f0, f1 = make_streams(n_streams, 0, n_frames + 1)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001)
    pts0 = [cv2.goodFeaturesToTrack(f0[0, s], NPTS, 0.01, 10).reshape(-1, 2) for s in range(n_streams)]
    ba = None
    try:
        from flvis_b200.pipeline import make_ba_batch
        from oracle import ba_ref                    # the cpu_baseline leg: the one place bench.py executes oracle/
        ba = make_ba_batch(n_streams, BA_WINDOW, seed=7, n_landmarks=BA_LANDMARKS)

        def cpu_ba_solve(batch, s):
            p = batch.problems[s]
            ba_ref.optimize(ba_ref.BAData(p.poses.copy(), p.lms.copy(), p.ep, p.el, p.uv, p.K, p.fixed_pose, p.fix_landmarks), 12, 8)
    except Exception:
        ba = None

    def work(s):
        pts = pts0[s]
        for t in range(1, n_frames + 1):
            pts = cpu_frame(cv2, FEATURE_PARA, f0[t - 1, s], f0[t, s], f1[t, s], pts, crit)
            if ba is not None and (t + s) % KF_EVERY == 0:
                cpu_ba_solve(ba, s)
        return 0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(work, range(n_streams)))
    dt = time.perf_counter() - t0
    return {"value": n_streams * n_frames / dt, "unit": "frames/s", "cores": cores,
            "kind": "port", "threads": workers * max(1, cores // workers),
            "sample": f"{n_streams} streams x {n_frames} stereo frames; OpenCV stages = cv2 {cv2.__version__} (the library "
                      f"the reference links: 2x calcOpticalFlowPyrLK + goodFeaturesToTrack), local BA = oracle/ba_ref.c "
                      f"(C port of the vendored g2o LM/Schur path, 1 thread per stream like g2o)"
                      + ("" if ba is not None else " [BA not in sample]")}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    S = args.streams
    n_streams = min(S, os.cpu_count() or 1)
    steps = max(1, min(args.steps, 6))
    for _ in range(min(args.warmup, 1)):
        cpu_baseline(min(n_streams, 4), 1, args)
    cpu = cpu_baseline(n_streams, steps, args)
    out = {"impl": "reference",
           "metric": "frames/sec (device-timed), EuRoC-shaped 752x480 stereo, LK frontend + 10-KF local BA",
           "value": cpu["value"], "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": min(args.warmup, 1),
           "ms_per_step": 1e3 * n_streams / cpu["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8/i16 fixed-point + f32 (OpenCV), f64 (g2o port)", "data": "synthetic",
           "config": {"workload": f"{S} concurrent EuRoC-shaped 752x480 stereo streams (bounded sample: {n_streams} streams x "
                                  f"{steps} frames on the host cores)", "streams_per_gpu": S, "image": [W, H],
                      "points_per_stream": NPTS},
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="euroc", choices=sorted(WORKLOADS), help="euroc = the headline (default)")
    args = ap.parse_args()
    global W, H, STEREO, BA_WINDOW, BA_LANDMARKS, WORKLOAD, P_PYR, LK_BYTES_PER_CALL
    wl = WORKLOADS[args.workload]
    WORKLOAD, W, H, STEREO, BA_WINDOW, BA_LANDMARKS = args.workload, wl["w"], wl["h"], wl["stereo"], wl["window"], wl["landmarks"]
    if args.workload != "euroc":
        P_PYR = sum(((W + (1 << l) - 1) >> l) * ((H + (1 << l) - 1) >> l) for l in range(4))
        LK_BYTES_PER_CALL = 6 * P_PYR + 29 * NPTS
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
