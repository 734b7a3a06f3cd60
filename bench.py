#!/usr/bin/env python
"""bench.py -- frames/sec of the FLVIS hot path on EuRoC-shaped 752x480 stereo+IMU streams (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--streams S] [--workload euroc|kitti|d435] [--impl reference]

A "step" is ONE CAMERA FRAME for every one of the S concurrent sequences on this GPU, through the product's public call
(flv_f2f_batch_imu_feed_many + flv_f2f_batch_image_feed, include/flvis_b200_host.h) -- the complete
F2FTracking::image_feed of the reference (src/frontend/f2f_tracking.cpp:59-400) per sequence:
    ingest (equalizeHist for EuRoC) + 2 pyramids -> IMU pose guess -> LK frame->frame -> keep rule -> F-matrix RANSAC ->
    PnP RANSAC -> IMU roll/pitch blend -> pose-only BA -> reprojection cull -> FeatureDEM redetect (Shi-Tomasi + region
    select) -> new landmarks -> LK left->right -> triangulation / depth filter -> keyframe rule, IMU bias feedback
and every keyframe goes to the local-map worker (10-keyframe sliding-window BA, 12 + cull + 8 LM iterations,
src/backend/vo_localmap.cpp:87-380), which runs concurrently like FLVIS's local-map nodelet; the timed region ends when
the last window it triggered is solved.
`value`  : device-timed (CUDA events on the compute stream), frames already resident in HBM (device-side frame pool).
`e2e`    : the same call with the frames in pinned HOST memory: both images of every sequence cross PCIe every step and the
           per-sequence results (pose, landmark lists) are read back every step.
Multi-GPU: sequences are independent; rank r owns global streams [r*S, (r+1)*S) ("weak" scaling), no
collective in the frame loop; the per-frame results of every rank are all-gathered (NCCL) once per timed region.
`--impl reference`: the reference's CPU path on the host cores -- the OpenCV calls through cv2 (the library the reference
links) and the C port of its g2o solver -- same frame, bounded sample.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

# every stream group, the image-upload, read-back and local-map streams need their own hardware queue: with the default of
# 8 connections streams alias and serialise behind one another (must be set before the CUDA context exists)
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from flvis_b200 import sharding  # noqa: E402
from synthdata import sequences  # noqa: E402

NPTS_MAX = 480                                    # 16 regions x 30 landmarks (feature_dem.cpp:21,188,250)
WORKLOADS = {   # BASELINE.json configs: image size, sensor, local-BA window
    "euroc": dict(w=752, h=480, stereo=True, window=10, imu=True,
                  name="EuRoC-shaped 752x480 raw stereo (radtan lenses, STEREO_UNRECT, equalizeHist) + IMU 200 Hz, 10-KF local BA (configs[1]/[2])"),
    "kitti": dict(w=1241, h=376, stereo=True, window=20, imu=False,
                  name="KITTI-shaped 1241x376 rectified stereo, no IMU, 20-KF local-BA window (configs[3])"),
    "d435": dict(w=640, h=480, stereo=False, window=8, imu=True,
                 name="640x480 D435i depth + IMU 200 Hz, 8-KF local BA (configs[0]/[4])"),
}
PERIOD = 40                                       # frames per period of the synthetic rig trajectory
# profiles/r02_final_step_kernels_ncu_full.csv (ncu --set full inside the bench step, 32 sequences x 420 landmarks in one launch,
# frame->frame call of lk_track_kernel_v4<TMA>): dram__bytes_read 91.6 MB + write 3.8 MB, 192.6 M warp instructions
LK_NCU_TRAFFIC_PER_SEQUENCE = {"euroc": (91.63e6 + 3.82e6) / 32}
LK_WARP_INSTR_PER_POINT = 192.6e6 / (32 * 420.5)


def pyramid_pixels(w, h):
    """P(w,h) of SURVEY.md 8(d): pixels of the levels OpenCV builds for a 31x31 window."""
    tot, lw, lh = w * h, w, h
    for _ in range(3):
        lw, lh = (lw + 1) // 2, (lh + 1) // 2
        if lw <= 31 or lh <= 31:
            break
        tot += lw * lh
    return tot


def lk_algorithmic_bytes(w, h, n_pts):
    """both u8 pyramids read once + 29 B of point I/O per tracked point."""
    return 2 * pyramid_pixels(w, h) + 29 * n_pts


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("sm_max_mhz", 1965.0)), "measured (MEASURED_PEAKS.json)"
    return 6650.0, 1965.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------------
def build_pools(workload, first_stream, n_streams, procs):
    """One period (+ start-up) of rendered frames and IMU samples per stream; rendering is CPU work done in a fork pool
    BEFORE any CUDA initialisation."""
    jobs = [(workload, first_stream + s, PERIOD) for s in range(n_streams)]
    if procs > 1 and n_streams > 1:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(min(procs, n_streams)) as pool:
            res = pool.map(sequences.render_pool, jobs)
    else:
        res = [sequences.render_pool(j) for j in jobs]
    f0 = np.stack([r[0] for r in res], 1)         # [n][S][h][w]
    f1 = np.stack([r[1] for r in res], 1)
    imu = [r[2] for r in res]                     # [S][n] arrays (k, 7)
    return f0, f1, imu


class FrameFeeder:
    """Maps step i to (pool frame, time offset) and holds the IMU samples of every pool frame for all streams."""

    def __init__(self, workload, first_stream, S, imu):
        self.seq0 = sequences.make_bench(workload, first_stream, PERIOD, render=False)
        self.n_startup, self.period, self.hz, self.S = self.seq0.n_startup, PERIOD, self.seq0.img_hz, S
        self.Tp = PERIOD / self.hz
        self.imu = []
        for f in range(self.n_startup + PERIOD):
            rows = [imu[s][f] for s in range(S)]
            st = np.concatenate([np.full(len(r), s, np.int32) for s, r in enumerate(rows)])
            a = np.concatenate(rows)
            self.imu.append((np.ascontiguousarray(st), np.ascontiguousarray(a[:, 0]), np.ascontiguousarray(a[:, 1:4]), np.ascontiguousarray(a[:, 4:7])))

    def locate(self, i):
        if i < self.n_startup:
            return i, 0.0
        k = i - self.n_startup
        return self.n_startup + k % self.period, (k // self.period) * self.Tp

    def time_of(self, i):
        return i / self.hz


# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    """SM clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe), NVML polling."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            bits = [pynvml.nvmlClocksThrottleReasonHwSlowdown, pynvml.nvmlClocksThrottleReasonHwThermalSlowdown,
                    pynvml.nvmlClocksThrottleReasonSwThermalSlowdown, pynvml.nvmlClocksThrottleReasonSwPowerCap]
            while not self.stop_flag.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append([float(sm), float(mx)] + [bool(r & b) for b in bits])
                self.stop_flag.wait(0.005)
            return
        except Exception:
            pass
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    c = [x.strip() for x in out.split(",")]
                    self.rows.append([float(c[0]), float(c[1])] + [x.lower().startswith("active") for x in c[2:6]])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i] for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows)}


def _cam_centre(T7):
    x, y, z, w = T7[:4]
    R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                  [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                  [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])
    return -R.T @ np.asarray(T7[4:], float)


def _issue_roofline(lk_us, S, n_pts, sm_mhz):
    """Secondary roofline for LK: warp instructions per launch against the chip's issue rate (4 schedulers x 148 SMs x clock)."""
    if not lk_us:
        return None
    instr = LK_WARP_INSTR_PER_POINT * S * n_pts
    peak = 4 * 148 * sm_mhz * 1e6
    return {"warp_instructions_per_launch": instr, "achieved_per_s": instr / (lk_us * 1e-6), "peak_per_s": peak,
            "frac": instr / (lk_us * 1e-6) / peak, "source": "instructions per point from the ncu capture in profiles/, time from this run"}


def _ba_roofline(t, peak):
    """per LM iteration K7 build 168 E + 392 P + 120 L, K8 Schur 144 E + 96 L + 288 Pf^2 + 48 Pf, K10 update 160 E + 120 L + 56 P
    bytes (one trial per iteration counted: a lower bound), against the wall time of the solver calls."""
    n = t["solves"]
    if not n or t["solve_ms"] <= 0:
        return None
    E, L, P, it = t["edges"] / n, t["landmarks"] / n, t["poses"] / n, t["lm_iterations"] / n
    Pf = max(P - 1, 0)
    per_iter = (168 * E + 392 * P + 120 * L) + (144 * E + 96 * L + 288 * Pf * Pf + 48 * Pf) + (160 * E + 120 * L + 56 * P)
    alg = per_iter * it
    gbs = alg * n / (t["solve_ms"] * 1e-3) / 1e9
    return {"kernel": "ba_kernel (local map windows, cluster of 4 CTAs per window)", "bound": "hbm", "unit": "GB/s", "achieved": gbs, "peak": peak,
            "frac": gbs / peak, "algorithmic_bytes_per_window": alg, "mean_edges": E, "mean_landmarks": L, "mean_poses": P, "mean_lm_iterations": it,
            "note": "aggregate over the windows solved concurrently; a window's working set (< 1 MB) lives in L2 / shared memory, "
                    "the kernel is latency-bound (barrier and dependent-load stalls, profiles/README.md), not HBM-bound"}


# ------------------------------------------------------------------------------------------------
def run_ours(args, wl, pools):
    import torch
    import torch.distributed as dist
    from flvis_b200 import batch as fb

    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # host threads per rank that wait on the GPU most of the time: one per stream group, the local-map shards, this thread.
    # CUDA's default is to spin while waiting; when the ranks of a node have more such threads than the node has cores, they
    # sleep instead (cudaDeviceScheduleBlockingSync, set by the library for every context it creates)
    waiting_threads = args.groups + 2 + 1
    if world * waiting_threads > (os.cpu_count() or 1):
        os.environ.setdefault("FLV_BLOCKING_SYNC", "1")
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")     # keep stdout to the one JSON line
        dist.init_process_group("nccl", init_method="env://", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    S = args.streams
    W, H = wl["w"], wl["h"]
    first = sharding.stream_ids(rank, world, S)[0]
    f0, f1, imu = pools
    feeder = FrameFeeder(args.workload, first, S, imu)
    cfg, lenses, equalize, K, _ = fb.config_for(feeder.seq0)
    h_pool0 = torch.from_numpy(f0).pin_memory()
    h_pool1 = torch.from_numpy(f1.view(np.uint8) if f1.dtype == np.uint16 else f1).pin_memory()
    d_pool0 = h_pool0.to(dev); d_pool1 = h_pool1.to(dev)
    stream = torch.cuda.Stream(dev)

    def make_tracker(n):
        t = fb.BatchTracker(cfg, n, device=local_rank, lenses=lenses, equalize=equalize, groups=args.groups if n == S else 1)
        t.set_stream(stream.cuda_stream)
        t.set_readback(False)                      # keyframe / pose consumers only need the lean landmark records
        lm = fb.LocalMapBatch(n, wl["window"], K, device=local_rank)
        t.attach_localmap(lm)
        return t, lm

    trk, lmap = make_tracker(S)
    tvec = np.zeros(S)

    def feed(t, i, mode, n=S, pipelined=False):
        f, off = feeder.locate(i)
        imu_now = None
        if wl["imu"]:
            st, ti, acc, gyro = feeder.imu[f]
            if n != S:
                m = st < n
                st, ti, acc, gyro = np.ascontiguousarray(st[m]), np.ascontiguousarray(ti[m]), np.ascontiguousarray(acc[m]), np.ascontiguousarray(gyro[m])
            imu_now = (st, (ti + off) if off else ti, acc, gyro)
        tv = tvec[:n]
        tv[:] = feeder.time_of(i)
        p0, p1 = (h_pool0[f].data_ptr(), h_pool1[f].data_ptr()) if mode == "host" else (d_pool0[f].data_ptr(), d_pool1[f].data_ptr())
        if pipelined:                             # groups take turns: one group's host work hides behind the others' kernels
            t.frame_async(tv, p0, p1, mode != "host", imu_now)
        else:
            if imu_now is not None:
                t.imu_feed_many(*imu_now)
            t.image_feed(tv, p0, p1, mode != "host")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # per-frame results of every rank, all-gathered once per timed region (north_star: NCCL only for the batch gather)
    gather = sharding.ResultGather(dist, args.steps, S, dev, stream) if world > 1 else None

    step_no = [0]
    wall = dict(loop_ms=0.0, drain_ms=0.0, regions=0)
    traj = [[] for _ in range(S)]

    def region(t, lm, steps, mode, n=S, record=False):
        barrier()
        l0 = t.launches(); s0 = lm.stats()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        tw0 = time.perf_counter()
        pipe = t.groups > 1 and not record
        if gather is not None and n == S:
            t.set_result_log(gather.h.data_ptr(), steps)      # the library writes every finished frame's records into the staging block
        for j in range(steps):
            feed(t, step_no[0], mode, n, pipelined=pipe)
            if record:
                tt = feeder.time_of(step_no[0])
                for s in range(n):
                    if t.state(s) == "Tracking":
                        traj[s].append((tt, t.pose(s)))
            step_no[0] += 1
        if pipe:
            t.sync()
        tw1 = time.perf_counter()
        lm.wait()                                  # the region ends when the last local-BA window it triggered is solved
        wall["loop_ms"] += (tw1 - tw0) * 1e3; wall["drain_ms"] += (time.perf_counter() - tw1) * 1e3; wall["regions"] += 1
        if gather is not None and n == S:
            t.set_result_log(0, 1)
            gather.flush()
            gather.wait()
        ev1.record(stream)
        barrier()
        ms = ev0.elapsed_time(ev1)
        s1 = lm.stats()
        return ms, t.launches() - l0 + (s1["launches"] - s0["launches"]), {k: s1[k] - s0[k] for k in s1}

    # start-up (UnInit -> IMU initialised -> first keyframe) + warm-up, untimed
    for _ in range(feeder.n_startup + max(args.warmup, 3)):
        feed(trk, step_no[0], "device"); step_no[0] += 1
    lmap.wait()
    n_tracking = sum(trk.state(s) == "Tracking" for s in range(S))

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    trk.set_profile(True)
    dev_runs, e2e_runs, launches, ba_tot = [], [], 0, dict(keyframes=0, solves=0, launches=0, solve_ms=0.0, host_ms=0.0, edges=0.0, landmarks=0.0, poses=0.0, lm_iterations=0.0)
    for r in range(args.reps):
        ms, nl, ba = region(trk, lmap, args.steps, "device")
        dev_runs.append(ms); launches = nl
        for k in ba_tot:
            ba_tot[k] += ba[k]
    stage_ms, prof_frames = trk.profile()
    host_ms = trk.host_profile()
    wall_dev = dict(wall)
    region(trk, lmap, args.steps, "device", record=True)      # untimed: the trajectory for the ATE figure
    trk.set_profile(False)
    n_lm_mean = float(np.mean([trk.n_landmarks(s) for s in range(S)]))
    for r in range(args.reps):
        e2e_runs.append(region(trk, lmap, args.steps, "host")[0])
    if sampler:
        sampler.stop_flag.set(); sampler.join(timeout=2)
    still_tracking = sum(trk.state(s) == "Tracking" for s in range(S))
    per_rank_ms = None
    if world > 1:
        t_all = [torch.zeros(1, device=dev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(t_all, torch.tensor([float(np.median(dev_runs))], device=dev, dtype=torch.float64))
        per_rank_ms = [float(x.item()) for x in t_all]
    # job time = slowest rank, repetition by repetition
    dev_runs = [sharding.max_over_ranks(m, dist if world > 1 else None, dev) for m in dev_runs]
    e2e_runs = [sharding.max_over_ranks(m, dist if world > 1 else None, dev) for m in e2e_runs]

    # BASELINE configs[1]: ONE stream on the GPU (latency-oriented), same call
    single = None
    if world == 1 and S > 1 and not args.no_single:
        t1, lm1 = make_tracker(1)
        keep = step_no[0]
        step_no[0] = 0
        for _ in range(feeder.n_startup + 3):
            feed(t1, step_no[0], "device", 1); step_no[0] += 1
        k1 = max(args.steps, 40)
        ms1 = sorted(region(t1, lm1, k1, "device", 1)[0] for _ in range(3))[1]
        ms1h = sorted(region(t1, lm1, k1, "host", 1)[0] for _ in range(3))[1]
        single = {"workload": "1 sequence, " + wl["name"], "steps": k1, "value": k1 / (ms1 * 1e-3), "e2e": k1 / (ms1h * 1e-3),
                  "unit": "frames/s", "ms_per_frame": ms1 / k1}
        step_no[0] = keep
        t1.close(); lm1.close()

    out = None
    if rank == 0:
        frames = args.steps * S * world
        med = float(np.median(dev_runs)); med_e = float(np.median(e2e_runs))
        peak, sm_max, peak_src = load_peaks()
        per_frame = stage_ms / max(prof_frames, 1)
        lk_us = 1e3 * float(per_frame[1])                                     # frame->frame call (maxLevel 10)
        lk_lr_us = 1e3 * float(per_frame[7]) if wl["stereo"] else None        # + the tiny right-point undistortion kernel
        alg = lk_algorithmic_bytes(W, H, n_lm_mean) * S / max(trk.groups, 1)     # one launch serves one group's sequences
        achieved = alg / (lk_us * 1e-6) / 1e9 if lk_us > 0 else 0.0
        # ATE of the GPU trajectories against the synthetic ground truth (camera centres, constant offset removed)
        ates = []
        for s in range(S):
            if len(traj[s]) < 5:
                continue
            gt_seq = sequences.make_bench(args.workload, first + s, PERIOD, render=False)
            c = np.array([_cam_centre(T) for _, T in traj[s]]); g = np.array([gt_seq.T_w_c0(tt).t for tt, _ in traj[s]])
            d = c - g
            ates.append(float(np.sqrt(np.mean(np.sum((d - d.mean(0)) ** 2, axis=1)))))
        img1_bytes = W * H * (1 if wl["stereo"] else 2)
        h2d = S * (W * H + img1_bytes) + S * (160 + 512 * 4)                  # images + control block + dummy-depth table
        d2h = S * 120 + S * (512 * 50 + 60)                                   # summaries + lean landmark records (50 B per landmark slot)
        cpu = cpu_baseline(args, wl, pools, min(S, os.cpu_count() or 1), 6) if world == 1 and not args.no_cpu else None
        if single is not None and cpu is not None:
            single["cpu_value"] = cpu_baseline(args, wl, pools, 1, 8)["value"]
        out = {
            "metric": f"frames/sec (device-timed), {wl['name']}",
            "value": frames / (med * 1e-3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": med / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/i32 fixed-point + f32 (LK, Shi-Tomasi), f64 (RANSAC, geometry, BA)", "data": "synthetic",
            "config": {"workload": f"{S} concurrent sequences per GPU, {wl['name']}; full F2FTracking::image_feed per frame (LK x2, F + PnP RANSAC, "
                                   f"pose-only BA, FeatureDEM redetect, depth innovation) + local BA W={wl['window']} on every keyframe"
                                   + (" (BASELINE configs[2])" if args.workload == "euroc" else ""),
                       "streams_per_gpu": S, "stream_groups": trk.groups, "image": [W, H], "landmarks_per_stream_mean": n_lm_mean,
                       "streams_tracking_before_after": [n_tracking, still_tracking],
                       "repetitions": args.reps, "value_min_max": [frames / (max(dev_runs) * 1e-3), frames / (min(dev_runs) * 1e-3)],
                       "l2_note": "no explicit L2 flush: every step ingests fresh images for all sequences from a rotating frame pool of "
                                  f"{f0.shape[0]} frames per sequence ({round((f0.nbytes + f1.nbytes) / 1e6)} MB, larger than the 126 MB L2) and "
                                  "rewrites every pyramid",
                       "parallelism": f"sequences sharded {S}/GPU x {world} GPU, no collective in the frame loop; per-frame results of all ranks "
                                      "all-gathered (NCCL) once per timed region" + ("" if world > 1 else " when N > 1"),
                       "per_rank_ms": per_rank_ms, "host_wait": "blocking" if os.environ.get("FLV_BLOCKING_SYNC") == "1" else "spin"},
            "e2e": {"value": frames / (med_e * 1e-3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": med_e / args.steps, "value_min_max": [frames / (max(e2e_runs) * 1e-3), frames / (min(e2e_runs) * 1e-3)]},
            "gpu_launches": int(launches),
            "stages_ms_per_step": {n: round(float(v), 4) for n, v in zip(fb.STAGES, per_frame)},
            "host_ms_per_step": {"decisions": float(host_ms[0] / max(prof_frames, 1)), "enqueue": float(host_ms[1] / max(prof_frames, 1)),
                                 "wait_for_device": float(host_ms[2] / max(prof_frames, 1)), "post_frame": float(host_ms[3] / max(prof_frames, 1)),
                                 "local_map_worker": ba_tot["host_ms"] / (args.reps * args.steps),
                                 "frame_loop_wall": wall_dev["loop_ms"] / (args.reps * args.steps),
                                 "local_map_drain_at_region_end": wall_dev["drain_ms"] / (args.reps * args.steps)},
            "ba": {"ms_per_kf": (ba_tot["solve_ms"] / ba_tot["solves"]) if ba_tot["solves"] else None,
                   "note": "wall time of one batched flv_ba_optimize call (H2D of the window arrays + ba_kernel + D2H) divided by the "
                           "windows it solved; the worker overlaps the tracker's kernels",
                   "window": wl["window"], "keyframes": ba_tot["keyframes"], "solves": ba_tot["solves"],
                   "windows_per_launch": ba_tot["solves"] / max(ba_tot["launches"], 1),
                   "roofline": _ba_roofline(ba_tot, peak),
                   "keyframes_per_step": ba_tot["keyframes"] / (args.reps * args.steps)},
            "ate": {"vs_ground_truth_m_mean": float(np.mean(ates)) if ates else None, "vs_ground_truth_m_max": float(np.max(ates)) if ates else None,
                    "frames": args.steps, "note": "RMSE of the camera centres against the synthetic rig trajectory over the first timed "
                                                  "region; ATE against the reference path is asserted in tests/test_configs_gpu.py"},
            "single_stream": single,
            "roofline": {"kernel": "lk_track_kernel_v4 (frame->frame call)", "bound": "hbm", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (LK_NCU_TRAFFIC_PER_SEQUENCE[args.workload] * S / max(trk.groups, 1)) if args.workload in LK_NCU_TRAFFIC_PER_SEQUENCE else None,
                         "traffic_note": "dram read + write of one ncu --set full capture (cold L2, 32 sequences per launch), scaled to this run's sequences per launch",
                         "peak_source": peak_src, "us_per_launch": lk_us, "us_per_launch_left_right": lk_lr_us,
                         "algorithmic_bytes_per_launch": alg,
                         "algorithmic_bytes_def": "SURVEY.md 8(d): (2 P(w,h) + 29 N) x S, P = pyramid pixels, N = mean tracked points per sequence",
                         "traffic_expected": (6 * pyramid_pixels(W, H) + 29 * n_lm_mean) * S / max(trk.groups, 1),
                         "traffic_expected_def": "what the kernel must read with its 4 B/px Scharr derivative pyramid of the first image",
                         "issue_slot_roofline": _issue_roofline(lk_us, S / max(trk.groups, 1), n_lm_mean, sm_max),
                         "launch_note": f"one launch tracks {S // max(trk.groups, 1)} sequences; with {trk.groups} stream groups the other group's "
                                        "kernels share the SMs while it runs, so us_per_launch is the in-situ duration, not the kernel alone",
                         "note": "LK is bound by instruction issue, not by HBM (ncu: DRAM < 3 % busy); the issue-slot figure is the "
                                 "roofline that explains it (DESIGN.md section 5)"},
            "clocks": sampler.summary() if sampler else None,
        }
        if cpu:
            out["cpu_baseline"] = cpu
    trk.close(); lmap.close()
    if world > 1:
        dist.destroy_process_group()              # the collective part is over: the side measurements below are rank 0's own
    if out is not None:
        out["roofline"]["isolated_launch"] = lk_isolated(f0, feeder.n_startup, W, H, load_peaks())
        if world == 1 and not args.no_others and args.workload == "euroc":
            out["other_workloads"] = other_workloads(args, S)
        print(json.dumps(out))
    return out


def lk_isolated(f0, first_frame, W, H, peaks):
    """The frame->frame LK call ALONE: all S sequences of this rank in ONE launch on an otherwise idle GPU, L2 flushed before every
    launch, CUDA events on the launching stream (the in-step figure above is taken while the other stream groups' kernels share the
    SMs).  Two consecutive pool frames, points = Shi-Tomasi corners of the first one (<= 480 per sequence), err = NULL as in the
    tracker.  A side measurement: any failure is reported in place of the numbers."""
    try:
        import torch
        from flvis_b200 import capi
        peak, sm_max, _ = peaks
        S = f0.shape[1]
        ctx = capi.Context(S, W, H, 512)
        st = torch.cuda.Stream()
        ctx.set_stream(st.cuda_stream)
        a = min(first_frame, f0.shape[0] - 2)
        ctx.upload(0, np.ascontiguousarray(f0[a])); ctx.upload(1, np.ascontiguousarray(f0[a + 1]))
        ctx.build_pyramid(0, S); ctx.build_pyramid(1, S)
        corners = ctx.gftt(0, S, NPTS_MAX, 0.01, 10)
        pts = np.zeros((S, 512, 2), np.float32); n = np.zeros(S, np.int32)
        for s_, c in enumerate(corners):
            k = min(len(c), NPTS_MAX)
            pts[s_, :k] = c[:k]; n[s_] = k
        with torch.cuda.stream(st):
            d_n = torch.from_numpy(n).cuda(); d_prev = torch.from_numpy(pts).cuda(); d_next = torch.empty_like(d_prev)
            d_st = torch.empty((S, 512), dtype=torch.uint8, device="cuda")
            flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

            def launch():
                ctx.lk_track_dev(0, 1, S, d_n.data_ptr(), d_prev.data_ptr(), d_prev.data_ptr(), d_next.data_ptr(), d_st.data_ptr(), None)
            for _ in range(3):
                launch()
            ts = []
            for _ in range(9):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st); launch(); e1.record(st)
                st.synchronize()
                ts.append(e0.elapsed_time(e1) * 1e3)
            tracked = int(d_st.sum().item())
        ctx.close()
        us = float(np.median(ts)); n_mean = float(n.mean())
        alg = lk_algorithmic_bytes(W, H, n_mean) * S
        gbs = alg / (us * 1e-6) / 1e9
        instr = LK_WARP_INSTR_PER_POINT * S * n_mean
        return {"sequences_per_launch": S, "points_per_sequence_mean": n_mean, "tracked": tracked, "us_per_launch": us,
                "us_min_max": [float(min(ts)), float(max(ts))], "algorithmic_bytes_per_launch": alg, "achieved": gbs, "unit": "GB/s",
                "frac": gbs / peak, "issue_slot_frac": instr / (us * 1e-6) / (4 * 148 * sm_max * 1e6),
                "l2": "256 MB written before every launch"}
    except Exception as e:                          # a side measurement must not take the headline line down
        return {"error": repr(e)[:300]}


def other_workloads(args, S):
    """BASELINE.json's other configurations (KITTI-shaped stereo, W=20; D435i depth + IMU, W=8) through the same bench, shortened
    (3 repetitions of <= 10 steps -- the median survives one slow repetition --, no CPU legs), each in its own process after this one released the GPU."""
    res = {}
    for name in ("kitti", "d435"):
        cmd = [sys.executable, os.path.abspath(__file__), "--workload", name, "--reps", "3", "--steps", str(min(args.steps, 10)), "--warmup", "3",
               "--streams", str(S), "--groups", str(args.groups), "--no-cpu", "--no-single", "--no-others"]
        try:
            r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
            line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
            j = json.loads(line)
            res[name] = {"workload": j["config"]["workload"], "value": j["value"], "e2e": j["e2e"]["value"], "unit": "frames/s",
                         "ms_per_step": j["ms_per_step"], "steps": j["steps"], "repetitions": j["config"]["repetitions"],
                         "ba_window": j["ba"]["window"], "ba_ms_per_kf": j["ba"]["ms_per_kf"], "keyframes_per_step": j["ba"]["keyframes_per_step"],
                         "ate_vs_ground_truth_m_mean": j["ate"]["vs_ground_truth_m_mean"]}
        except Exception as e:                      # a side measurement must not take the headline line down
            res[name] = {"error": repr(e)[:200]}
    return res


# ------------------------------------------------------------------------------------------------
# The reference's CPU path for the same frame: every OpenCV call FLVIS makes per frame through cv2 (the same library),
# its g2o solves through the C port (oracle/ba_ref.c).  The Python between the calls is array slicing only.
def cpu_frame_chain(cv2, ba_ref, wl, rig, prev0, cur0, cur1, pts, p3d, T_guess, ba_window_problem, is_kf):
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001)
    if rig["equalize"]:
        cur0 = cv2.equalizeHist(cur0)                                                         # f2f_tracking.cpp:141-145
        if cur1 is not None:
            cur1 = cv2.equalizeHist(cur1)
    nxt, st, _ = cv2.calcOpticalFlowPyrLK(prev0, cur0, pts, pts.copy(), winSize=(31, 31), maxLevel=10, criteria=crit,
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW)                 # lkorb_tracking.cpp:64-73
    ok = st.ravel() == 1
    a, b = pts[ok], nxt[ok]
    if rig["lens0"] is not None and len(a):                                                   # lkorb_tracking.cpp:86-89
        K0, D0, R0, P0 = rig["lens0"]
        a = cv2.undistortPoints(a.reshape(-1, 1, 2), K0, D0, R=R0, P=P0[:3, :3]).reshape(-1, 2)
        b = cv2.undistortPoints(b.reshape(-1, 1, 2), K0, D0, R=R0, P=P0[:3, :3]).reshape(-1, 2)
    if len(a) >= 8:
        cv2.findFundamentalMat(a, b, cv2.FM_RANSAC, 5.0, 0.99)                                # :134-135
    X = p3d[ok]
    if len(X) >= 10:
        rvec, tvec = T_guess
        cv2.solvePnPRansac(X, b, rig["Km"], np.zeros(4), rvec.copy(), tvec.copy(), True, 100, 3.0, 0.99, flags=cv2.SOLVEPNP_ITERATIVE)   # :172-176
        n = len(X)                                                                            # optimize_in_frame.cpp:10-90
        d = ba_ref.BAData(rig["pose7"][None].copy(), X.astype(np.float64), np.zeros(n, np.int32), np.arange(n, dtype=np.int32),
                          b.astype(np.float64), rig["K4"], fixed_pose=-1, fix_landmarks=1)
        ba_ref.optimize(d, 2, 2, min_edges_after_cull=10)
    mask = np.full(cur0.shape, 255, np.uint8)
    cv2.goodFeaturesToTrack(cur0, rig["gftt"][0], rig["gftt"][1], rig["gftt"][2], mask=mask)  # feature_dem.cpp:160
    if cur1 is not None:
        r, _, _ = cv2.calcOpticalFlowPyrLK(cur0, cur1, nxt, nxt.copy(), winSize=(31, 31), maxLevel=5, criteria=crit,
                                           flags=cv2.OPTFLOW_USE_INITIAL_FLOW)                # camera_frame.cpp:124-128
        if rig["lens1"] is not None:
            K1, D1, R1, P1 = rig["lens1"]
            cv2.undistortPoints(r.reshape(-1, 1, 2), K1, D1, R=R1, P=P1[:3, :3])              # :130
    if is_kf and ba_window_problem is not None:                                               # vo_localmap.cpp:292-319
        p = ba_window_problem
        ba_ref.optimize(ba_ref.BAData(p.poses.copy(), p.lms.copy(), p.ep, p.el, p.uv, p.K, p.fixed_pose, p.fix_landmarks), 12, 8)
    nxt[~ok] = pts[~ok]
    return nxt, cur0


def cpu_baseline(args, wl, pools, n_streams, n_frames):
    """Reference CPU path on this box's host cores, bounded sample: n_streams sequences x n_frames frames."""
    import cv2
    from concurrent.futures import ThreadPoolExecutor
    from oracle import ba_ref                      # the cpu_baseline leg: the one place bench.py executes oracle/
    from flvis_b200 import batch as fb
    from synthdata import ba_problems
    cores = os.cpu_count() or 1
    workers = min(n_streams, cores)
    cv2.setNumThreads(max(1, cores // workers))
    f0, f1, _ = pools
    seq0 = sequences.make_bench(args.workload, 0, PERIOD, render=False)
    cfg, lenses, equalize, K, rect = fb.config_for(seq0)
    c = seq0.cfg
    rig = dict(equalize=equalize, K4=tuple(float(v) for v in K), Km=np.array([[K[0], 0, K[2]], [0, K[1], K[3]], [0, 0, 1.0]]),
               gftt=(int(c["feature_para"][3]), float(c["feature_para"][4]), float(c["feature_para"][5])),
               lens0=None, lens1=None, pose7=np.array([0, 0, 0, 1.0, 0, 0, 0]))
    if rect is not None:
        rig["lens0"] = (rect["K0"], rect["D0"], rect["R0"], rect["P0"]); rig["lens1"] = (rect["K1"], rect["D1"], rect["R1"], rect["P1"])
    base = seq0.n_startup
    kf_every = 4                                                                              # the GPU arm's keyframe rate on this trajectory
    obs = 330                                                                                 # inlier landmarks with depth per keyframe in the GPU arm
    windows = [ba_problems.make_problem(wl["window"], 3 * obs, obs, seed=900 + s, K=rig["K4"], w=wl["w"], h=wl["h"]) for s in range(n_streams)]
    prev, pts0 = [], []
    for s in range(n_streams):
        img = cv2.equalizeHist(f0[base, s]) if equalize else f0[base, s]
        prev.append(img)
        pts0.append(cv2.goodFeaturesToTrack(img, NPTS_MAX, 0.01, 10).reshape(-1, 2))
    rv, tv = np.zeros((3, 1)), np.zeros((3, 1))
    Z = 3.0 if args.workload != "kitti" else 12.0

    def work(s):
        pts, last = pts0[s], prev[s]
        for t in range(1, n_frames + 1):
            # a fronto-parallel cloud at the plane depth gives PnP / BA the right problem size and conditioning
            p3d = np.stack([(pts[:, 0] - K[2]) / K[0] * Z, (pts[:, 1] - K[3]) / K[1] * Z, np.full(len(pts), Z)], 1).astype(np.float32)
            i1 = f1[base + t, s] if wl["stereo"] else None
            pts, last = cpu_frame_chain(cv2, ba_ref, wl, rig, last, f0[base + t, s], i1, pts, p3d, (rv, tv), windows[s], (t + s) % kf_every == 0)
        return 0

    t0 = time.perf_counter()
    with ThreadPoolExecutor(workers) as ex:
        list(ex.map(work, range(n_streams)))
    dt = time.perf_counter() - t0
    return {"value": n_streams * n_frames / dt, "unit": "frames/s", "cores": cores, "kind": "port",
            "threads": workers * max(1, cores // workers),
            "sample": f"{n_streams} sequences x {n_frames} frames; per frame the OpenCV calls of F2FTracking::image_feed through cv2 "
                      f"{cv2.__version__} (equalizeHist, 2x calcOpticalFlowPyrLK, undistortPoints, findFundamentalMat, solvePnPRansac, "
                      f"goodFeaturesToTrack) + pose-only BA and, every {kf_every}th frame, a W={wl['window']} local-BA window "
                      f"({obs} observations per keyframe) through oracle/ba_ref.c (C port of the vendored g2o LM/Schur path, one thread "
                      f"per solve like g2o); Python between the calls is array slicing only"}


def run_reference(args, wl, pools):
    world = int(os.environ.get("WORLD_SIZE", "1"))
    S = args.streams
    n_streams = min(S, os.cpu_count() or 1)
    steps = max(1, min(args.steps, 6))
    cpu_baseline(args, wl, pools, min(n_streams, 4), 1)
    cpu = cpu_baseline(args, wl, pools, n_streams, steps)
    out = {"impl": "reference", "metric": f"frames/sec (device-timed), {wl['name']}",
           "value": cpu["value"], "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": 1,
           "ms_per_step": 1e3 * n_streams / cpu["value"], "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "u8/i16 fixed-point + f32 (OpenCV), f64 (g2o port)", "data": "synthetic",
           "config": {"workload": f"{S} concurrent sequences, {wl['name']} (bounded sample: {n_streams} sequences x {steps} frames on the host cores)",
                      "streams_per_gpu": S, "image": [wl["w"], wl["h"]]},
           "cpu_baseline": cpu,
           "e2e": {"value": cpu["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--streams", type=int, default=32)
    ap.add_argument("--groups", type=int, default=4, help="stream groups per GPU (own CUDA stream + host thread each)")
    ap.add_argument("--reps", type=int, default=7, help="repetitions of the K-step timed region (median reported)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-single", action="store_true", help="skip the single-stream leg")
    ap.add_argument("--no-others", action="store_true", help="skip the short kitti / d435 side measurements (N=1, euroc only)")
    ap.add_argument("--workload", default="euroc", choices=sorted(WORKLOADS), help="euroc = the headline (default)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    cores = os.cpu_count() or 1
    if args.impl == "reference":
        if rank != 0:
            return
        n = min(args.streams, cores)
        run_reference(args, wl, build_pools(args.workload, 0, n, cores))
        return
    first = sharding.stream_ids(rank, world, args.streams)[0]
    pools = build_pools(args.workload, first, args.streams, max(1, cores // world))
    run_ours(args, wl, pools)


if __name__ == "__main__":
    main()
