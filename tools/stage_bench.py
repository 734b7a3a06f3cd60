"""Per-stage device times of one bench step (CUDA events on the library stream); run on the GPU box.
usage: python tools/stage_bench.py [streams] [reps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from flvis_b200.pipeline import FrontendBench

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
f0, f1 = bench.make_streams(S, 0, 8)
fb = FrontendBench(S, bench.W, bench.H, bench.MAX_PTS, bench.NPTS, bench.FEATURE_PARA, 0, ba_window=10, kf_every=5)
fb.load_pool(f0, f1)
fb.reset()
for i in range(3):
    fb.step(i, "device")
torch.cuda.synchronize()
ctx = fb.ctx
st = fb.stream

def timed(name, fn):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    with torch.cuda.stream(st):
        fn()                       # warm
        ev[0].record(st)
        for _ in range(reps):
            fn()
        ev[1].record(st)
    torch.cuda.synchronize()
    print(f"{name:28s} {1e3 * ev[0].elapsed_time(ev[1]) / reps:10.1f} us")

prev0, cur0, cur1 = fb.slots
timed("upload_dev x2", lambda: (ctx.upload_dev(cur0, S, fb.d_pool0[1].data_ptr()), ctx.upload_dev(cur1, S, fb.d_pool1[1].data_ptr())))
timed("pyramid x2", lambda: (ctx.build_pyramid(cur0, S), ctx.build_pyramid(cur1, S)))
timed("lk f2f", lambda: fb._lk(prev0, cur0, fb.d_pts, fb.d_pts, fb.d_next, fb.d_status, fb.d_err, 10))
timed("select", lambda: ctx.select_tracked_dev(S, fb.d_npts.data_ptr(), fb.d_pts.data_ptr(), fb.d_next.data_ptr(), fb.d_status.data_ptr(), fb.d_keep.data_ptr(), fb.d_cur.data_ptr(), fb.d_cur64.data_ptr()))
timed("redetect (gftt+region)", lambda: ctx.feature_redetect_dev(cur0, S, fb.fp, fb.d_cur64.data_ptr(), fb.d_npts.data_ptr(), fb.d_new.data_ptr(), fb.d_nnew.data_ptr()))
timed("lk stereo", lambda: fb._lk(cur0, cur1, fb.d_cur, fb.d_cur, fb.d_right, fb.d_rstatus, fb.d_rerr, 5))
if fb.has_ba:
    with torch.cuda.stream(st):
        timed("ba (S/5 windows)", lambda: fb.ba.step(0, "device", 5))
timed("full step", lambda: fb._step(3, "device"))
print("tracked fraction f2f:", float(fb.d_keep[:, :].float().sum() / fb.d_npts.sum()))
