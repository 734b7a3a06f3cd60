"""LK stage bench (run on the GPU box): 32 streams x 480 points, frame->frame call, device-resident arguments, CUDA events.
FLV_LK_VARIANT selects the kernel (6 = v4 default, 7 = v4 with the second image's patch staged by TMA); prints the median
time per launch and checks that the result equals variant 6's bit for bit when FLV_LK_CHECK points at a saved result."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from flvis_b200 import capi
from synthdata import textures

S = int(sys.argv[1]) if len(sys.argv) > 1 else 32
N = 480
h, w = 480, 752
ctx = capi.Context(S, w, h, 512)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
ctx.set_stream(stream.cuda_stream)
Is, Js = [], []
for s in range(min(S, 4)):
    I, J, _ = textures.frame_pair(10 + s, h, w, (4.2 + s, -2.7), 0.5, 1.0)
    Is.append(I); Js.append(J)
ctx.upload(0, np.stack([Is[s % len(Is)] for s in range(S)])); ctx.upload(1, np.stack([Js[s % len(Js)] for s in range(S)]))
ctx.build_pyramid(0, S); ctx.build_pyramid(1, S)
pts = np.zeros((S, 512, 2), np.float32)
for s in range(S):
    c = ctx.gftt(0, S, N, 0.01, 10)[s] if s < len(Is) else None
    if c is not None:
        base = c
    pts[s, :len(base)] = base[:N]
n = np.full(S, min(N, len(base)), np.int32)
dev = "cuda"
d_n = torch.from_numpy(n).to(dev); d_prev = torch.from_numpy(pts).to(dev); d_next = torch.empty_like(d_prev)
d_st = torch.empty((S, 512), dtype=torch.uint8, device=dev); d_err = torch.empty((S, 512), dtype=torch.float32, device=dev)
def run():
    ctx.lk_track_dev(0, 1, S, d_n.data_ptr(), d_prev.data_ptr(), d_prev.data_ptr(), d_next.data_ptr(), d_st.data_ptr(), d_err.data_ptr())
for _ in range(5):
    run()
torch.cuda.synchronize()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
ts = []
for _ in range(15):
    flush.zero_()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); run(); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1) * 1e3)
res = np.concatenate([d_next.cpu().numpy().view(np.uint32).ravel(), d_st.cpu().numpy().ravel().astype(np.uint32)])
tag = os.environ.get("FLV_LK_VARIANT", "7")
out = os.environ.get("FLV_LK_CHECK")
same = None
if out:
    if os.path.exists(out):
        same = bool(np.array_equal(np.load(out), res))
    else:
        np.save(out, res)
print(f"LK variant {tag}: S={S} pts={int(n[0])} tracked={int(d_st.sum())} median {np.median(ts):.1f} us  min {min(ts):.1f}  max {max(ts):.1f}  identical_to_saved={same}")
