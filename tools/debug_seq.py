"""Debug aid (GPU box): run a config sequence, on the first LK-position mismatch dump the oracle's LK call and the GPU's."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import numpy as np
from flvis_b200 import capi
from synthdata import sequences
from tests import seq_harness as sh
from oracle import lk_ref

name = sys.argv[1]; n = int(sys.argv[2])
lib = capi.load_library()
seq = getattr(sequences, "make_" + name)(n)
h, ref, K = sh.make_pair(lib, seq, True)
# wrap the oracle LK to record its last call
calls = []
orig = lk_ref.calc_optical_flow_pyr_lk_c
def rec(I, J, p, q, **kw):
    out = orig(I, J, p, q, **kw)
    calls.append((I.copy(), J.copy(), np.array(p, np.float32).copy(), np.array(q, np.float32).copy(), kw, out))
    return out
lk_ref.calc_optical_flow_pyr_lk_c = rec
vp = sh._vp
for k, (t, img0, img1, imu) in enumerate(seq.frames()):
    for (ti, acc, gyro) in imu:
        a = np.ascontiguousarray(acc); g = np.ascontiguousarray(gyro)
        lib.flv_f2f_imu_feed(h, float(ti), vp(a), vp(g)); ref.imu_feed(float(ti), acc, gyro)
    calls.clear()
    kf = C.c_int(0); rs = C.c_int(0)
    lib.flv_f2f_image_feed(h, float(t), vp(np.ascontiguousarray(img0)), vp(np.ascontiguousarray(img1)), C.byref(kf), C.byref(rs))
    ref.image_feed(float(t), img0, img1)
    cap = 600
    T = np.zeros(7); ids = np.zeros(cap, np.int64); pl = np.zeros((cap, 2))
    nn = lib.flv_f2f_get_frame(h, vp(T), vp(ids), vp(pl), None, None, None, None, cap)
    cur = ref.curr
    rpl = np.array([l.plane for l in cur.lms]).reshape(-1, 2)
    print(k, sh.STATE[lib.flv_f2f_state(h)], ref.state, nn, len(cur.lms), "pose d", np.abs(T - cur.T_c_w.to7()).max())
    if nn != len(cur.lms) or (nn and np.abs(pl[:nn] - rpl).max() > 1e-3):
        d = np.abs(pl[:nn] - rpl).max(axis=1) if nn == len(cur.lms) else None
        print("MISMATCH at frame", k, "n", nn, len(cur.lms))
        if d is not None:
            bad = np.nonzero(d > 1e-3)[0]
            print("bad idx", bad[:10], "gpu", pl[bad[:5]], "ref", rpl[bad[:5]], "ids", ids[bad[:5]])
        # replay the oracle's f2f LK call on the GPU kernel
        I, J, p, q, kw, out = calls[0]
        ctx = capi.Context(1, I.shape[1], I.shape[0], 512)
        if ref.equalize: pass
        ctx.upload(0, I); ctx.upload(1, J); ctx.build_pyramid(0, 1); ctx.build_pyramid(1, 1)
        nxt, st, err = ctx.lk_track(0, 1, p, q, max_level=kw.get("max_level", 10))
        dd = np.abs(nxt - out[0]).max(axis=1)
        print("replay: status equal", np.array_equal(st, out[1]), "max d", dd.max(), "n bad", int((dd > 1e-3).sum()))
        b = np.nonzero(dd > 1e-3)[0][:8]
        for i in b:
            print("  pt", i, "prev", p[i], "init", q[i], "gpu", nxt[i], st[i], "ref", out[0][i], out[1][i])
        pyo = lk_ref.calc_optical_flow_pyr_lk(I, J, p[b], q[b], max_level=kw.get("max_level", 10))
        print("  python oracle on the bad points:", pyo[0], pyo[1])
        np.savez("gpurun_out/lk_mismatch.npz", I=I, J=J, p=p, q=q, gpu=nxt, gst=st, ref=out[0], rst=out[1])
        break
