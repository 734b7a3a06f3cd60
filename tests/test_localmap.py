"""LocalMap (C++ host class over the GPU BA) against the Python restatement of vo_localmap.cpp + poselmbag.cpp.
CPU part: PoseLMBag/LocalMap oracle self-consistency.  GPU part: keyframe-by-keyframe CorrectionInf parity."""
import ctypes as C

import numpy as np
import pytest

from oracle import localmap_ref


def test_oracle_localmap_runs_and_slides():
    kfs = localmap_ref.make_keyframe_sequence(9, seed=1, n_per_kf=80)
    lm = localmap_ref.LocalMap(5, (458.654, 457.296, 367.215, 248.375))
    outs = [lm.frame_callback(k) for k in kfs]
    assert outs[:4] == [None] * 4 and all(o is not None for o in outs[4:])
    assert outs[4]["frame_id"] == kfs[4]["frame_id"]
    # landmarks reported are those seen >= 4 times, in bag order
    assert all(len(o["lm_id"]) == len(o["lm_3d"]) for o in outs[4:])
    # bag slots cycle: newest slot of the 6th keyframe is slot 0 (the oldest one was overwritten)
    assert lm.bag.newest == (len(kfs) - 5 - 1) % 5


@pytest.mark.gpu
@pytest.mark.parametrize("window,n_kf", [(5, 12), (10, 16)])
def test_localmap_matches_oracle_keyframe_by_keyframe(window, n_kf):
    from flvis_b200 import capi
    K = (458.654, 457.296, 367.215, 248.375)
    kfs = localmap_ref.make_keyframe_sequence(n_kf, seed=window, n_per_kf=200)
    ref = localmap_ref.LocalMap(window, K)
    ctx = capi.Context(1, 752, 480)
    lib = ctx.lib
    lib.flv_localmap_create.restype = C.c_void_p
    lib.flv_localmap_create.argtypes = [C.c_void_p, C.c_int] + [C.c_double] * 4
    lib.flv_localmap_destroy.argtypes = [C.c_void_p]
    lib.flv_localmap_add_keyframe.argtypes = [C.c_void_p, C.c_int64, C.c_int] + [C.c_void_p] * 4 + [C.c_void_p] * 5 + \
        [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(capi.BAStats)]
    h = lib.flv_localmap_create(ctx.h, window, *K)
    assert h
    cap = 8192
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    for k, kf in enumerate(kfs):
        o = ref.frame_callback(kf)
        ids = np.ascontiguousarray(kf["lm_id"], np.int64); uv = np.ascontiguousarray(kf["lm_2d"], np.float64)
        p3 = np.ascontiguousarray(kf["lm_3d"], np.float64); T = np.ascontiguousarray(kf["T_c_w"], np.float64)
        fid = np.zeros(1, np.int64); oT = np.zeros(7); nlm = np.zeros(1, np.int32); olm = np.zeros(cap, np.int64)
        o3 = np.zeros((cap, 3)); nout = np.zeros(1, np.int32); oout = np.zeros(cap, np.int64)
        st = capi.BAStats()
        rc = lib.flv_localmap_add_keyframe(h, int(kf["frame_id"]), len(ids), vp(ids), vp(uv), vp(p3), vp(T), vp(fid), vp(oT),
                                           vp(nlm), vp(olm), vp(o3), cap, vp(nout), vp(oout), cap, C.byref(st))
        if o is None:
            assert rc == 0
            continue
        assert rc == 1, (k, rc, lib.flv_last_error(ctx.h))
        assert fid[0] == o["frame_id"]
        assert list(olm[:nlm[0]]) == o["lm_id"]                                   # landmark id list: bit-exact, same order
        assert sorted(oout[:nout[0]]) == sorted(o["outlier_id"])                  # outlier ids: as a multiset (A.5)
        assert st.iterations_run == o["stats"].iterations_run
        assert np.abs(oT[4:] - o["T_c_w"][4:]).max() <= 1e-6                      # per-keyframe pose: 1e-6 m
        assert 2 * np.arccos(min(1.0, abs(float(np.dot(oT[:4], o["T_c_w"][:4]))))) <= 1e-6
        if nlm[0]:
            assert np.abs(o3[:nlm[0]] - o["lm_3d"]).max() <= 1e-5
    lib.flv_localmap_destroy(h)
    ctx.close()
