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
@pytest.mark.parametrize("window,n_kf", [(5, 12), (10, 16), (30, 34)])     # 30: the large-window solver (ba_big.cu)
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


@pytest.mark.gpu
def test_localmap_batch_worker_matches_oracle_per_stream(lib):
    """flv_localmap_batch: one worker thread, S LocalMap state machines, all due windows of a submission in ONE launch.
    Three sequences of different length per keyframe; every CorrectionInf against the oracle of its own sequence."""
    K = (458.654, 457.296, 367.215, 248.375)
    window, n_kf, S = 5, 11, 3
    seqs = [localmap_ref.make_keyframe_sequence(n_kf, seed=30 + s, n_per_kf=120 + 40 * s) for s in range(S)]
    refs = [localmap_ref.LocalMap(window, K) for _ in range(S)]
    vp = C.c_void_p
    lib.flv_localmap_batch_create.restype = vp
    lib.flv_localmap_batch_create.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_double] * 4
    lib.flv_localmap_batch_destroy.argtypes = [vp]
    lib.flv_localmap_batch_submit.argtypes = [vp, C.c_int] + [vp] * 7
    lib.flv_localmap_batch_wait.argtypes = [vp]
    lib.flv_localmap_batch_result.argtypes = [vp, C.c_int, vp, vp, vp, vp, vp, C.c_int, vp, vp, C.c_int]
    lib.flv_localmap_batch_stats.argtypes = [vp] * 5
    lib.flv_localmap_batch_last_error.restype = C.c_char_p
    lib.flv_localmap_batch_last_error.argtypes = [vp]
    b = lib.flv_localmap_batch_create(0, S, window, *K)
    assert b
    p = lambda a: a.ctypes.data_as(vp)
    cap = 8192
    n_checked = 0
    for k in range(n_kf):
        # stream 2 skips every third keyframe: the streams' windows are not in phase
        sel = [s for s in range(S) if not (s == 2 and k % 3 == 1)]
        kfs = [seqs[s][k] for s in sel]
        streams = np.array(sel, np.int32); fids = np.array([kf["frame_id"] for kf in kfs], np.int64)
        cnt = np.array([len(kf["lm_id"]) for kf in kfs], np.int32)
        ids = np.ascontiguousarray(np.concatenate([np.asarray(kf["lm_id"], np.int64) for kf in kfs]))
        uv = np.ascontiguousarray(np.concatenate([np.asarray(kf["lm_2d"], np.float64).reshape(-1, 2) for kf in kfs]))
        p3 = np.ascontiguousarray(np.concatenate([np.asarray(kf["lm_3d"], np.float64).reshape(-1, 3) for kf in kfs]))
        T = np.ascontiguousarray(np.stack([np.asarray(kf["T_c_w"], np.float64) for kf in kfs]))
        assert lib.flv_localmap_batch_submit(b, len(sel), p(streams), p(fids), p(cnt), p(ids), p(uv), p(p3), p(T)) == 0
        assert lib.flv_localmap_batch_wait(b) == 0, lib.flv_localmap_batch_last_error(b)
        for s, kf in zip(sel, kfs):
            o = refs[s].frame_callback(kf)
            fid = np.zeros(1, np.int64); oT = np.zeros(7); nlm = np.zeros(1, np.int32); olm = np.zeros(cap, np.int64)
            o3 = np.zeros((cap, 3)); nout = np.zeros(1, np.int32); oout = np.zeros(cap, np.int64)
            rc = lib.flv_localmap_batch_result(b, s, p(fid), p(oT), p(nlm), p(olm), p(o3), cap, p(nout), p(oout), cap)
            if o is None:
                assert rc == 0
                continue
            assert rc >= 1 and fid[0] == o["frame_id"] and list(olm[:nlm[0]]) == o["lm_id"]
            assert sorted(oout[:nout[0]]) == sorted(o["outlier_id"])
            assert np.abs(oT[4:] - o["T_c_w"][4:]).max() <= 1e-6
            if nlm[0]:
                assert np.abs(o3[:nlm[0]] - o["lm_3d"]).max() <= 1e-5
            n_checked += 1
    nk = C.c_longlong(); ns = C.c_longlong(); nl = C.c_longlong(); ms = (C.c_double * 2)()
    lib.flv_localmap_batch_stats(b, C.byref(nk), C.byref(ns), C.byref(nl), ms)
    assert ns.value == n_checked >= 15 and nl.value < ns.value          # several windows per launch
    lib.flv_localmap_batch_destroy(b)
