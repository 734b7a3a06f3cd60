"""VIMOTION host class (C++, through the C handles) against the Python restatement of vi_motion.cpp.
Pure host code: runs in the CPU suite (the shared library loads without a GPU)."""
import ctypes as C

import numpy as np

from oracle import vimotion_ref as vr


def _lib(lib):
    lib.flv_vimotion_create.restype = C.c_void_p
    lib.flv_vimotion_create.argtypes = [C.c_void_p] + [C.c_double] * 7
    lib.flv_vimotion_destroy.argtypes = [C.c_void_p]
    lib.flv_vimotion_imu_feed.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 5
    lib.flv_vimotion_vision_trigger.argtypes = [C.c_void_p, C.c_void_p]
    lib.flv_vimotion_correction.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_void_p]
    lib.flv_vimotion_corr_frame_state.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    lib.flv_vimotion_rp_compensation.argtypes = [C.c_void_p, C.c_double, C.c_void_p]
    lib.flv_vimotion_get_bias.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    lib.flv_vimotion_queue_size.argtypes = [C.c_void_p]
    return lib


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


def test_vimotion_matches_oracle_sample_by_sample(lib):
    lib = _lib(lib)
    # EuRoC-like camera-IMU extrinsics (rotation about z by ~90 deg + small offset)
    T_i_c7 = np.array([0.0, 0.0, 0.7071067811865476, 0.7071067811865476, -0.02, -0.06, 0.01])
    para = (0.1, 0.01, 0.001, 0.001, 0.5, 0.1)          # euroc.yaml vifusion_para1..4 + default saturations
    h = lib.flv_vimotion_create(vp(T_i_c7), 9.81, *para)
    ref = vr.VIMOTION(vr.SE3.from7(T_i_c7), 9.81, *para)
    t, acc, gyro = vr.synth_imu(900, seed=3)
    q = np.zeros(4); p = np.zeros(3); v = np.zeros(3)
    triggered = False
    last_vis = None
    for i in range(len(t)):
        a = np.ascontiguousarray(acc[i]); g = np.ascontiguousarray(gyro[i])
        rc = lib.flv_vimotion_imu_feed(h, float(t[i]), vp(a), vp(g), vp(q), vp(p), vp(v))
        rq, rp, rv = ref.imu_feed(float(t[i]), acc[i], gyro[i])
        assert rc == (1 if ref.imu_initialized else 0)
        assert np.array_equal(q, rq) and np.array_equal(p, rp) and np.array_equal(v, rv)
        if ref.imu_initialized and not triggered and i > 60:
            qt = np.zeros(4)
            assert lib.flv_vimotion_vision_trigger(h, vp(qt)) == 0
            assert np.array_equal(qt, ref.vision_trigger())
            assert lib.flv_vimotion_queue_size(h) == 1
            triggered = True
        # a "vision" pose every 10 samples (20 Hz): the IMU-predicted camera pose, nudged
        if triggered and i % 10 == 0 and i > 100:
            T = np.zeros(7)
            found = lib.flv_vimotion_corr_frame_state(h, float(t[i]) + 1e-4, vp(T))
            rT = ref.corr_frame_state(float(t[i]) + 1e-4)
            assert found == (1 if rT is not None else 0)
            if rT is None:
                continue
            assert np.array_equal(T, rT.to7())
            vis = rT.to7().copy(); vis[4:] += np.array([0.002, -0.001, 0.0015]) * np.sin(i)
            T2 = vis.copy()
            lib.flv_vimotion_rp_compensation(h, float(t[i]) + 1e-4, vp(T2))
            r2 = ref.rp_compensation(float(t[i]) + 1e-4, vr.SE3.from7(vis))
            assert np.array_equal(T2, r2.to7())
            if last_vis is not None:
                lib.flv_vimotion_correction(h, float(t[i]) + 1e-4, vp(T2), last_vis[0], vp(last_vis[1]))
                ref.correction_from_vision(float(t[i]) + 1e-4, vr.SE3.from7(T2), last_vis[0], vr.SE3.from7(last_vis[1]))
                ab = np.zeros(3); gb = np.zeros(3)
                lib.flv_vimotion_get_bias(h, vp(ab), vp(gb))
                assert np.abs(ab - ref.acc_bias).max() < 1e-10 and np.abs(gb - ref.gyro_bias).max() < 1e-10
            last_vis = (float(t[i]) + 1e-4, T2.copy())
    assert triggered and last_vis is not None
    assert np.abs(ref.acc_bias).max() > 0                      # the bias feedback path was exercised
    lib.flv_vimotion_destroy(h)


def test_vimotion_is_safe_under_concurrent_imu_and_vision_threads(lib):
    """imu_feed runs on other threads than image_feed in the reference (ROS callback threads, vo_tracking.cpp:326-371), so
    VIMOTION guards `states` with mtx_states_RW (vi_motion.cpp:119-205, :220-339, :390-433).  ctypes releases the GIL: the
    feeder thread and the vision-side calls really run concurrently here."""
    import threading
    lib = _lib(lib)
    T_i_c7 = np.array([0.0, 0.0, 0.7071067811865476, 0.7071067811865476, -0.02, -0.06, 0.01])
    h = lib.flv_vimotion_create(vp(T_i_c7), 9.81, 0.1, 0.01, 0.001, 0.001, 0.5, 0.1)
    t, acc, gyro = vr.synth_imu(6000, seed=5)
    q = np.zeros(4); p = np.zeros(3); v = np.zeros(3)
    for i in range(60):
        lib.flv_vimotion_imu_feed(h, float(t[i]), vp(np.ascontiguousarray(acc[i])), vp(np.ascontiguousarray(gyro[i])), vp(q), vp(p), vp(v))
    qt = np.zeros(4)
    assert lib.flv_vimotion_vision_trigger(h, vp(qt)) == 0
    now = [60]
    stop = threading.Event()

    def feeder():
        import time
        q2 = np.zeros(4); p2 = np.zeros(3); v2 = np.zeros(3)
        for i in range(60, len(t)):
            a = np.ascontiguousarray(acc[i]); g = np.ascontiguousarray(gyro[i])
            lib.flv_vimotion_imu_feed(h, float(t[i]), vp(a), vp(g), vp(q2), vp(p2), vp(v2))
            now[0] = i
            if i % 25 == 0:
                time.sleep(0.0005)          # paced like a sensor: the vision thread gets its turns whatever the machine load
        stop.set()

    th = threading.Thread(target=feeder)
    th.start()
    n_found = 0
    last = None
    while not stop.is_set():
        i = now[0]
        T = np.zeros(7)
        if lib.flv_vimotion_corr_frame_state(h, float(t[max(i - 3, 0)]), vp(T)) == 1:
            n_found += 1
            assert np.all(np.isfinite(T)) and abs(np.linalg.norm(T[:4]) - 1) < 1e-9
            lib.flv_vimotion_rp_compensation(h, float(t[max(i - 3, 0)]), vp(T))
            if last is not None and t[max(i - 3, 0)] > last[0]:
                lib.flv_vimotion_correction(h, float(t[max(i - 3, 0)]), vp(T), last[0], vp(last[1]))
            last = (float(t[max(i - 3, 0)]), T.copy())
        assert 0 < lib.flv_vimotion_queue_size(h) <= 400
    th.join()
    assert n_found > 10
    ab = np.zeros(3); gb = np.zeros(3)
    lib.flv_vimotion_get_bias(h, vp(ab), vp(gb))
    assert np.all(np.isfinite(ab)) and np.all(np.isfinite(gb))
    lib.flv_vimotion_destroy(h)
