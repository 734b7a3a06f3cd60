"""Large-window bundle adjustment (26 .. 100 poses, flvis_b200/csrc/ba_big.cu): the SAME driver and phase code the GPU kernel runs,
executed sequentially on the host (flv_ba_big_emulate_host, a test aid), against the CPU oracle oracle/ba_ref.c.  Bars as in
tests/test_ba_gpu.py: culled-edge sets bit-exact, iteration counts equal, poses <= 1e-6, chi2 rel 1e-7.  Also checks that the
result does not depend on the number of emulated threads (the partition of every sum is fixed by design)."""
import ctypes as C

import numpy as np

from flvis_b200 import capi
from oracle import ba_ref
from synthdata import ba_problems

from .util import oracle_data


def _emulate(lib, p, threads, prm):
    lib.flv_ba_big_emulate_host.argtypes = [C.POINTER(capi.BAProblem), C.POINTER(capi.BAParams)] + [C.c_void_p] * 6 + [C.POINTER(capi.BAStats), C.c_int]
    poses = np.ascontiguousarray(p.poses, np.float64).copy(); lms = np.ascontiguousarray(p.lms, np.float64).copy()
    ep = np.ascontiguousarray(p.ep, np.int32); el = np.ascontiguousarray(p.el, np.int32); uv = np.ascontiguousarray(p.uv, np.float64)
    act = np.ones(len(ep), np.uint8)
    pb = capi.BAProblem(len(poses), len(lms), len(ep), p.fixed_pose, p.fix_landmarks, *p.K)
    st = capi.BAStats()
    vp = lambda a: a.ctypes.data_as(C.c_void_p)
    assert lib.flv_ba_big_emulate_host(C.byref(pb), C.byref(prm), vp(poses), vp(lms), vp(ep), vp(el), vp(uv), vp(act), C.byref(st), threads) == 0
    return poses, lms, act, st


def _rot_angle(q1, q2):
    return 2 * np.arccos(min(1.0, abs(float(np.dot(q1, q2)))))


def test_big_window_solver_matches_oracle_w30_and_is_partition_independent():
    lib = capi.load_library()
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0, 0)
    p = ba_problems.make_problem(window=30, n_landmarks=500, obs_per_frame=120, seed=5)
    poses, lms, act, st = _emulate(lib, p, 64, prm)
    d = oracle_data(p)
    ref = ba_ref.optimize(d, 12, 8, 1.0, 3.0, 0)
    assert st.ok == ref.ok == 1
    assert st.iterations_run == ref.iterations_run and st.n_culled == ref.n_culled and st.n_culled > 0
    assert np.array_equal(act, d.active)
    assert abs(st.chi2_initial - ref.chi2_initial) <= 1e-9 * ref.chi2_initial
    assert abs(st.chi2_final - ref.chi2_final) <= 1e-7 * ref.chi2_final
    assert np.abs(poses[:, 4:] - d.poses[:, 4:]).max() <= 1e-6
    assert max(_rot_angle(poses[i, :4], d.poses[i, :4]) for i in range(len(poses))) <= 1e-6
    assert np.abs(lms - d.lms).max() <= 1e-5
    poses2, lms2, act2, st2 = _emulate(lib, p, 2048, prm)                # the kernel's thread count
    assert np.array_equal(act, act2) and st2.iterations_run == st.iterations_run
    assert np.abs(poses - poses2).max() <= 1e-9 and np.abs(lms - lms2).max() <= 1e-8


def test_big_window_solver_small_window_and_pose_only():
    lib = capi.load_library()
    p = ba_problems.make_problem(window=6, n_landmarks=120, obs_per_frame=80, seed=7)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0, 0)
    poses, lms, act, st = _emulate(lib, p, 96, prm)
    d = oracle_data(p)
    ref = ba_ref.optimize(d, 12, 8, 1.0, 3.0, 0)
    assert st.iterations_run == ref.iterations_run and np.array_equal(act, d.active)
    assert np.abs(poses - d.poses).max() <= 1e-6 and np.abs(lms - d.lms).max() <= 1e-5


def test_big_window_solver_at_the_reference_limit_of_100_poses():
    """The reference's largest window (vo_localmap.cpp:441-447): 99 free poses, reduced camera system 594 x 594."""
    lib = capi.load_library()
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0, 0)
    p = ba_problems.make_problem(window=100, n_landmarks=1500, obs_per_frame=70, seed=9)
    assert len(p.poses) == 100
    poses, lms, act, st = _emulate(lib, p, 256, prm)
    d = oracle_data(p)
    ref = ba_ref.optimize(d, 12, 8, 1.0, 3.0, 0)
    assert st.ok == ref.ok == 1 and st.iterations_run == ref.iterations_run and st.n_culled == ref.n_culled
    assert np.array_equal(act, d.active)
    assert abs(st.chi2_final - ref.chi2_final) <= 1e-7 * ref.chi2_final
    assert np.abs(poses - d.poses).max() <= 1e-6 and np.abs(lms - d.lms).max() <= 1e-5
