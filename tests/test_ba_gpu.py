"""GPU parity tests of the device-side Levenberg-Marquardt / Schur bundle adjustment against the CPU oracle
(oracle/ba_ref.c, the restatement of the vendored g2o path).  fp64 with a different (tree) summation order,
so the bar is a stated tolerance: culled-edge sets bit-exact, poses <= 1e-6 m / 1e-6 rad, chi2 rel 1e-9."""
import numpy as np
import pytest

from oracle import ba_ref
from flvis_b200 import ba_synth, capi

pytestmark = pytest.mark.gpu


def _rot_angle(q1, q2):
    d = abs(float(np.dot(q1, q2)))
    return 2 * np.arccos(min(1.0, d))


def _check(batch, poses, lms, active, stats, prm):
    for s, p in enumerate(batch.problems):
        d = p.oracle_data()
        st = ba_ref.optimize(d, prm.iters1, prm.iters2, prm.huber_delta, prm.cull_chi2, prm.min_edges_after_cull)
        g = stats[s]
        P, L, E = len(p.poses), len(p.lms), len(p.ep)
        assert g.ok == st.ok
        assert g.iterations_run == st.iterations_run, (s, g.iterations_run, st.iterations_run)
        assert g.n_culled == st.n_culled
        assert np.array_equal(active[s, :E], d.active)                      # outlier edge set: bit-exact
        assert abs(g.chi2_initial - st.chi2_initial) <= 1e-9 * st.chi2_initial
        if st.ok:
            assert abs(g.chi2_final - st.chi2_final) <= 1e-7 * max(st.chi2_final, 1e-12) + 1e-9
            assert np.abs(poses[s, :P, 4:] - d.poses[:, 4:]).max() <= 1e-6
            assert max(_rot_angle(poses[s, i, :4], d.poses[i, :4]) for i in range(P)) <= 1e-6
            if not p.fix_landmarks:
                assert np.abs(lms[s, :L] - d.lms).max() <= 1e-5


def test_local_ba_matches_oracle_small_and_euroc_sized():
    probs = [ba_synth.make_problem(window=6, n_landmarks=150, obs_per_frame=90, seed=1),
             ba_synth.make_problem(window=10, n_landmarks=1500, obs_per_frame=480, seed=2),
             ba_synth.make_problem(window=10, n_landmarks=400, obs_per_frame=160, seed=3, outlier_frac=0.2),
             ba_synth.make_problem(window=3, n_landmarks=60, obs_per_frame=60, seed=4)]
    batch = ba_synth.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_synth.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_kitti_sized_window_20():
    probs = [ba_synth.make_problem(window=20, n_landmarks=2000, obs_per_frame=480, seed=11, w=1241, h=376,
                                   K=(718.856, 718.856, 607.1928, 185.2157))]
    batch = ba_synth.Batch(probs)
    ctx = capi.Context(1, 1241, 376)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_synth.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_pose_only_ba_and_min_edge_failure():
    probs = [ba_synth.make_pose_only(300, seed=2), ba_synth.make_pose_only(480, seed=3, outlier_frac=0.3),
             ba_synth.make_pose_only(12, seed=4, outlier_frac=0.9)]
    batch = ba_synth.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    prm = capi.BAParams(2, 2, 1.0, 3.0, 10)
    poses, lms, active, stats = ba_synth.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    assert stats[2].ok == 0
    ctx.close()


def test_zero_noise_known_answer_and_determinism():
    p = ba_synth.make_problem(window=6, n_landmarks=200, obs_per_frame=120, seed=3, noise_px=0.0, outlier_frac=0.0)
    batch = ba_synth.Batch([p, p])
    ctx = capi.Context(2, 752, 480)
    poses, lms, active, stats = ba_synth.solve_batch_host(ctx, batch)
    assert stats[0].chi2_final < 1e-8 * stats[0].chi2_initial
    assert np.array_equal(poses[0], poses[1]) and np.array_equal(lms[0], lms[1])     # fixed summation order
    poses2, lms2, _, _ = ba_synth.solve_batch_host(ctx, batch)
    assert np.array_equal(poses, poses2) and np.array_equal(lms, lms2)               # run-to-run deterministic
    ctx.close()


def test_unsupported_window_is_reported():
    ctx = capi.Context(1, 752, 480)
    rc = ctx.lib.flv_ba_reserve(ctx.h, 40, 100, 100)
    assert rc == -4 and b"supports" in ctx.lib.flv_last_error(ctx.h)
    ctx.close()
