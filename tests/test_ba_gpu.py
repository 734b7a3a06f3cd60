"""GPU parity tests of the device-side Levenberg-Marquardt / Schur bundle adjustment against the CPU oracle
(oracle/ba_ref.c, the restatement of the vendored g2o path).  fp64 with a different (tree) summation order,
so the bar is a stated tolerance: culled-edge sets bit-exact, poses <= 1e-6 m / 1e-6 rad, chi2 rel 1e-9."""
import numpy as np
import pytest

from oracle import ba_ref
from flvis_b200 import ba_batch, capi
from synthdata import ba_problems

from .util import oracle_data

pytestmark = pytest.mark.gpu


def _rot_angle(q1, q2):
    d = abs(float(np.dot(q1, q2)))
    return 2 * np.arccos(min(1.0, d))


def _check(batch, poses, lms, active, stats, prm):
    for s, p in enumerate(batch.problems):
        d = oracle_data(p)
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
    probs = [ba_problems.make_problem(window=6, n_landmarks=150, obs_per_frame=90, seed=1),
             ba_problems.make_problem(window=10, n_landmarks=1500, obs_per_frame=480, seed=2),
             ba_problems.make_problem(window=10, n_landmarks=400, obs_per_frame=160, seed=3, outlier_frac=0.2),
             ba_problems.make_problem(window=3, n_landmarks=60, obs_per_frame=60, seed=4)]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_large_windows_30_and_60_poses_global_memory_solver():
    """Windows beyond the shared-memory solver's 25 poses (the reference allows up to 100, vo_localmap.cpp:441-447) run on
    ba_big.cu: one cluster of 8 CTAs per window, reduced camera system in global memory.  Same bars as the small windows."""
    probs = [ba_problems.make_problem(window=30, n_landmarks=500, obs_per_frame=120, seed=5),
             ba_problems.make_problem(window=60, n_landmarks=1200, obs_per_frame=150, seed=6)]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_kitti_sized_window_20():
    probs = [ba_problems.make_problem(window=20, n_landmarks=2000, obs_per_frame=480, seed=11, w=1241, h=376,
                                   K=(718.856, 718.856, 607.1928, 185.2157))]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(1, 1241, 376)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_pose_only_ba_and_min_edge_failure():
    probs = [ba_problems.make_pose_only(300, seed=2), ba_problems.make_pose_only(480, seed=3, outlier_frac=0.3),
             ba_problems.make_pose_only(12, seed=4, outlier_frac=0.9)]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    prm = capi.BAParams(2, 2, 1.0, 3.0, 10)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, prm)
    _check(batch, poses, lms, active, stats, prm)
    assert stats[2].ok == 0
    ctx.close()


def test_zero_noise_known_answer_and_determinism():
    p = ba_problems.make_problem(window=6, n_landmarks=200, obs_per_frame=120, seed=3, noise_px=0.0, outlier_frac=0.0)
    batch = ba_batch.Batch([p, p])
    ctx = capi.Context(2, 752, 480)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch)
    assert stats[0].chi2_final < 1e-8 * stats[0].chi2_initial
    assert np.array_equal(poses[0], poses[1]) and np.array_equal(lms[0], lms[1])     # fixed summation order
    poses2, lms2, _, _ = ba_batch.solve_batch_host(ctx, batch)
    assert np.array_equal(poses, poses2) and np.array_equal(lms, lms2)               # run-to-run deterministic
    ctx.close()


def test_unsupported_window_is_reported():
    ctx = capi.Context(1, 752, 480)
    assert ctx.lib.flv_ba_reserve(ctx.h, 100, 100, 100) == 0             # the reference's largest window (vo_localmap.cpp:441-447)
    rc = ctx.lib.flv_ba_reserve(ctx.h, 101, 100, 100)
    assert rc == -4 and b"supports" in ctx.lib.flv_last_error(ctx.h)
    ctx.close()


def test_kernel_jacobians_match_central_differences_of_the_kernel_residual():
    """g2o's own Jacobian check (unit_test/test_helper/evaluate_jacobian.h:62-88: central differences, delta 1e-9 there;
    1e-6 here because the residual is O(100 px)) applied to the CUDA kernel's OWN edge function: analytic A, B from
    flv_ba_debug_edges against finite differences of the residual returned by the same device code, perturbing the pose
    with the oracle-independent exp map of oracle/ba_numpy.py."""
    from oracle import ba_numpy
    rng = np.random.default_rng(7)
    K = ba_problems.EUROC_K
    n = 64
    poses = np.zeros((n, 7)); X = np.zeros((n, 3)); uv = rng.uniform(0, 480, (n, 2))
    for i in range(n):
        aa = rng.normal(0, 0.4, 3); th = np.linalg.norm(aa)
        poses[i] = np.concatenate([np.sin(th / 2) * aa / th, [np.cos(th / 2)], rng.normal(0, 1, 3)])
        X[i] = (rng.normal(0, 1), rng.normal(0, 1), rng.uniform(4, 10))
    ctx = capi.Context(1, 752, 480)
    r, A, B = ctx.ba_debug_edges(poses, X, uv, K)
    d = 1e-6
    for k in range(3):
        e = np.zeros(3); e[k] = d
        rp, _, _ = ctx.ba_debug_edges(poses, X + e, uv, K); rm, _, _ = ctx.ba_debug_edges(poses, X - e, uv, K)
        fd = (rp - rm) / (2 * d)
        assert np.abs(fd - A[:, :, k]).max() <= 1e-6 * max(1.0, np.abs(A).max()) + 2e-4
    for k in range(6):
        e = np.zeros(6); e[k] = d
        pp = np.array([ba_numpy.pose_oplus(p, e) for p in poses]); pm = np.array([ba_numpy.pose_oplus(p, -e) for p in poses])
        rp, _, _ = ctx.ba_debug_edges(pp, X, uv, K); rm, _, _ = ctx.ba_debug_edges(pm, X, uv, K)
        fd = (rp - rm) / (2 * d)
        assert np.abs(fd - B[:, :, k]).max() <= 1e-6 * max(1.0, np.abs(B).max()) + 2e-4
    # and against both CPU derivations (g2o's table in ba_ref.c, the chain rule in ba_numpy)
    for i in range(n):
        rc, Ac, Bc = ba_ref.edge(poses[i], X[i], uv[i], K)
        rn, An, Bn = ba_numpy.edge_jacobians(poses[i], X[i], uv[i], K)
        assert np.abs(r[i] - rc).max() <= 1e-9 and np.abs(A[i] - Ac).max() <= 1e-9 * np.abs(Ac).max() and np.abs(B[i] - Bc).max() <= 1e-9 * np.abs(Bc).max()
        assert np.abs(A[i] - An).max() <= 1e-9 * np.abs(An).max() and np.abs(B[i] - Bn).max() <= 1e-9 * np.abs(Bn).max()
    ctx.close()


def _trace_close(tk, tr, rel_chi=1e-7, rel_lam=1e-4):
    assert len(tk) == len(tr), (len(tk), len(tr))
    for a, b in zip(tk, tr):
        assert abs(a[0] - b[0]) <= rel_chi * max(abs(b[0]), 1e-9), (a, b)
        assert abs(a[1] - b[1]) <= rel_lam * abs(b[1]), (a, b)
        assert int(a[3]) == int(b[3]), (a, b)


def test_kernel_lm_trace_matches_c_port_and_independent_numpy_oracle():
    """chi2 / lambda / trials after EVERY Levenberg-Marquardt iteration (optimization_algorithm_levenberg.cpp:58-150) of the
    device solver against (i) oracle/ba_ref.c and (ii) oracle/ba_numpy.py (dense normal equations, scipy Cholesky, no Schur)."""
    from oracle import ba_numpy
    probs = [ba_problems.make_problem(window=5, n_landmarks=120, obs_per_frame=70, seed=11),
             ba_problems.make_problem(window=8, n_landmarks=200, obs_per_frame=90, seed=12, outlier_frac=0.15),
             ba_problems.make_ba_demo(n_poses=8, n_points=120, seed=3, pixel_noise=1.0, outlier_ratio=0.05)]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(len(probs), 752, 480)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch)
    for s, p in enumerate(probs):
        tk = ctx.ba_trace(s, stats[s].iterations_run)
        d = oracle_data(p); tc = []
        st = ba_ref.optimize(d, 12, 8, trace=tc)
        assert stats[s].iterations_run == st.iterations_run
        _trace_close(tk, tc)
        tw = ba_numpy.DenseBA(p.poses.copy(), p.lms.copy(), p.ep, p.el, p.uv, p.K, p.fixed_pose, p.fix_landmarks)
        o = tw.optimize(12, 8)
        assert o["n_culled"] == stats[s].n_culled and np.array_equal(tw.active, active[s, :len(p.ep)])
        _trace_close(tk, tw.trace, rel_chi=1e-6)
        P = len(p.poses)
        assert np.abs(poses[s, :P, 4:] - tw.poses[:, 4:]).max() <= 1e-6
    ctx.close()


def test_ba_demo_fixture_on_the_gpu():
    """3rdPartLib/g2o/g2o/examples/ba/ba_demo.cpp:126-293: 15 poses, 500 points, f = 1000, 640x480."""
    p0 = ba_problems.make_ba_demo(seed=1, pixel_noise=0.0)
    p1 = ba_problems.make_ba_demo(seed=2, pixel_noise=1.0, outlier_ratio=0.05)
    batch = ba_batch.Batch([p0, p1])
    ctx = capi.Context(2, 640, 480)
    prm = capi.BAParams(12, 8, 1.0, 3.0, 0)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, prm)
    assert stats[0].ok and stats[0].chi2_final < 1e-12 * stats[0].chi2_initial          # zero noise: exact
    assert np.abs(lms[0, :len(p0.lms)] - p0.gt[1]).max() < 1e-6 or stats[0].chi2_final < 1e-9
    assert stats[1].ok and stats[1].chi2_final < 0.01 * stats[1].chi2_initial
    _check(batch, poses, lms, active, stats, prm)
    ctx.close()


def test_reserve_is_idempotent_and_grows():
    """flv_ba_reserve is called per frame by the host tracker: covered requests must not reallocate."""
    ctx = capi.Context(1, 752, 480)
    assert ctx.lib.flv_ba_reserve(ctx.h, 1, 512, 512) == 0
    ws = ctx.lib.flv_ba_reserve
    import time
    t0 = time.perf_counter()
    for _ in range(200):
        assert ctx.lib.flv_ba_reserve(ctx.h, 1, 512, 512) == 0
    assert (time.perf_counter() - t0) / 200 < 50e-6                                       # no cudaFree / cudaMalloc inside
    assert ctx.lib.flv_ba_reserve(ctx.h, 10, 1500, 4800) == 0                             # grow
    p = ba_problems.make_pose_only(300, seed=2)
    batch = ba_batch.Batch([p])
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch, capi.BAParams(2, 2, 1.0, 3.0, 10))
    assert stats[0].ok == 1
    ctx.close()
