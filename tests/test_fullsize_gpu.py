"""Size-independent properties at BASELINE.json's full configuration (32 concurrent 752x480 stereo streams, 480 points
per stream, W=10 windows) -- sizes the CPU oracle cannot finish in seconds, so parity is argued through invariants of the
domain, plus spot checks of a few streams against the oracle."""
import numpy as np
import pytest

from oracle import lk_ref, gftt_ref
from synthdata import textures as synth

pytestmark = pytest.mark.gpu

S, W, H, NPTS = 32, 752, 480, 480


@pytest.fixture(scope="module")
def world():
    from flvis_b200 import capi
    rng = np.random.default_rng(11)
    base = np.stack([synth.texture(100 + s, H + 32, W + 32, blur=2) for s in range(S)])
    I = np.ascontiguousarray(base[:, 8:8 + H, 8:8 + W])
    shifts = [(int(rng.integers(-6, 7)), int(rng.integers(-6, 7))) for _ in range(S)]
    J = np.stack([base[s, 8 + dy:8 + dy + H, 8 + dx:8 + dx + W] for s, (dx, dy) in enumerate(shifts)])   # J(x,y) = I(x+dx, y+dy)
    pts = np.zeros((S, 512, 2), np.float32)
    pts[:, :NPTS, 0] = rng.uniform(40, W - 40, (S, NPTS)); pts[:, :NPTS, 1] = rng.uniform(40, H - 40, (S, NPTS))
    ctx = capi.Context(S, W, H, 512)
    ctx.upload(0, I); ctx.upload(1, np.ascontiguousarray(J))
    ctx.build_pyramid(0, S); ctx.build_pyramid(1, S)
    yield ctx, I, J, pts, shifts
    ctx.close()


def test_lk_identity_is_a_fixed_point(world):
    """Tracking an image against itself: the residual of every window is exactly zero, so every point converges in the
    first iteration with zero motion (the returned position is (p - 15) + 15 in float: within two ulps of p)."""
    ctx, I, J, pts, _ = world
    n = np.full(S, NPTS, np.int32)
    nxt, st, err = ctx.lk_track(0, 0, pts, pts, n_pts=n)
    ok = st[:, :NPTS] == 1
    assert ok.mean() > 0.97                                   # the rest fail the min-eigenvalue test (flat texture)
    assert np.abs(nxt[:, :NPTS][ok] - pts[:, :NPTS][ok]).max() <= 2e-4
    assert err[:, :NPTS][ok].max() < 0.05      # the error pass re-blends at the (<= 1 ulp) moved position


def test_lk_recovers_integer_translations_and_matches_oracle_on_samples(world):
    ctx, I, J, pts, shifts = world
    n = np.full(S, NPTS, np.int32)
    nxt, st, err = ctx.lk_track(0, 1, pts, pts, n_pts=n)
    for s, (dx, dy) in enumerate(shifts):
        m = st[s, :NPTS] == 1
        assert m.mean() > 0.9
        d = nxt[s, :NPTS][m] - pts[s, :NPTS][m]
        # the scene point seen at (x, y) in I sits at (x - dx, y - dy) in J
        assert np.abs(np.median(d[:, 0]) + dx) < 0.02 and np.abs(np.median(d[:, 1]) + dy) < 0.02
        assert (np.abs(d + np.array([dx, dy])) < 0.1).mean() > 0.98
    for s in (0, 17, 31):                                       # bit-exact spot checks
        o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I[s], J[s], pts[s, :64], pts[s, :64], max_level=10)
        assert np.array_equal(st[s, :64], o_st)
        assert np.array_equal(nxt[s, :64].view(np.uint32), o_nxt.view(np.uint32))


def test_lk_is_independent_of_batch_position(world):
    """A stream's result does not depend on which slot of the batch it occupies (no cross-stream state)."""
    from flvis_b200 import capi
    ctx, I, J, pts, _ = world
    n = np.full(S, NPTS, np.int32)
    nxt, st, err = ctx.lk_track(0, 1, pts, pts, n_pts=n)
    perm = np.random.default_rng(2).permutation(S)
    c2 = capi.Context(S, W, H, 512)
    c2.upload(0, np.ascontiguousarray(I[perm])); c2.upload(1, np.ascontiguousarray(J[perm]))
    c2.build_pyramid(0, S); c2.build_pyramid(1, S)
    nxt2, st2, err2 = c2.lk_track(0, 1, np.ascontiguousarray(pts[perm]), np.ascontiguousarray(pts[perm]), n_pts=n)
    assert np.array_equal(st2[:, :NPTS], st[perm][:, :NPTS])
    assert np.array_equal(nxt2[:, :NPTS].view(np.uint32), nxt[perm][:, :NPTS].view(np.uint32))
    c2.close()


def test_gftt_invariants_all_streams(world):
    """goodFeaturesToTrack post-conditions for all 32 streams: descending response, response >= q * max, pairwise distance
    >= d, every corner a strict 3x3 local maximum candidate; two streams bit-exact against the oracle."""
    ctx, I, J, pts, _ = world
    N, q, d = 1000, 0.01, 10
    corners = ctx.gftt(0, S, N, q, d)
    for s in (3, 29):
        assert np.array_equal(corners[s], gftt_ref.good_features_to_track(I[s], N, q, d))
    ctx.keep_response(True)
    ctx.gftt(0, S, N, q, d)
    for s in range(0, S, 5):
        eig = ctx.download_eig(s)
        c = corners[s].astype(int)
        assert 0 < len(c) <= N
        v = eig[c[:, 1], c[:, 0]]
        assert np.all(np.diff(v) <= 0)
        assert v.min() >= np.float32(eig.max() * q) * (1 - 1e-6)
        dd = np.linalg.norm(c[:, None, :] - c[None, :, :], axis=2) + np.eye(len(c)) * 1e9
        assert dd.min() >= d
    ctx.keep_response(False)


def test_ba_full_batch_reduces_chi2_and_recovers_poses():
    """32 EuRoC-sized windows (W=10, ~4.6k edges) in one launch: chi2 drops by > 10x, the culled fraction matches the
    noise model, pose errors are at the centimetre level, and stream 0 matches the fp64 oracle."""
    from flvis_b200 import capi, ba_batch
    from oracle import ba_ref
    from synthdata import ba_problems
    from .util import oracle_data
    probs = [ba_problems.make_problem(window=10, n_landmarks=1500, obs_per_frame=480, seed=50 + s) for s in range(S)]
    batch = ba_batch.Batch(probs)
    ctx = capi.Context(S, W, H)
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch)
    for s in range(S):
        assert stats[s].ok == 1 and stats[s].iterations_run == 20
        assert stats[s].chi2_final < 0.1 * stats[s].chi2_initial
        frac = stats[s].n_culled / float(len(probs[s].ep))
        assert 0.08 < frac < 0.4
        gt_poses = probs[s].gt[0]
        assert np.abs(poses[s, :len(gt_poses), 4:] - gt_poses[:, 4:]).max() < 0.2     # 1 px noise, 10 m scene (oracle: 0.01 .. 0.10)
    d = oracle_data(probs[0])
    st = ba_ref.optimize(d, 12, 8)
    assert st.n_culled == stats[0].n_culled
    assert np.abs(poses[0, :d.poses.shape[0]] - d.poses).max() < 1e-6
    ctx.close()
