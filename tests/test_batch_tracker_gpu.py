"""flv_f2f_batch (flvis_b200/csrc/tracker.cu + batch_tracker.cu): S camera sequences advanced together with ONE launch per
stage and device-resident landmark lists, against
  (i)  the restated reference pipeline (oracle/f2f_ref.py), each stream against its own oracle, frame by frame, with the IMU
       in the loop (teacher-forced, see tests/seq_harness.py) -- ids / flags / LK positions bit-exact, poses <= 1e-6;
  (ii) S separate flv::F2FTracking handles (the host-hand-off twin) on IMU-less sequences with the product's own device
       RANSAC: every output bit-identical.
Reference: src/frontend/f2f_tracking.cpp:59-453, src/processing/lkorb_tracking.cpp:9-202, camera_frame.cpp."""
import numpy as np
import pytest

from synthdata import sequences

from . import seq_harness as sh

pytestmark = pytest.mark.gpu


def test_batch_of_three_depth_imu_streams_matches_oracles(lib):
    seqs = [sequences.make_c0(48, seed=s, skip=6, t_rest=0.25) for s in range(3)]
    outs = sh.run(lib, seqs, batch=True, tol_pose=1e-6, tol_und=0.0, tol_p3=1e-6, lockstep=True)
    for r in outs:
        assert r["states"][:6] == ["UnInit"] * 6 and r["final_state"] == "Tracking" and r["has_imu"] == 1
        assert r["frames_tracked"] >= 40 and r["guess_used"] >= 38 and r["kf"] >= 3
        assert np.abs(r["ref_acc_bias"]).max() > 0
        assert r["path"] > 0.05 and r["ate_vs_ref"] <= 0.01 * r["path"]
    # the streams really are different sequences
    assert abs(outs[0]["path"] - outs[1]["path"]) > 0 or not np.allclose(outs[0]["traj"][-1], outs[1]["traj"][-1])


def test_batch_of_two_euroc_unrect_imu_streams_with_forced_failure(lib):
    """STEREO_UNRECT + equalizeHist + IMU on both streams; stream-independent failure handling: the unrelated frames 26 / 27
    push both streams through TrackingFail and the IMU-pose re-initialisation; a 5-keyframe local map is chained per stream."""
    seqs = [sequences.make_c1(44, seed=s, blank_frames=(26, 27), t_rest=0.3) for s in range(2)]
    outs = sh.run(lib, seqs, batch=True, tol_pose=1e-5, tol_und=6.2e-5, tol_p3=1e-5, lockstep=True, window=5)
    for r in outs:
        st = r["states"]
        assert st[27] == "TrackingFail" and st[30] == "Tracking" and r["reset"] >= 1 and r["final_state"] == "Tracking"
        assert r["kf"] >= 6 and r["n_solved"] >= 2 and r["guess_used"] >= 30


def test_batch_equals_separate_single_trackers_bit_for_bit(lib):
    """No IMU, no hooks: the batch and S separate flv_f2f handles run the same kernels (device RANSAC included) on the same
    inputs, so every per-stream output must be identical -- including a stream that fails on its own (stream 1 gets two
    unrelated frames) while the others keep tracking."""
    n = 20
    seqs = [sequences.make_c3(n, seed=0), sequences.make_c3(n, seed=1), sequences.make_c3(n, seed=2)]
    seqs[1].blank = {9, 10}
    single = sh.SingleTrackers(lib, seqs, hooks=False)
    batch = sh.BatchTracker(lib, seqs, hooks=False)
    gens = [q.frames() for q in seqs]
    saw_fail = False
    for k in range(n):
        fr = [next(g) for g in gens]
        a = single.image_feed([f[0] for f in fr], [f[1] for f in fr], [f[2] for f in fr])
        b = batch.image_feed([f[0] for f in fr], [f[1] for f in fr], [f[2] for f in fr])
        assert a == b, (k, a, b)
        for s in range(3):
            assert single.state(s) == batch.state(s), (k, s)
            saw_fail |= batch.state(s) == "TrackingFail"
            fa, fb = single.get_frame(s), batch.get_frame(s)
            assert fa[0] == fb[0], (k, s, fa[0], fb[0])
            for x, y in zip(fa[1:], fb[1:]):
                assert np.array_equal(x, y), (k, s)
            ea, eb = single.get_ex(s, fa[0]), batch.get_ex(s, fb[0])
            for x, y in zip(ea[1:], eb[1:]):
                assert np.array_equal(x, y), (k, s)
            if single.state(s) == "Tracking" and k > 0:
                assert single.counts(s) == batch.counts(s)
    assert saw_fail
    assert [batch.state(s) for s in range(3)][0] == "Tracking"
    single.close(); batch.close()


def test_grouped_batch_two_groups_matches_oracles(lib):
    """flv_f2f_batch_create_grouped: 4 depth + IMU streams in 2 groups (own context, CUDA stream and host thread each, the
    hooks are called from the group threads); per-stream results are unchanged."""
    seqs = [sequences.make_c0(30, seed=10 + s, skip=6, t_rest=0.25) for s in range(4)]
    outs = sh.run(lib, seqs, batch=True, tol_pose=1e-6, tol_und=0.0, tol_p3=1e-6, lockstep=True, groups=2)
    for r in outs:
        assert r["final_state"] == "Tracking" and r["frames_tracked"] >= 22 and r["guess_used"] >= 20
