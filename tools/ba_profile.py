"""Per-phase SM cycle counters of the BA kernel on an EuRoC-sized window (debug aid; run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from flvis_b200 import capi, ba_batch
from synthdata import ba_problems

W = int(sys.argv[1]) if len(sys.argv) > 1 else 10
S = int(sys.argv[2]) if len(sys.argv) > 2 else 1
OBS = int(sys.argv[3]) if len(sys.argv) > 3 else 480
probs = [ba_problems.make_problem(window=W, n_landmarks=(3 * OBS if W <= 10 else 2000), obs_per_frame=OBS, seed=2 + s) for s in range(S)]
batch = ba_batch.Batch(probs)
ctx = capi.Context(S, 752, 480)
for rep in range(2):
    t = time.perf_counter()
    poses, lms, active, stats = ba_batch.solve_batch_host(ctx, batch)
    dt = time.perf_counter() - t
names = ["chi2", "build:pose", "schur:prod", "cholesky", "subst", "update", "setup", "-", "schur:init", "schur:stage", "build:edge", "build:lm", "-", "-", "-", "-"]
pr = ctx.ba_profile(0)
tot = pr.sum()
print(f"W={W} S={S} E={len(probs[0].ep)} L={len(probs[0].lms)} iters={stats[0].iterations_run} culled={stats[0].n_culled} "
      f"host wall {dt*1e3:.2f} ms, kernel cycles {tot} (~{tot/1.9e6:.2f} ms @1.9GHz)")
for n, v in zip(names, pr):
    print(f"  {n:12s} {v:12d} cycles {100.0*v/max(tot,1):5.1f}%")
