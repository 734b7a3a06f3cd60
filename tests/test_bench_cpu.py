"""bench.py contract pieces that do not need a GPU: the reference arm prints one JSON line with the agreed keys, and the
workload table / byte accounting are self-consistent."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                          "--streams", "2"], capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["value"] > 0 and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1


def test_workload_table_and_lk_bytes():
    sys.path.insert(0, ROOT)
    import bench
    assert set(bench.WORKLOADS) == {"euroc", "kitti", "d435"}
    # pyramid pixels P(w,h) of SURVEY.md 8(d)
    assert bench.pyramid_pixels(752, 480) == 360960 + 90240 + 22560 + 5640 == 479400
    assert bench.pyramid_pixels(1241, 376) == 619930 and bench.pyramid_pixels(640, 480) == 408000
    # algorithmic bytes of one LK call = 2 P + 29 N: 972 720 B at 752x480, N = 480
    assert bench.lk_algorithmic_bytes(752, 480, 480) == 972720
    assert bench.lk_algorithmic_bytes(1241, 376, 480) == 1253780 and bench.lk_algorithmic_bytes(640, 480, 480) == 829920
    for wl in bench.WORKLOADS.values():
        assert wl["w"] >= 64 and wl["h"] >= 64 and wl["window"] <= 25


def test_periodic_bench_sequences_repeat_exactly():
    from synthdata import sequences
    for wl in ("euroc", "d435", "kitti"):
        seq = sequences.make_bench(wl, 3, 40, render=False)
        t0 = seq.n_startup / seq.img_hz; t1 = (seq.n_startup + seq.period) / seq.img_hz
        a, b = seq.T_w_c0(t0), seq.T_w_c0(t1)
        import numpy as np
        assert np.abs(a.t - b.t).max() < 1e-12 and np.abs(a.q - b.q).max() < 1e-12
