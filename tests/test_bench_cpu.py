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
    # pyramid pixels of 752x480 with 4 levels: 360960 + 90240 + 22560 + 5640
    assert bench.P_PYR == 360960 + 90240 + 22560 + 5640
    assert bench.LK_BYTES_PER_CALL == 6 * bench.P_PYR + 29 * bench.NPTS
    for wl in bench.WORKLOADS.values():
        assert wl["w"] >= 64 and wl["h"] >= 64 and wl["window"] <= 24
