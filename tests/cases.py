"""Shared test-case builders (inputs regenerated from seeds; expected outputs in tests/golden/*.npz)."""
import glob
import os
import numpy as np

from synthdata import textures as synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def lk_cases():
    return sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "lk_*.npz")))


def gftt_cases():
    return sorted(os.path.basename(p)[5:-4] for p in glob.glob(os.path.join(GOLDEN, "gftt_*.npz")))


def load_lk(name):
    g = np.load(os.path.join(GOLDEN, f"lk_{name}.npz"))
    h, w = [int(v) for v in g["hw"]]
    I, J, _ = synth.frame_pair(int(g["seed"]), h, w, tuple(float(v) for v in g["shift"]), float(g["rot"]),
                               float(g["scale"]))
    return g, I, J


def load_gftt(name):
    g = np.load(os.path.join(GOLDEN, f"gftt_{name}.npz"))
    h, w = [int(v) for v in g["hw"]]
    img = synth.texture(int(g["seed"]), h, w, int(g["blur"]))
    return g, img


# FLVIS per-sensor feature parameters: max/region, min/region, spacing, N, q, d
FEATURE_PARA = {
    "euroc": [30, 20, 5, 1000, 0.01, 10],
    "kitti": [30, 15, 10, 2000, 0.0001, 10],
    "d435": [30, 15, 5, 500, 0.01, 15],
}
