"""CPU tests: the oracle restatements against the golden vectors produced by the reference's own library
(cv2 4.13.0, tests/golden/make_golden.py), and the C-ABI library's symbol table (no compute calls)."""
import numpy as np
import pytest

from oracle import lk_ref, gftt_ref, feature_dem_ref
from tests import cases


@pytest.mark.parametrize("name", cases.lk_cases())
def test_lk_oracle_matches_cv2_golden(name):
    g, I, J = cases.load_lk(name)
    # pyramid: bit-exact vs cv2.pyrDown
    pyr = lk_ref.build_pyramid(I, 31, 10)
    assert len(pyr) == 4
    assert [int(p.astype(np.uint64).sum()) for p in pyr] == [int(v) for v in g["pyr_crc"]]
    assert np.array_equal(pyr[3], g["pyr3"])
    nxt, st, err = lk_ref.calc_optical_flow_pyr_lk(I, J, g["pts"], g["init"], max_level=int(g["max_level"]))
    assert np.array_equal(st, g["status"])          # status: bit-exact
    m = st == 1
    d = np.abs(nxt - g["next"])[m]
    assert d.max() <= 1e-3                          # positions: stated tolerance 1e-3 px (cv2 sums in SIMD lanes)
    assert np.abs(err - g["err"])[m].max() <= 2e-3       # err follows the (<=1e-3 px different) end point; FLVIS ignores it


@pytest.mark.parametrize("name", cases.gftt_cases())
def test_gftt_oracle_bit_exact_vs_cv2_golden(name):
    g, img = cases.load_gftt(name)
    eig = gftt_ref.corner_min_eigen_val(img)
    assert np.array_equal(eig[::7, ::5], g["eig_sample"])           # response map: bit-exact
    assert eig.max() == g["eig_max"]
    assert np.array_equal(eig.astype(np.float64).sum(axis=1), g["eig_rowsum"])
    c = gftt_ref.good_features_to_track(img, int(g["N"]), float(g["q"]), float(g["d"]), eig=eig)
    assert np.array_equal(c, g["corners"])                          # corner list + order: bit-exact


def test_gftt_small_distance_and_empty():
    img = np.full((80, 96), 7, np.uint8)        # flat image: no corners
    assert len(gftt_ref.good_features_to_track(img, 10, 0.01, 5)) == 0


def test_std_sort_helper_is_a_permutation_and_sorted():
    rng = np.random.default_rng(0)
    s = rng.integers(0, 5, 200).astype(np.float32)      # many ties, n > 16 => introsort path
    p = feature_dem_ref.std_sort_desc(s)
    assert sorted(p.tolist()) == list(range(200))
    assert np.all(np.diff(s[p]) <= 0)


def test_feature_dem_oracle_properties():
    g, img = cases.load_gftt("euroc")
    h, w = img.shape
    fd = feature_dem_ref.FeatureDEM(w, h, cases.FEATURE_PARA["euroc"])
    pts = fd.detect(img)
    assert 0 < len(pts) <= 16 * 30
    # the `||` spacing rule: inside a region no two kept points share a row/column band of +-boundary_dis
    reg = [fd._region_of(p) for p in pts]
    assert reg == sorted(reg)
    for r in range(16):
        q = pts[np.array(reg) == r]
        assert len(q) <= 30
        for i in range(len(q)):
            for j in range(i):
                assert abs(q[i, 0] - q[j, 0]) > fd.boundary_dis and abs(q[i, 1] - q[j, 1]) > fd.boundary_dis
    new = fd.redetect(img, pts[::2].astype(np.float64) + 0.25)
    assert len(new) > 0


def test_library_exports_every_declared_symbol(lib):
    from flvis_b200 import capi
    import re, os
    hdr = open(os.path.join(os.path.dirname(capi._HERE), "include", "flvis_b200.h")).read()
    declared = set(re.findall(r"\b(flv_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(capi.SYMBOLS), declared ^ set(capi.SYMBOLS)
    for s in declared:
        assert hasattr(lib, s), s
    assert b"sm_100a" in lib.flv_version()
