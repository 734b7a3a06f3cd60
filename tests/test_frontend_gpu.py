"""GPU parity tests (run with -m gpu on the B200): every call goes through the C ABI (ctypes) and is compared
with the CPU oracle on the same seeded inputs and with the cv2 golden vectors.  Integer / index results and
the oracle-contract floats are bit-exact; positions vs cv2 itself within 1e-3 px (stated in oracle/lk_ref.py)."""
import numpy as np
import pytest

from oracle import lk_ref, gftt_ref, feature_dem_ref
from synthdata import textures as synth
from tests import cases

pytestmark = pytest.mark.gpu


def _ctx(S, w, h, max_pts=512):
    from flvis_b200 import capi
    return capi.Context(S, w, h, max_pts)


@pytest.mark.parametrize("name", cases.lk_cases())
def test_pyramid_bit_exact(name):
    g, I, J = cases.load_lk(name)
    h, w = I.shape
    ctx = _ctx(2, w, h)
    ctx.upload(0, np.stack([I, J]))
    ctx.build_pyramid(0, 2)
    assert ctx.num_levels == 4
    for s, img in enumerate((I, J)):
        ref = lk_ref.build_pyramid(img, 31, 10)
        for l in range(4):
            assert np.array_equal(ctx.download_level(0, s, l), ref[l]), (s, l)
    assert np.array_equal(ctx.download_level(0, 0, 3), g["pyr3"])     # cv2.pyrDown chain
    ctx.close()


@pytest.mark.parametrize("name", cases.lk_cases())
def test_lk_bit_exact_vs_oracle_and_golden(name):
    g, I, J = cases.load_lk(name)
    h, w = I.shape
    ctx = _ctx(1, w, h)
    ctx.upload(0, I); ctx.upload(1, J)
    ctx.build_pyramid(0, 1); ctx.build_pyramid(1, 1)
    ml = int(g["max_level"])
    nxt, st, err = ctx.lk_track(0, 1, g["pts"], g["init"], max_level=ml)
    o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I, J, g["pts"], g["init"], max_level=ml)
    assert np.array_equal(st, o_st)
    assert np.array_equal(nxt.view(np.uint32), o_nxt.view(np.uint32))       # bit-exact vs the oracle
    assert np.array_equal(err.view(np.uint32), o_err.view(np.uint32))
    assert np.array_equal(st, g["status"])                                  # status bit-exact vs cv2
    m = st == 1
    assert np.abs(nxt - g["next"])[m].max() <= 1e-3                         # positions vs cv2: 1e-3 px
    ctx.close()


@pytest.mark.parametrize("variant", ["6", "7"])
@pytest.mark.parametrize("name", cases.lk_cases())
def test_lk_v4_packed_patch_all_cases(name, variant, monkeypatch):
    """LK v4 (precomputed Scharr pyramid + packed register patch + DP2A blend) on every golden case, incl. border points
    and a second call that reuses the cached derivative pyramid.  Variant 7 (the default) stages the second image's patch
    with TMA tile loads, variant 6 with plain loads: same bits."""
    monkeypatch.setenv("FLV_LK_VARIANT", variant)
    g, I, J = cases.load_lk(name)
    h, w = I.shape
    ctx = _ctx(1, w, h)
    ctx.upload(0, I); ctx.upload(1, J)
    ctx.build_pyramid(0, 1); ctx.build_pyramid(1, 1)
    o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I, J, g["pts"], g["init"], max_level=int(g["max_level"]))
    for rep in range(2):
        nxt, st, err = ctx.lk_track(0, 1, g["pts"], g["init"], max_level=int(g["max_level"]))
        assert np.array_equal(st, o_st)
        assert np.array_equal(nxt.view(np.uint32), o_nxt.view(np.uint32))
        assert np.array_equal(err.view(np.uint32), o_err.view(np.uint32))
    # new images in the template slot invalidate the cached derivatives
    ctx.upload(0, J); ctx.build_pyramid(0, 1)
    ctx.upload(1, I); ctx.build_pyramid(1, 1)
    nxt, st, err = ctx.lk_track(0, 1, g["pts"], g["init"], max_level=int(g["max_level"]))
    o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(J, I, g["pts"], g["init"], max_level=int(g["max_level"]))
    assert np.array_equal(st, o_st)
    assert np.array_equal(nxt.view(np.uint32), o_nxt.view(np.uint32))
    ctx.close()


@pytest.mark.parametrize("variant", ["1", "3", "4", "5", "6", "7"])
def test_lk_kernel_variants_bit_exact(variant, monkeypatch):
    """All register-budget / shared-memory variants of the tracker obey the same arithmetic contract."""
    monkeypatch.setenv("FLV_LK_VARIANT", variant)
    g, I, J = cases.load_lk("euroc_rot")
    h, w = I.shape
    ctx = _ctx(1, w, h)
    ctx.upload(0, I); ctx.upload(1, J)
    ctx.build_pyramid(0, 1); ctx.build_pyramid(1, 1)
    nxt, st, err = ctx.lk_track(0, 1, g["pts"], g["init"], max_level=int(g["max_level"]))
    o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I, J, g["pts"], g["init"], max_level=int(g["max_level"]))
    assert np.array_equal(st, o_st)
    assert np.array_equal(nxt.view(np.uint32), o_nxt.view(np.uint32))
    assert np.array_equal(err.view(np.uint32), o_err.view(np.uint32))
    ctx.close()


def test_lk_batched_streams_ragged_and_empty():
    names = cases.lk_cases()
    e = [n for n in names if n.startswith("euroc")]
    data = [cases.load_lk(n) for n in e]
    h, w = data[0][1].shape
    S = 5
    ctx = _ctx(S, w, h)
    Is = np.stack([data[i % 2][1] for i in range(S)]); Js = np.stack([data[i % 2][2] for i in range(S)])
    ctx.upload(0, Is); ctx.upload(1, Js)
    ctx.build_pyramid(0, S); ctx.build_pyramid(1, S)
    npts = np.array([480, 17, 0, 333, 1], np.int32)
    prev = np.zeros((S, 512, 2), np.float32); init = np.zeros((S, 512, 2), np.float32)
    for s in range(S):
        g = data[s % 2][0]
        prev[s, :npts[s]] = g["pts"][:npts[s]]; init[s, :npts[s]] = g["init"][:npts[s]]
    nxt, st, err = ctx.lk_track(0, 1, prev, init, npts)
    for s in range(S):
        g, I, J = data[s % 2]
        n = npts[s]
        o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I, J, g["pts"][:n], g["init"][:n])
        assert np.array_equal(st[s, :n], o_st)
        assert np.array_equal(nxt[s, :n].view(np.uint32), o_nxt.view(np.uint32))
    ctx.close()


def test_lk_points_outside_and_flat_image():
    h, w = 480, 640
    I = synth.texture(3, h, w); J = I.copy()
    I[:, :200] = 90; J[:, :200] = 90                       # flat band: minEig gate must fire
    pts = np.array([[-50.0, 10.0], [700.0, 500.0], [100.0, 240.0], [5.0, 5.0], [400.3, 200.7], [639.0, 479.0],
                    [-31.5, -31.5], [670.9, 510.9]], np.float32)
    ctx = _ctx(1, w, h)
    ctx.upload(0, I); ctx.upload(1, J)
    ctx.build_pyramid(0, 1); ctx.build_pyramid(1, 1)
    nxt, st, err = ctx.lk_track(0, 1, pts, pts)
    o_nxt, o_st, o_err = lk_ref.calc_optical_flow_pyr_lk(I, J, pts, pts)
    assert np.array_equal(st, o_st)
    assert np.array_equal(nxt.view(np.uint32), o_nxt.view(np.uint32))
    assert st[2] == 0 and st[4] == 1
    ctx.close()


@pytest.mark.parametrize("name", cases.gftt_cases())
def test_gftt_bit_exact(name):
    g, img = cases.load_gftt(name)
    h, w = img.shape
    ctx = _ctx(1, w, h)
    ctx.upload(0, img)
    ctx.keep_response(True)
    N, q, d = int(g["N"]), float(g["q"]), float(g["d"])
    c = ctx.gftt(0, 1, N, q, d)[0]
    eig = ctx.download_eig(0)
    o_eig = gftt_ref.corner_min_eigen_val(img)
    assert np.array_equal(eig.view(np.uint32), o_eig.view(np.uint32))       # response map bit-exact (oracle)
    assert np.array_equal(eig[::7, ::5], g["eig_sample"])                   # ... and vs cv2
    assert np.array_equal(c, g["corners"])                                  # corner list and order vs cv2
    ctx.close()


def test_gftt_batched_and_flat():
    names = ["euroc", "euroc_2n"]
    imgs = [cases.load_gftt(n)[1] for n in names]
    h, w = imgs[0].shape
    flat = np.full((h, w), 31, np.uint8)
    ctx = _ctx(4, w, h)
    ctx.upload(0, np.stack([imgs[0], flat, imgs[1], imgs[0]]))
    out = ctx.gftt(0, 4, 1000, 0.01, 10)
    assert np.array_equal(out[0], gftt_ref.good_features_to_track(imgs[0], 1000, 0.01, 10))
    assert len(out[1]) == 0
    assert np.array_equal(out[2], gftt_ref.good_features_to_track(imgs[1], 1000, 0.01, 10))
    assert np.array_equal(out[3], out[0])
    ctx.close()


@pytest.mark.parametrize("sensor,name", [("euroc", "euroc"), ("kitti", "kitti"), ("d435", "d435"), ("euroc", "euroc_2n")])
def test_feature_dem_detect_redetect_bit_exact(sensor, name):
    from flvis_b200 import capi
    g, img = cases.load_gftt(name)
    h, w = img.shape
    para = cases.FEATURE_PARA[sensor]
    fd = feature_dem_ref.FeatureDEM(w, h, para)
    fp = capi.FeatureParams(fd.max_region_feature_num, fd.min_region_feature_num, fd.boundary_dis, fd.gftt_num,
                            fd.gftt_ql, fd.gftt_dis)
    ctx = _ctx(2, w, h)
    ctx.upload(0, np.stack([img, img[::-1].copy()]))
    det = ctx.feature_detect(0, 2, fp)
    o0 = fd.detect(img); o1 = fd.detect(img[::-1].copy())
    assert np.array_equal(det[0], o0) and np.array_equal(det[1], o1)
    # redetect with a thinned, sub-pixel-shifted subset as the existing features
    ex0 = o0[::3].astype(np.float64) + 0.37
    ex1 = o1[1::2].astype(np.float64) - 0.21
    red = ctx.feature_redetect(0, 2, fp, [ex0, ex1])
    assert np.array_equal(red[0], fd.redetect(img, ex0))
    assert np.array_equal(red[1], fd.redetect(img[::-1].copy(), ex1))
    # no existing features at all
    red = ctx.feature_redetect(0, 2, fp, [np.zeros((0, 2)), ex1])
    assert np.array_equal(red[0], fd.redetect(img, np.zeros((0, 2))))
    # flv_feature_prepare (GFTT started early on the auxiliary stream) must not change any result; a prepared GFTT
    # that does not match the next call (detect wants 2N corners) is discarded
    ctx.feature_prepare(0, 2, fp, redetect=True)
    red = ctx.feature_redetect(0, 2, fp, [ex0, ex1])
    assert np.array_equal(red[0], fd.redetect(img, ex0)) and np.array_equal(red[1], fd.redetect(img[::-1].copy(), ex1))
    ctx.feature_prepare(0, 2, fp, redetect=True)
    det = ctx.feature_detect(0, 2, fp)
    assert np.array_equal(det[0], o0) and np.array_equal(det[1], o1)
    ctx.feature_prepare(0, 2, fp, redetect=False)
    det = ctx.feature_detect(0, 2, fp)
    assert np.array_equal(det[0], o0) and np.array_equal(det[1], o1)
    ctx.close()


def test_equalize_hist_on_ingest_bit_exact():
    """flv_set_equalize_hist: level 0 (and the pyramid built from it) of an ingested image == cv2.equalizeHist(image);
    host and device sources, a low-contrast image, and a constant image (OpenCV's early-out)."""
    import cv2
    import torch
    g, I, J = cases.load_lk("euroc_shift")
    h, w = I.shape
    low = (I // 4 + 60).astype(np.uint8)
    const = np.full_like(I, 77)
    imgs = np.stack([I, low, const, J])
    ctx = _ctx(4, w, h)
    ctx.set_equalize_hist(True)
    ctx.upload(0, imgs)
    ctx.build_pyramid(0, 4)
    for s in range(4):
        ref = cv2.equalizeHist(imgs[s])
        assert np.array_equal(ctx.download_level(0, s, 0), ref)
        assert np.array_equal(ctx.download_level(0, s, 1), lk_ref.pyr_down(ref))
    d = torch.from_numpy(imgs).cuda()
    ctx.upload_dev(1, 4, d.data_ptr())
    torch.cuda.synchronize()
    for s in range(4):
        assert np.array_equal(ctx.download_level(1, s, 0), cv2.equalizeHist(imgs[s]))
    ctx.set_equalize_hist(False)
    ctx.upload(2, imgs)
    assert np.array_equal(ctx.download_level(2, 1, 0), low)
    ctx.close()


@pytest.mark.parametrize("ch,rgb", [(3, False), (3, True), (4, False), (4, True)])
def test_color_ingest_matches_cv2_cvtcolor(ch, rgb):
    """flv_upload_color_images == cv2.cvtColor(BGR/RGB/BGRA/RGBA -> GRAY) (f2f_tracking.cpp:78-110), bit-exact."""
    import cv2
    rng = np.random.default_rng(ch * 2 + rgb)
    h, w = 480, 752
    imgs = rng.integers(0, 256, (2, h, w, ch), dtype=np.uint8)
    code = {(3, False): cv2.COLOR_BGR2GRAY, (3, True): cv2.COLOR_RGB2GRAY, (4, False): cv2.COLOR_BGRA2GRAY,
            (4, True): cv2.COLOR_RGBA2GRAY}[(ch, rgb)]
    ctx = _ctx(2, w, h)
    ctx.upload_color(0, imgs, is_rgb=rgb)
    for s in range(2):
        assert np.array_equal(ctx.download_level(0, s, 0), cv2.cvtColor(imgs[s], code))
    ctx.close()
