"""Parity of the CUDA path (libsdorb.so, called through its C ABI via ctypes) with the CPU oracle and the committed
cv2-generated fixtures.  Bit-exact everywhere: keypoint x / y / size / angle / response / octave / class_id and their
ORDER, descriptors, pyramid bytes (the bar of BASELINE.md section 6; the 1e-3 degree angle tolerance of the north
star is met with 0).  Needs a B200: run with  pytest -m gpu.
"""
import ctypes as C
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")) if "hamming" not in p)
C1 = (1000, 1.2, 8, 20)


def kp_view(a):
    return np.ascontiguousarray(a).view(api.KP_DTYPE)


def assert_same(ok, od, gk, gd, what=""):
    assert len(ok) == len(gk), "%s: %d oracle vs %d gpu keypoints" % (what, len(ok), len(gk))
    for f in api.KP_DTYPE.names:
        bad = np.flatnonzero(ok[f] != gk[f])
        assert len(bad) == 0, "%s: kp.%s differs at %s (first: oracle %r gpu %r)" % (what, f, bad[:5], ok[f][bad[0]], gk[f][bad[0]])
    assert ok.tobytes() == gk.tobytes()
    assert np.array_equal(od, gd), "%s: %d descriptor rows differ" % (what, int((od != gd).any(axis=1).sum()))


@pytest.fixture(scope="module")
def ex_c1():
    e = api.ORBextractor(*C1, max_width=752, max_height=480, max_batch=8)
    yield e
    e.close()


# ------------------------------------------------------------------ fixtures made with real cv2
@pytest.mark.parametrize("name", CASES)
def test_golden_fixture(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = g["params"]
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]))
    img = g["image"]
    ex = api.ORBextractor(*params, max_width=img.shape[1], max_height=img.shape[0], max_batch=2)
    k, d, pyr = ex(img)
    assert_same(g["kps"].astype(api.KP_DTYPE), g["desc"].reshape(-1, 32), k, d, name)
    assert len(pyr) == params[2]
    for lvl, ((h, w), sha) in enumerate(zip(g["pyramid_shape"], g["pyramid_sha256"])):
        assert pyr[lvl].shape == (int(h), int(w))
        assert hashlib.sha256(np.ascontiguousarray(pyr[lvl]).tobytes()).hexdigest() == str(sha), "pyramid level %d" % lvl
    ex.close()


# ------------------------------------------------------------------ stage by stage against the oracle
@pytest.mark.parametrize("label,img,params", [
    ("c1_smooth", synth.smooth_noise(40), C1),
    ("c1_rects", synth.rects(41), C1),
    ("c0_default", synth.smooth_noise(42), (1000, 2.0, 5, 20)),
    ("c2", synth.smooth_noise(43, 752, 480), C1),
    ("ini_2000", synth.smooth_noise(44), (2000, 1.2, 8, 20)),
    ("odd_size", synth.smooth_noise(45, 333, 257), (700, 1.2, 6, 12)),
    ("portrait", synth.smooth_noise(46, 240, 400), (400, 1.3, 5, 20)),
    ("c5_small", synth.smooth_noise(47, 960, 540), (4000, 1.2, 12, 20)),
    ("noise_th7", np.random.default_rng(48).integers(0, 256, (300, 400), dtype=np.uint8), (1500, 1.2, 8, 7)),
    ("th_high", np.random.default_rng(49).integers(0, 256, (240, 320), dtype=np.uint8), (500, 1.2, 4, 140)),
])
def test_stages_match_oracle(label, img, params):
    h, w = img.shape
    o = orc.Extractor(*params)
    ok, od, st = o.extract(img, dump=True)
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=3)
    gk, gd, pyr = ex(img)
    off = cell_off = 0
    for l, g in enumerate(st["geometry"]):
        lw, lh = int(g["width"]), int(g["height"])
        n = lw * lh
        assert np.array_equal(st["pyramid"][off:off + n].reshape(lh, lw), pyr[l]), "%s pyramid level %d" % (label, l)
        nc = max(int(g["level_cols"]), 0) * max(int(g["level_rows"]), 0)
        okl = ok[ok["octave"] == l]
        if len(okl):  # the reference only blurs levels that have keypoints; the device blurs all of them
            bl = ex.debug_read(api.DBG_BLURRED_LEVEL, 0, l, n).reshape(lh, lw)
            assert np.array_equal(st["blurred"][off:off + n].reshape(lh, lw), bl), "%s blurred level %d" % (label, l)
        if nc:
            cc = ex.debug_read(api.DBG_CELL_COUNTS, 0, l, nc * 4, np.int32)
            assert np.array_equal(cc, st["raw_cell_count"][cell_off:cell_off + nc]), "%s FAST counts level %d" % (label, l)
        sel = ex.debug_read(api.DBG_LEVEL_SELECTED, 0, l, 4 * max(int(g["n_desired"]), 1), np.uint32)
        assert len(sel) == len(okl) == int(st["level_count"][l])
        off += n
        cell_off += nc
    assert_same(ok, od, gk, gd, label)
    ex.close()


def test_c5_full_size_1080p_matches_oracle():
    """BASELINE config C5 at its real size: 1920x1080, 4000 keypoints, 12 levels (144-cell grids, lists beyond the
    shared-memory capacity of the selection kernel, 4-, 8- and 16-word narrow FAST tiles)."""
    params = (4000, 1.2, 12, 20)
    img = synth.smooth_noise(300, 1920, 1080)
    ok, od = orc.Extractor(*params).extract(img)
    ex = api.ORBextractor(*params, max_width=1920, max_height=1080, max_batch=2)
    k, d, pyr = ex(img)
    assert len(ok) == 4000
    assert_same(ok, od, k, d, "C5 1080p")
    kk, dd, cc = ex.extract_batch_host(np.stack([img, synth.rects(301, 1920, 1080)]))
    assert_same(ok, od, kk[0, :cc[0]], dd[0, :cc[0]], "C5 batch frame 0")
    ok1, od1 = orc.Extractor(*params).extract(synth.rects(301, 1920, 1080))
    assert_same(ok1, od1, kk[1, :cc[1]], dd[1, :cc[1]], "C5 rects")
    ex.close()


@pytest.mark.parametrize("w,h", [(39 + 64, 39 + 40), (128 + 19, 101), (641, 479), (1241, 376)])
def test_odd_sizes_and_tile_edges(w, h):
    """Widths around the 120-column FAST tiles / 128-column blur tiles and heights around the 30-row tiles."""
    params = (600, 1.2, 5, 15)
    img = np.random.default_rng(w * 7 + h).integers(0, 256, (h, w), dtype=np.uint8)
    img = (img // 3 + synth.smooth_noise(w, w, h) // 2).astype(np.uint8)
    try:
        ok, od = orc.Extractor(*params).extract(img)
    except RuntimeError:
        ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=1)
        with pytest.raises(api.SdorbError):
            ex(img)
        ex.close()
        return
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=1)
    k, d, _ = ex(img)
    assert_same(ok, od, k, d, "%dx%d" % (w, h))
    ex.close()


# ------------------------------------------------------------------ entry-point equivalence
def test_batch_host_equals_single_and_oracle(ex_c1):
    imgs = synth.frames(5, 640, 480, start=60)
    imgs[3] = synth.rects(3)
    kps, desc, cnt = ex_c1.extract_batch_host(imgs)
    o = orc.Extractor(*C1)
    for f in range(len(imgs)):
        ok, od = o.extract(imgs[f])
        assert_same(ok, od, kps[f, :cnt[f]], desc[f, :cnt[f]], "batch frame %d" % f)
        sk, sd, _ = ex_c1(imgs[f], want_pyramid=False)
        assert sk.tobytes() == ok.tobytes() and np.array_equal(sd, od)


def test_batch_larger_than_max_batch_and_ragged_tail(ex_c1):
    """19 frames through passes of max_batch=8 (two full passes + a tail of 3), three-stream pipeline."""
    base = synth.frames(4, 640, 480, start=70)
    imgs = np.concatenate([base] * 5)[:19]
    kps, desc, cnt = ex_c1.extract_batch_host(imgs)
    for f in range(19):
        assert cnt[f] == cnt[f % 4] and kps[f].tobytes() == kps[f % 4].tobytes() and desc[f].tobytes() == desc[f % 4].tobytes()
    o = orc.Extractor(*C1)
    ok, od = o.extract(base[2])
    assert_same(ok, od, kps[18, :cnt[18]], desc[18, :cnt[18]])


@pytest.mark.parametrize("const", [0, 3])
def test_two_compute_lanes_equal_one(ex_c1, const, monkeypatch):
    """Host pipeline with two compute lanes (SDORB_PIPE_DUAL: odd passes on a twin scratch arena and stream): 23 frames through
    passes of at most 4 frames, geometric and constant pass schedules, byte-identical to the one-lane result, the oracle and the
    stage accounting of one lane (launch counts, stage events of both lanes)."""
    base = synth.frames(5, 640, 480, start=90)
    base[3] = synth.rects(4)
    imgs = np.concatenate([base] * 5)[:23]
    monkeypatch.setenv("SDORB_PIPE_DUAL", "0")  # the schedule switches are read when the handle is created
    monkeypatch.setenv("SDORB_PIPE_CONST", "0")
    ex1 = api.ORBextractor(*C1, max_width=640, max_height=480, max_batch=4)
    try:
        kps, desc, cnt = ex1.extract_batch_host(imgs)
    finally:
        ex1.close()
    monkeypatch.setenv("SDORB_PIPE_DUAL", "1")
    monkeypatch.setenv("SDORB_PIPE_CONST", str(const))
    ex2 = api.ORBextractor(*C1, max_width=640, max_height=480, max_batch=4)
    try:
        for rep in range(2):  # the second call reuses the twin
            k2, d2, c2 = ex2.extract_batch_host(imgs)
            assert np.array_equal(c2, cnt), "call %d" % rep
            for f in range(len(imgs)):
                assert k2[f, :cnt[f]].tobytes() == kps[f, :cnt[f]].tobytes() and d2[f, :cnt[f]].tobytes() == desc[f, :cnt[f]].tobytes(), \
                    "call %d frame %d" % (rep, f)
        o = orc.Extractor(*C1)
        for f in (3, 22):
            ok, od = o.extract(imgs[f])
            assert_same(ok, od, k2[f, :c2[f]], d2[f, :c2[f]], "frame %d" % f)
        before = ex2.kernel_launches()
        ex2.set_profiling(True)
        ex2.stage_times(reset=True)
        ex2.extract_batch_host(imgs)
        ms, launches = ex2.stage_times()
        ex2.set_profiling(False)
        npass = launches["fast"]
        assert npass >= 6 and launches["pyramid"] == 7 * npass and ex2.kernel_launches() - before == 12 * npass
        assert all(ms[s] > 0 for s in ("pyramid", "fast", "select", "blur", "describe"))
        ex2(imgs[0], want_pyramid=False)  # the single-frame entry on the same handle afterwards
    finally:
        ex2.close()


def test_device_entry_point_equals_host_entry_point(ex_c1):
    torch = pytest.importorskip("torch")
    imgs = synth.frames(6, 640, 480, start=80)
    hk, hd, hc = ex_c1.extract_batch_host(imgs)
    dev = torch.device("cuda:0")
    cap = ex_c1.max_keypoints
    t_img = torch.from_numpy(imgs).to(dev)
    k = torch.zeros((6, cap, 7), dtype=torch.float32, device=dev)
    d = torch.zeros((6, cap, 32), dtype=torch.uint8, device=dev)
    c = torch.zeros(6, dtype=torch.int32, device=dev)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ex_c1.extract_batch_device(t_img, k, d, c, stream=s.cuda_stream)
    s.synchronize()
    ex_c1.batch_status()
    assert np.array_equal(c.cpu().numpy(), hc)
    for f in range(6):
        n = int(hc[f])
        assert kp_view(k[f, :n].cpu().numpy()).reshape(-1).tobytes() == hk[f, :n].tobytes()
        assert np.array_equal(d[f, :n].cpu().numpy(), hd[f, :n])


def test_strided_and_unaligned_inputs(ex_c1):
    """cv::Mat inputs may be ROIs: row stride > width, base pointer not 16-byte aligned."""
    big = np.zeros((500, 700), np.uint8)
    img = synth.smooth_noise(90)
    big[7:487, 13:653] = img
    view = big[7:487, 13:653]
    assert view.strides[0] == 700 and view.ctypes.data % 4 != 0
    k, d, _ = ex_c1(view, want_pyramid=False)
    ok, od = orc.Extractor(*C1).extract(img)
    assert_same(ok, od, k, d, "strided view")
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    t = torch.zeros((2, 480, 651), dtype=torch.uint8, device=dev)  # odd pitch: the library repacks level 0
    t[:, :, :640] = torch.from_numpy(np.stack([img, img])).to(dev)
    cap = ex_c1.max_keypoints
    kk = torch.zeros((2, cap, 7), dtype=torch.float32, device=dev)
    dd = torch.zeros((2, cap, 32), dtype=torch.uint8, device=dev)
    cc = torch.zeros(2, dtype=torch.int32, device=dev)
    ex_c1.extract_batch_device(t[:, :, :640], kk, dd, cc)
    torch.cuda.synchronize()
    ex_c1.batch_status()
    for f in range(2):
        assert int(cc[f]) == len(ok)
        assert kp_view(kk[f, :len(ok)].cpu().numpy()).reshape(-1).tobytes() == ok.tobytes()
        assert np.array_equal(dd[f, :len(ok)].cpu().numpy(), od)


def test_geometry_switch_on_one_handle(ex_c1):
    """One handle serves different image sizes (Tracking feeds whatever the camera delivers)."""
    o = orc.Extractor(*C1)
    for i, (w, h) in enumerate([(640, 480), (752, 480), (320, 240), (640, 480)]):
        img = synth.smooth_noise(100 + i, w, h)
        k, d, pyr = ex_c1(img)
        ok, od = o.extract(img)
        assert_same(ok, od, k, d, "%dx%d" % (w, h))


# ------------------------------------------------------------------ edge cases of the reference
def test_empty_image_leaves_outputs_untouched(ex_c1):
    assert ex_c1(np.zeros((0, 0), np.uint8)) == (None, None, None)  # src/ORBextractor.cc:622-623
    n = C.c_int(1234)
    rc = api.lib().sdorb_extract(ex_c1._h, None, 0, 0, 0, None, None, 0, C.byref(n), None)
    assert rc == 0 and n.value == 1234


def test_zero_corner_images(ex_c1):
    for img in (np.full((480, 640), 0, np.uint8), np.full((480, 640), 255, np.uint8), np.tile(np.arange(640, dtype=np.uint8), (480, 1))):
        k, d, pyr = ex_c1(img)
        ok, od = orc.Extractor(*C1).extract(img)
        assert len(k) == len(ok) and (len(ok) == 0 or k.tobytes() == ok.tobytes())
        assert d.shape == (len(ok), 32)


def test_flat_and_corner_dense_regions_in_one_frame(ex_c1):
    """The FAST kernel skips pair slots whose compass test fails, stops testing where every slot is needed, and writes the map
    of a tile without any score straight away: a frame with flat halves / quadrants / stripes next to noise exercises all three
    paths and their borders (tiles are 120 x 60 outputs, half rows of 64 pixels are the skip unit)."""
    rng = np.random.default_rng(77)
    noise = rng.integers(0, 256, (480, 752), dtype=np.uint8)
    cases = []
    a = np.full((480, 752), 128, np.uint8)
    a[:, 376:] = noise[:, 376:]                      # right half noise
    cases.append(a)
    b = np.full((480, 752), 30, np.uint8)
    b[240:, :] = noise[240:, :]                      # lower half noise
    cases.append(b)
    c = noise.copy()
    c[100:380, 150:600] = 200                        # a flat window inside noise
    cases.append(c)
    d = np.full((480, 752), 90, np.uint8)
    d[::97, :] = 255                                 # thin lines on a flat background
    d[:, ::131] = 0
    d[200:203, 300:303] = 255                        # and one isolated blob
    cases.append(d)
    e = np.full((480, 640), 128, np.uint8)
    e[200:280, 250:390] = noise[200:280, 250:390]    # noise only in the middle tiles
    cases.append(e)
    for i, img in enumerate(cases):
        k, dsc, _ = ex_c1(img)
        ok, od = orc.Extractor(*C1).extract(img)
        assert len(ok) > 0
        assert_same(ok, od, k, dsc, "case %d" % i)


def test_fewer_corners_than_quota():
    img = np.full((240, 320), 90, np.uint8)
    img[60:100, 80:140] = 200
    img[150:190, 200:260] = 10
    params = (500, 1.2, 6, 20)
    ex = api.ORBextractor(*params, max_width=320, max_height=240, max_batch=1)
    k, d, _ = ex(img)
    ok, od = orc.Extractor(*params).extract(img)
    assert 0 < len(ok) < 500
    assert_same(ok, od, k, d)
    ex.close()


def test_massive_response_ties():
    """Every corner of a regular dot grid has the same response, so all trims cut inside a tie."""
    img = np.full((300, 400), 60, np.uint8)
    img[10::6, 10::6] = 250
    params = (300, 1.2, 4, 20)
    ex = api.ORBextractor(*params, max_width=400, max_height=300, max_batch=1)
    k, d, _ = ex(img)
    ok, od = orc.Extractor(*params).extract(img)
    assert len(ok) > 100 and len(np.unique(ok[ok["octave"] == 0]["response"])) <= 2
    assert_same(ok, od, k, d)
    ex.close()


@pytest.mark.parametrize("params", [(0, 1.2, 3, 20), (1, 1.2, 8, 20), (7, 1.2, 2, 20), (50, 1.2, 1, 20), (300, 1.05, 12, 5),
                                    (1000, 1.9, 4, 20), (20000, 1.2, 8, 7)])
def test_extreme_parameters(params):
    """Degenerate but legal Config values (src/Config.cc:108-111): no features, one feature, one level, a very fine and
    a very coarse pyramid, and far more features requested than the image has corners."""
    img = synth.smooth_noise(700 + params[0] % 97, 400, 300)
    try:
        ok, od = orc.Extractor(*params).extract(img)
    except RuntimeError:
        ex = api.ORBextractor(*params, max_width=400, max_height=300, max_batch=1)
        with pytest.raises(api.SdorbError):
            ex(img)
        ex.close()
        return
    ex = api.ORBextractor(*params, max_width=400, max_height=300, max_batch=2)
    k, d, pyr = ex(img)
    assert_same(ok, od, k, d, str(params))
    assert len(pyr) == params[2]
    ex.close()


def test_too_small_image_is_geometry_error():
    ex = api.ORBextractor(*C1, max_width=64, max_height=64, max_batch=1)
    with pytest.raises(api.SdorbError) as e:
        ex(np.zeros((30, 40), np.uint8))
    assert e.value.code == -4  # the reference throws cv::Exception; the oracle reports the same
    with pytest.raises(RuntimeError):
        orc.Extractor(*C1).extract(np.zeros((30, 40), np.uint8))
    ex.close()


def test_error_codes(ex_c1):
    img = synth.smooth_noise(0)
    L = api.lib()
    k = np.zeros(10, api.KP_DTYPE)
    d = np.zeros((10, 32), np.uint8)
    n = C.c_int(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    assert L.sdorb_extract(ex_c1._h, p(img), 640, 480, 640, p(k), p(d), 10, C.byref(n), None) == -2  # capacity
    big = np.zeros((481, 800), np.uint8)
    assert L.sdorb_extract(ex_c1._h, p(big), 800, 481, 800, p(k), p(d), 10, C.byref(n), None) in (-1, -2)
    kk = np.zeros(1000, api.KP_DTYPE)
    dd = np.zeros((1000, 32), np.uint8)
    assert L.sdorb_extract(ex_c1._h, p(big), 800, 481, 800, p(kk), p(dd), 1000, C.byref(n), None) == -1  # > max_width
    assert L.sdorb_extract(ex_c1._h, p(img), 640, 480, 100, p(kk), p(dd), 1000, C.byref(n), None) == -1  # stride < width


def test_getters_match_oracle(ex_c1):
    t = orc.Extractor(*C1).tables()
    assert ex_c1.GetLevels() == 8 and ex_c1.GetScaleFactor() == float(np.float32(1.2))
    assert ex_c1.GetScaleFactors().tobytes() == t["scale"].tobytes()
    assert ex_c1.GetInverseScaleFactors().tobytes() == t["inv_scale"].tobytes()
    assert ex_c1.GetScaleSigmaSquares().tobytes() == t["sigma2"].tobytes()
    assert ex_c1.GetInverseScaleSigmaSquares().tobytes() == t["inv_sigma2"].tobytes()
    assert ex_c1.features_per_level().tolist() == t["n_per_level"].tolist()


def test_pyramid_output_has_reflect101_border(ex_c1):
    """imagePyramid[l] is a view into a buffer padded by 19 px of BORDER_REFLECT_101 (src/ORBextractor.cc:684-696)."""
    img = synth.smooth_noise(110)
    _, _, pyr = ex_c1(img)
    for l, lvl in enumerate(pyr):
        padded = lvl.base if lvl.base is not None else None
        assert padded is not None and padded.shape == (lvl.shape[0] + 38, lvl.shape[1] + 38)
        assert np.array_equal(padded, orc.border_reflect101(np.ascontiguousarray(lvl), 19)), "level %d" % l
    assert np.array_equal(pyr[0], img)


def test_pyramid_then_match_then_pyramid_on_one_handle():
    """The reference's sequence on ONE handle (ORBdistance shares the extractor's): operator() with imagePyramid, a host-memory
    match that has to grow the matcher scratch, operator() with imagePyramid again.  The pinned pyramid staging buffer must
    survive the matcher's reallocation (round-1 advisor finding: it was freed there and then used and freed again)."""
    ex = api.ORBextractor(*C1, max_width=640, max_height=480, max_batch=2)
    o = orc.Extractor(*C1)
    imgs = [synth.smooth_noise(120), synth.rects(121)]
    k0, d0, pyr0 = ex(imgs[0])
    k1, d1, _ = ex(imgs[1])
    for grow in (1, 3, 9):  # each call needs a larger scratch buffer than the one before
        A = np.repeat(d0[None], grow, axis=0)
        Bm = np.repeat(d1[None], grow, axis=0)
        nA = np.full(grow, len(d0), np.int32)
        nB = np.full(grow, len(d1), np.int32)
        m = ex.match_batch(A, nA, Bm, nB)
        assert m[0, :len(d0)].tobytes() == orc.match_best2(d0, d1).tobytes()
        assert ex.hamming_matrix(d0[:50 * grow], d1[:40 * grow]).tobytes() == orc.hamming_matrix(d0[:50 * grow], d1[:40 * grow]).tobytes()
        for img in imgs:
            k, d, pyr = ex(img)
            ok, od, st = o.extract(img, dump=True)
            assert_same(ok, od, k, d, "after match (grow %d)" % grow)
            off = 0
            for l, g in enumerate(st["geometry"]):
                w, h = int(g["width"]), int(g["height"])
                assert np.array_equal(pyr[l], st["pyramid"][off:off + w * h].reshape(h, w)), "pyramid level %d after match" % l
                assert np.array_equal(pyr[l].base, orc.border_reflect101(np.ascontiguousarray(pyr[l]), 19))
                off += w * h
    ex.close()


@pytest.mark.parametrize("first_level", [0, 1])
def test_batch_pyramid_output_equals_oracle(first_level):
    """sdorb_extract_batch_pyr: the reference's fifth output (imagePyramid, src/ORBextractor.cc:620-621) for a batch, host and
    device memory, more frames than one pass, levels compared byte for byte with the oracle's pyramid."""
    import torch
    w, h, nf = 333, 257, 11
    params = (700, 1.2, 6, 12)
    imgs = synth.frames(nf, w, h, start=300)
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=4)
    off, fb = ex.pyramid_layout(w, h)
    assert fb % 16 == 0 and all(o % 16 == 0 for o in off)
    slab = np.full(nf * fb, 0xAB, np.uint8)
    k, d, c = ex.extract_batch_host(imgs, pyramid=slab, first_level=first_level)
    o = orc.Extractor(*params)
    dev = torch.device("cuda", 0)
    dimgs = torch.from_numpy(imgs).to(dev)
    cap = ex.max_keypoints
    dk = torch.zeros((nf, cap, 7), dtype=torch.float32, device=dev)
    dd = torch.zeros((nf, cap, 32), dtype=torch.uint8, device=dev)
    dc = torch.zeros(nf, dtype=torch.int32, device=dev)
    dslab = torch.full((nf * fb,), 0xAB, dtype=torch.uint8, device=dev)
    ex.extract_batch_device(dimgs, dk, dd, dc, pyramid=dslab, first_level=first_level)
    torch.cuda.synchronize()
    ex.batch_status()
    hslab = dslab.cpu().numpy()
    for f in range(nf):
        ok, od, st = o.extract(imgs[f], dump=True)
        assert c[f] == len(ok) and k[f, :c[f]].tobytes() == ok.tobytes() and np.array_equal(d[f, :c[f]], od)
        po = 0
        for which, sl in (("host", slab), ("device", hslab)):
            lv = ex.pyramid_levels(sl, f, w, h)
            po = 0
            for l, g in enumerate(st["geometry"]):
                lw, lh = int(g["width"]), int(g["height"])
                exp = st["pyramid"][po:po + lw * lh].reshape(lh, lw)
                po += lw * lh
                if l == 0 and first_level == 1:
                    assert (lv[0] == 0xAB).all(), "level 0 must stay untouched with first_level = 1"
                else:
                    assert np.array_equal(lv[l], exp), "%s slab: frame %d level %d" % (which, f, l)
    ex.close()


@pytest.mark.parametrize("border,extra_stride", [(0, 0), (0, 13), (5, 0), (5, 7), (19, 3), (25, 0)])
def test_pyramid_views_of_any_border_and_stride(border, extra_stride):
    """sdorb_pyr_view with layouts other than the reference's own (19 px border, rows w + 38 apart): tight levels, smaller and
    larger borders, padded strides -- the level bytes and the BORDER_REFLECT_101 margin must equal the oracle's."""
    w, h = 200, 160
    params = (300, 1.2, 4, 20)
    img = synth.smooth_noise(810, w, h)
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=1)
    cap = ex.max_keypoints
    kps, desc, n = np.zeros(cap, api.KP_DTYPE), np.zeros((cap, 32), np.uint8), C.c_int(0)
    views = (api._PyrView * 4)()
    bufs = []
    for l in range(4):
        lw, lh = ex.level_size(w, h, l)
        buf = np.full((lh + 2 * border, lw + 2 * border + extra_stride), 0xCD, np.uint8)
        bufs.append(buf)
        inner = buf[border:border + lh, border:border + lw]
        views[l] = api._PyrView(inner.ctypes.data, lw, lh, buf.strides[0], border)
    rc = api.lib().sdorb_extract(ex._h, img.ctypes.data_as(C.c_void_p), w, h, img.strides[0], kps.ctypes.data_as(C.c_void_p),
                                 desc.ctypes.data_as(C.c_void_p), cap, C.byref(n), C.cast(views, C.c_void_p))
    assert rc == 0
    ok, od, st = orc.Extractor(*params).extract(img, dump=True)
    assert kps[:n.value].tobytes() == ok.tobytes()
    po = 0
    for l, g in enumerate(st["geometry"]):
        lw, lh = int(g["width"]), int(g["height"])
        exp = st["pyramid"][po:po + lw * lh].reshape(lh, lw)
        po += lw * lh
        got = bufs[l][:, :lw + 2 * border]
        want = orc.border_reflect101(np.ascontiguousarray(exp), border) if border else exp
        assert np.array_equal(got, want), "level %d" % l
        if extra_stride:
            assert (bufs[l][:, lw + 2 * border:] == 0xCD).all(), "bytes beyond the view were written"
    ex.close()


def test_fused_pyramid_tail_is_bit_identical(monkeypatch):
    """SDORB_PYRAMID_TAIL=1 (the upper pyramid levels in one launch; measured slower, so opt-in): same pyramid, same keypoints."""
    monkeypatch.setenv("SDORB_PYRAMID_TAIL", "1")
    for (w, h, params) in ((640, 480, C1), (333, 257, (700, 1.2, 6, 12)), (960, 540, (4000, 1.2, 12, 20))):
        imgs = synth.frames(3, w, h, start=650)
        ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=3)
        k, d, c = ex.extract_batch_host(imgs)
        before = ex.kernel_launches()
        ex(imgs[0], want_pyramid=False)
        assert ex.kernel_launches() - before < (params[2] - 1) + 5  # fewer launches than one per level + the five other kernels
        o = orc.Extractor(*params)
        for f in range(3):
            ok, od = o.extract(imgs[f])
            assert_same(ok, od, k[f, :c[f]], d[f, :c[f]], "%dx%d frame %d" % (w, h, f))
        k1, d1, pyr = ex(imgs[0])
        ok, od, st = o.extract(imgs[0], dump=True)
        po = 0
        for l, g in enumerate(st["geometry"]):
            lw, lh = int(g["width"]), int(g["height"])
            assert np.array_equal(pyr[l], st["pyramid"][po:po + lw * lh].reshape(lh, lw)), "level %d" % l
            po += lw * lh
        ex.close()


def test_single_frame_call_with_and_without_graph(monkeypatch):
    """sdorb_extract replays two captured CUDA graphs per call (pyramid | FAST .. describe + result copies); SDORB_GRAPH=0
    enqueues the same sequence on the streams.  Same bytes either way, over changing images, sizes and pyramid on / off."""
    o = orc.Extractor(*C1)
    outs = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("SDORB_GRAPH", mode)
        ex = api.ORBextractor(*C1, max_width=752, max_height=480, max_batch=1)
        res = []
        for i, (w, h, want) in enumerate([(640, 480, True), (640, 480, False), (640, 480, True), (752, 480, True), (752, 480, False),
                                          (640, 480, True), (320, 240, True)]):
            img = synth.smooth_noise(700 + i, w, h)
            k, d, pyr = ex(img, want_pyramid=want)
            ok, od, st = o.extract(img, dump=True)
            assert_same(ok, od, k, d, "graph=%s call %d" % (mode, i))
            if want:
                po = 0
                for l, g in enumerate(st["geometry"]):
                    lw, lh = int(g["width"]), int(g["height"])
                    assert np.array_equal(pyr[l], st["pyramid"][po:po + lw * lh].reshape(lh, lw)), "graph=%s call %d level %d" % (mode, i, l)
                    assert np.array_equal(pyr[l].base, orc.border_reflect101(np.ascontiguousarray(pyr[l]), 19))
                    po += lw * lh
            res.append((k.tobytes(), d.tobytes()))
        assert ex.kernel_launches() > 0
        outs[mode] = res
        ex.close()
    assert outs["1"] == outs["0"]


@pytest.mark.parametrize("ndev", [2, 3])
def test_multi_handle_driver_equals_single_handle(ndev):
    """sdorb_extract_batch_multi (the C-level multi-GPU driver): handles on their own host threads over contiguous frame ranges
    of ONE host batch.  With a single B200 the handles share device 0 (two contexts' worth of streams running concurrently);
    with several devices each handle gets its own (N-GPU output == 1-GPU output byte for byte)."""
    import torch
    ngpu = torch.cuda.device_count()
    params = (500, 1.2, 6, 20)
    w, h, nf = 320, 240, 23
    imgs = synth.frames(nf, w, h, start=900)
    single = api.ORBextractor(*params, device=0, max_width=w, max_height=h, max_batch=5)
    off, fb = single.pyramid_layout(w, h)
    slab1 = np.zeros(nf * fb, np.uint8)
    k1, d1, c1 = single.extract_batch_host(imgs, pyramid=slab1, first_level=0)
    single.close()
    exs = [api.ORBextractor(*params, device=g % ngpu, max_width=w, max_height=h, max_batch=4) for g in range(ndev)]
    slab = np.zeros(nf * fb, np.uint8)
    k, d, c = api.extract_batch_multi(exs, imgs, pyramid=slab, first_level=0)
    for e in exs:
        e.close()
    assert np.array_equal(c, c1) and k.tobytes() == k1.tobytes() and d.tobytes() == d1.tobytes() and slab.tobytes() == slab1.tobytes()
    o = orc.Extractor(*params)
    for f in (0, nf // ndev, nf - 1):
        ok, od = o.extract(imgs[f])
        assert c[f] == len(ok) and k[f, :c[f]].tobytes() == ok.tobytes() and np.array_equal(d[f, :c[f]], od)


def test_every_device_gives_identical_bytes():
    """Hardware N-GPU == 1-GPU identity: the same frames extracted on every visible device (skips below two devices)."""
    import torch
    ngpu = torch.cuda.device_count()
    if ngpu < 2:
        pytest.skip("needs at least two GPUs")
    imgs = synth.frames(16, start=40)
    ref = None
    for g in range(ngpu):
        ex = api.ORBextractor(*C1, device=g, max_width=640, max_height=480, max_batch=16)
        out = ex.extract_batch_host(imgs)
        ex.close()
        blob = out[0].tobytes() + out[1].tobytes() + out[2].tobytes()
        ref = ref or blob
        assert blob == ref, "device %d differs from device 0" % g


def test_guarded_allocations_stay_intact():
    """The bounds-checked debug mode (compute-sanitizer is closed on this pool): a second pytest process runs the extraction,
    matcher and search suites with SDORB_GUARD=1 -- 256 KB guard bands around every device buffer of the library, payloads
    poisoned -- and tests/conftest.py checks all bands after every test (buffers freed meanwhile are checked as they go).  An
    out-of-bounds write of any kernel fails there by name; an out-of-bounds or uninitialised read shows up as a parity failure."""
    import subprocess
    import sys
    env = dict(os.environ, SDORB_GUARD="1")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sel = ("stages_match or odd_sizes or batch_host or ragged_tail or device_entry or strided or geometry_switch or zero_corner or fewer_corners "
           "or massive or extreme or pyramid or match_batch or match_greedy or hamming or distinctive or frame_post or undistort or multi_handle "
           "or single_frame")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-k", sel, os.path.join(root, "tests", "test_gpu_parity.py"),
                        os.path.join(root, "tests", "test_gpu_search.py"), os.path.join(root, "tests", "test_orbslam2_mode.py")],
                       env=env, cwd=root, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
    assert " passed" in r.stdout
    # and the mode really was on in that process
    probe = subprocess.run([sys.executable, "-c", "from sdslam_b200 import api; e = api.ORBextractor(10, 1.2, 2, 20, max_width=64, max_height=64, "
                            "max_batch=1); print('guard', e.guard_check())"], env=env, cwd=root, capture_output=True, text=True, timeout=300)
    assert "guard 0" in probe.stdout, probe.stdout + probe.stderr


def test_oversized_batches_are_rejected(ex_c1):
    """Entry points whose batch index rides on gridDim.y refuse more than SDORB_MAX_GRID_BATCH frames up front."""
    n = 65536
    kps = np.zeros((n, 1), api.KP_DTYPE)
    cnt = np.zeros(n, np.int32)
    with pytest.raises(api.SdorbError) as e:
        ex_c1.undistort_keypoints_batch(kps, cnt, (500., 500., 320., 240.), (0.1, 0., 0., 0.))
    assert e.value.code == -1  # SDORB_ERR_BAD_ARG
    # the matcher's batch index is on gridDim.x: 70000 tiny pairs in one call
    npairs = 70000
    rng = np.random.default_rng(3)
    A = rng.integers(0, 256, (npairs, 2, 32), dtype=np.uint8)
    Bm = rng.integers(0, 256, (npairs, 3, 32), dtype=np.uint8)
    nA = np.full(npairs, 2, np.int32)
    nB = np.full(npairs, 3, np.int32)
    got = ex_c1.match_batch(A, nA, Bm, nB)
    exp = orc.match_many(A, nA, Bm, nB)
    assert got.tobytes() == exp.tobytes()


# ------------------------------------------------------------------ the selection order: device nth_element vs libstdc++
def _std_nth_element(entries, nth):
    """Ground truth from the real std::nth_element (oracle retainBest with n = nth + 1 keeps the arrangement of the
    first nth + 1 elements; the tail order is checked through the full-permutation harness on the CPU)."""
    order = orc.retain_best_order((entries & 0xFF).astype(np.float32), nth + 1)
    return entries[order]


def test_warp_nth_element_equals_std(ex_c1):
    """The warp-parallel Hoare partition must leave the first nth + 1 elements exactly where libstdc++ leaves them
    (that prefix, in that order, is what retainBest + resize keeps)."""
    rng = np.random.default_rng(123)
    cases = []
    for n in (2, 3, 4, 5, 8, 31, 32, 33, 64, 100, 257, 400, 777, 1024):
        for distinct in (1, 2, 5, 40, 255):
            e = (np.arange(n, dtype=np.uint32) << 8) | rng.integers(1, distinct + 1, n).astype(np.uint32)
            for nth in sorted({0, n // 7, n // 2, n - 2} & set(range(n - 1))):
                cases.append((e, nth))
    base = np.arange(600, dtype=np.uint32)
    for pat in (base % 200 + 1, (600 - base) % 200 + 1, np.minimum(base, 600 - base) % 250 + 1):
        cases.append(((base << 8) | pat.astype(np.uint32), 150))
        cases.append(((base << 8) | pat.astype(np.uint32), 3))
    big = (np.arange(3000, dtype=np.uint32) << 8) | rng.integers(1, 60, 3000).astype(np.uint32)  # > shared capacity: one-lane path
    cases.append((big, 700))
    for e, nth in cases:
        got = ex_c1.debug_nth_element(e, nth)
        exp = _std_nth_element(e, nth)
        assert sorted(got.tolist()) == sorted(e.tolist())
        assert np.array_equal(got[:nth + 1], exp[:nth + 1]), "n=%d nth=%d" % (len(e), nth)


# ------------------------------------------------------------------ full-size, size-independent properties
def test_full_size_batch_properties():
    """256 frames of the C3 shape: every frame yields exactly 1000 keypoints; the run is idempotent; a frame's
    result does not depend on its batch neighbours; all keypoints respect the 19-px edge and level-major order."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    base = synth.frames(8, 640, 480, start=200)
    imgs = torch.from_numpy(np.concatenate([base] * 32)).to(dev)
    ex = api.ORBextractor(*C1, max_width=640, max_height=480, max_batch=256)
    cap = ex.max_keypoints
    outs = []
    for _ in range(2):
        k = torch.zeros((256, cap, 7), dtype=torch.float32, device=dev)
        d = torch.zeros((256, cap, 32), dtype=torch.uint8, device=dev)
        c = torch.zeros(256, dtype=torch.int32, device=dev)
        ex.extract_batch_device(imgs, k, d, c)
        torch.cuda.synchronize()
        ex.batch_status()
        outs.append((k.cpu().numpy(), d.cpu().numpy(), c.cpu().numpy()))
    (k1, d1, c1), (k2, d2, c2) = outs
    assert (c1 == 1000).all()
    assert k1.tobytes() == k2.tobytes() and d1.tobytes() == d2.tobytes() and np.array_equal(c1, c2)
    for f in range(8, 256):
        assert k1[f].tobytes() == k1[f % 8].tobytes() and d1[f].tobytes() == d1[f % 8].tobytes()
    o = orc.Extractor(*C1)
    for f in (0, 5):
        ok, od = o.extract(base[f])
        assert_same(ok, od, kp_view(k1[f]).reshape(-1)[:1000], d1[f, :1000], "frame %d" % f)
    kv = kp_view(k1).reshape(256, cap)
    assert (np.diff(kv["octave"], axis=1) >= 0).all()
    sf = ex.GetScaleFactors()
    x0, y0 = kv["x"] / sf[kv["octave"]], kv["y"] / sf[kv["octave"]]
    assert (x0 > 18.99).all() and (y0 > 18.99).all()
    ex.close()


# ------------------------------------------------------------------ ORBmatcher::DescriptorDistance batched
def test_hamming_matrix_golden(ex_c1):
    g = np.load(os.path.join(GOLD, "hamming_96x80.npz"))
    assert np.array_equal(ex_c1.hamming_matrix(g["A"], g["B"]), g["dist"])
    m = api.ORBmatcher(ex_c1)
    assert m.DescriptorDistance(g["A"][7], g["B"][5]) == 0
    assert m.DescriptorDistance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_match_batch_equals_oracle_ragged(ex_c1):
    rng = np.random.default_rng(5)
    npairs, sa, sb = 7, 300, 260
    A = rng.integers(0, 256, (npairs, sa, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (npairs, sb, 32), dtype=np.uint8)
    for p in range(npairs):  # near-duplicates so that some matches pass TH_LOW and the ratio test
        idx = rng.integers(0, sb, 40)
        flip = rng.integers(0, 256, (40, 32), dtype=np.uint8) & rng.integers(0, 256, (40, 32), dtype=np.uint8) & rng.integers(0, 256, (40, 32), dtype=np.uint8) & 0x11
        A[p, :40] = B[p, idx] ^ flip
    B[2, 17] = B[2, 3]  # exact duplicate train rows: first index wins
    A[2, 0] = B[2, 3]
    nA = np.array([300, 0, 299, 1, 128, 129, 257], np.int32)
    nB = np.array([260, 100, 259, 260, 0, 1, 2], np.int32)
    for ratio, th in ((0.75, 50), (0.6, 100), (0.9, 256)):
        got = ex_c1.match_batch(A, nA, B, nB, ratio=ratio, th_low=th)
        exp = orc.match_many(A, nA, B, nB, ratio=ratio, th_low=th)
        for p in range(npairs):
            assert got[p, :nA[p]].tobytes() == exp[p, :nA[p]].tobytes(), "pair %d ratio %g" % (p, ratio)
    got = ex_c1.match_batch(A, nA, B, nB)
    assert got[2, 0]["best_idx"] == 3 and got[2, 0]["best_dist"] == 0 and got[2, 0]["second_dist"] == 0
    assert (got[4, :128]["best_idx"] == -1).all() and (got[4, :128]["best_dist"] == 256).all()
    assert got["accepted"].sum() > 50


def test_match_greedy_equals_oracle(ex_c1):
    """SearchByPoints' vbMatched2 rule (src/ORBmatcher.cc:1228-1270)."""
    rng = np.random.default_rng(6)
    A = rng.integers(0, 256, (3, 200, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (3, 180, 32), dtype=np.uint8)
    A[:, :60] = B[:, rng.integers(0, 180, 60)] ^ (rng.integers(0, 256, (3, 60, 32), dtype=np.uint8) & 0x21)
    A[1, 100:130] = A[1, 10:40]  # several queries want the same train row
    nA = np.array([200, 200, 150], np.int32)
    nB = np.array([180, 170, 180], np.int32)
    got = ex_c1.match_batch(A, nA, B, nB, ratio=0.9, th_low=60, greedy=True)
    for p in range(3):
        exp = orc.match_best2(A[p, :nA[p]], B[p, :nB[p]], ratio=0.9, th_low=60, greedy=True)
        assert got[p, :nA[p]].tobytes() == exp.tobytes()
    plain = ex_c1.match_batch(A, nA, B, nB, ratio=0.9, th_low=60)
    assert plain.tobytes() != got.tobytes()  # the greedy mask changed something


def test_match_full_size_properties(ex_c1):
    """1000 x 1000 (the BASELINE C4 shape) on real ORB descriptors: best/second-best agree with the distance matrix,
    and matching a set against itself returns the identity with distance 0."""
    imgs = synth.frames(2, 640, 480, start=300)
    k, d, c = ex_c1.extract_batch_host(imgs)
    assert (c == 1000).all()
    A, B = d[0:1], d[1:2]
    n = np.array([1000], np.int32)
    m = ex_c1.match_batch(A, n, B, n)[0]
    dist = ex_c1.hamming_matrix(A[0], B[0]).astype(np.int32)
    assert np.array_equal(dist, orc.hamming_matrix(A[0], B[0]))
    assert np.array_equal(dist, ex_c1.hamming_matrix(B[0], A[0]).T)
    assert np.array_equal(m["best_idx"], dist.argmin(axis=1)) and np.array_equal(m["best_dist"], dist.min(axis=1))
    assert np.array_equal(m["second_dist"], np.sort(dist, axis=1)[:, 1])
    assert m.tobytes() == orc.match_best2(A[0], B[0]).tobytes()
    self_m = ex_c1.match_batch(A, n, A, n)[0]
    dself = ex_c1.hamming_matrix(A[0], A[0])
    assert (np.diag(dself) == 0).all() and (self_m["best_dist"] == 0).all()
    assert np.array_equal(self_m["best_idx"], dself.argmin(axis=1))


def test_distinctive_descriptors_equal_oracle(ex_c1):
    """MapPoint::ComputeDistinctiveDescriptors batched (src/MapPoint.cc:225-284): ragged sets incl. empty, single,
    duplicates, a set larger than one CTA's threads, and real ORB descriptors."""
    rng = np.random.default_rng(31)
    sizes = [1, 2, 3, 0, 7, 16, 64, 65, 130, 300, 5, 0, 33]
    sets = []
    for n in sizes:
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if n >= 5:
            d[1:n // 2] = d[0] ^ (rng.integers(0, 256, (n // 2 - 1, 32), dtype=np.uint8) & 0x09)
            d[n - 1] = d[2]
        sets.append(d)
    k, dd, c = ex_c1.extract_batch_host(synth.frames(1, 640, 480, start=400))
    sets.append(dd[0, :200].copy())
    sizes.append(200)
    desc = np.concatenate(sets)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    gi, gm = ex_c1.distinctive_batch(desc, offsets)
    oi, om = orc.distinctive_many(desc, offsets)
    assert np.array_equal(gi, oi) and np.array_equal(gm, om)
    assert gi[3] == -1 and gi[0] == 0 and gm[0] == 0
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    td, to = torch.from_numpy(desc).to(dev), torch.from_numpy(offsets).to(dev)
    ti = torch.zeros(len(sizes), dtype=torch.int32, device=dev)
    tm = torch.zeros(len(sizes), dtype=torch.int32, device=dev)
    ex_c1.distinctive_batch(td, to, ti, tm, device=True)
    torch.cuda.synchronize()
    assert np.array_equal(ti.cpu().numpy(), oi) and np.array_equal(tm.cpu().numpy(), om)


def test_frame_postprocessing_equals_oracle(ex_c1):
    """Frame::AssignFeaturesToGrid / PosInGrid and ComputeStereoFromRGBD (src/Frame.cc:179-192, 323-332, 399-417) on the
    extractor's own output (no distortion: mvKeysUn = mvKeys) plus keypoints pushed outside the image bounds."""
    imgs = synth.frames(3, 640, 480, start=500)
    imgs[2] = 0  # a frame without keypoints
    kps, desc, cnt = ex_c1.extract_batch_host(imgs)
    kun = kps.copy()
    kun["x"][1] += np.float32(31.3)  # as if undistortion had moved them: some leave the grid on the right
    kun["y"][1] -= np.float32(27.6)  # ... and at the top
    inv_w, inv_h = float(np.float32(64) / np.float32(640)), float(np.float32(48) / np.float32(480))
    cs, idx = ex_c1.assign_grid_batch(kun, cnt, 0.0, 0.0, inv_w, inv_h)
    for f in range(3):
        ocs, oidx = orc.assign_grid(kun[f, :cnt[f]], 0.0, 0.0, inv_w, inv_h)
        assert np.array_equal(cs[f], ocs), "frame %d cell_start" % f
        assert np.array_equal(idx[f, :ocs[-1]], oidx), "frame %d indices" % f
    assert cs[2, -1] == 0 and 0 < cs[1, -1] < cnt[1] and cs[0, -1] == cnt[0]
    rng = np.random.default_rng(7)
    depth = rng.uniform(-2, 9, (3, 480, 640)).astype(np.float32)
    depth[depth < 0.3] = 0
    ur, z = ex_c1.stereo_from_rgbd_batch(kps, kun, cnt, depth, 40.0)
    for f in range(3):
        our, oz = orc.stereo_from_rgbd(kps[f, :cnt[f]], kun[f, :cnt[f]], depth[f], 40.0)
        assert ur[f, :cnt[f]].tobytes() == our.tobytes() and z[f, :cnt[f]].tobytes() == oz.tobytes()
        assert (ur[f, cnt[f]:] == -1).all() and (z[f, cnt[f]:] == -1).all()


@pytest.mark.parametrize("cam", ["tum1", "euroc", "none"])
def test_undistort_keypoints_equals_oracle(ex_c1, cam):
    """Frame::UndistortKeyPoints (src/Frame.cc:335-366) on the extractor's output; the oracle itself is pinned against the
    real cv2.undistortPoints (tests/test_oracle_primitives.py)."""
    from test_oracle_primitives import CAMERAS
    K4, dist = (np.array(v, np.float32) for v in CAMERAS[cam])
    kps, desc, cnt = ex_c1.extract_batch_host(synth.frames(2, 752, 480, start=600))
    out = ex_c1.undistort_keypoints_batch(kps, cnt, K4, dist)
    for f in range(2):
        ref = orc.undistort_keypoints(kps[f, :cnt[f]], K4, dist)
        assert out[f, :cnt[f]].tobytes() == ref.tobytes()
    moved = float(np.abs(out["x"][0, :cnt[0]] - kps["x"][0, :cnt[0]]).max())
    assert (moved > 0.5) == (cam != "none")
    assert api.host_image_bounds(752, 480, K4, dist).tobytes() == orc.image_bounds(752, 480, K4, dist).tobytes()


def test_kernel_launch_accounting(ex_c1):
    before = ex_c1.kernel_launches()
    ex_c1(synth.smooth_noise(0), want_pyramid=False)
    assert ex_c1.kernel_launches() - before >= 5
    ex_c1.set_profiling(True)
    ex_c1.stage_times(reset=True)
    ex_c1(synth.smooth_noise(0), want_pyramid=False)
    ms, launches = ex_c1.stage_times()
    ex_c1.set_profiling(False)
    assert all(ms[s] > 0 for s in ("pyramid", "fast", "select", "blur", "describe"))
    assert launches["fast"] >= 1 and launches["pyramid"] >= 1
