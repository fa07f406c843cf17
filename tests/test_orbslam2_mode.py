"""The ORB-SLAM2-style mode (iniThFAST / minThFAST fallback on 30-pixel cells +
DistributeOctTree).  This mode is NOT in /root/reference; the oracle restates the public ORB-SLAM2
algorithm (oracle/sdorb_oracle.cc: compute_keypoints_octree, distribute_oct_tree -- literal std::list form) and is pinned
here against a second restatement written separately (tests/cv2_pipeline.py: real cv2.FAST per cell, pass-based
DistributeOctTree on arrays) -- live and through the committed fixtures tests/golden/orbslam2/*.npz
(tests/golden/make_orbslam2_golden.py).  Parity with ORB-SLAM2 itself is UNPINNED (its source is not in this image).
CPU tests first; the GPU parity tests (marked gpu) call the library through its C ABI."""
import glob
import os

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import api, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "orbslam2")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")))


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = g["params"]
    return g, (int(p[0]), float(p[1]), int(p[2]), int(p[3]), int(p[4]))


def test_fixture_inventory():
    assert len(CASES) >= 6


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_golden(name):
    g, params = load_case(name)
    k, d = orc.Extractor(*params[:4], min_th_fast=params[4]).extract(g["image"])
    assert k.tobytes() == g["kps"].astype(orc.KP_DTYPE).tobytes()
    assert np.array_equal(d, g["desc"].reshape(-1, 32))


def test_fixtures_cover_the_regimes():
    g, p = load_case("os2_smooth_640x480")
    k = g["kps"]
    assert 1000 <= len(k) <= 1000 + 2 * 8  # every level stops at N .. N + 2 nodes
    g, p = load_case("os2_few_320x240")
    assert len(g["kps"]) < p[0] // 4  # fewer corners than wanted: every node ends with one keypoint
    g, p = load_case("os2_fallback_480x360")
    resp = g["kps"]["response"]
    assert (resp < p[3]).any() and (resp >= p[3]).any() and (resp >= p[4]).all()  # cells served by minThFAST and by iniThFAST
    g, p = load_case("os2_wide_620x188")
    assert len(g["kps"]) >= p[0]


def test_oracle_equals_cv2_pipeline_live():
    cvp = pytest.importorskip("cv2_pipeline")
    img = synth.smooth_noise(31, 240, 176)
    params = (200, 1.2, 3, 20, 7)
    k, d, _ = cvp.extract(img, *params[:4], min_th_fast=params[4])
    ok, od = orc.Extractor(*params[:4], min_th_fast=params[4]).extract(img)
    assert len(k) == len(ok) >= 200
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


def test_distribute_oct_tree_restatements_agree_on_adversarial_sets():
    """The std::list form (oracle) and the pass-based array form (cv2_pipeline.distribute_oct_tree, the form the CUDA kernel
    uses) on tie-heavy inputs: equal responses, clustered points, sets smaller and much larger than N, several initial nodes."""
    cvp = pytest.importorskip("cv2_pipeline")
    rng = np.random.default_rng(3)
    for trial in range(60):
        w, h = int(rng.integers(40, 700)), int(rng.integers(40, 300))
        if round(w / h) < 1:
            w, h = h, w
        n = int(rng.integers(1, 1500))
        if trial % 3 == 0:  # clustered
            cx, cy = rng.integers(0, w, 6), rng.integers(0, h, 6)
            pick = rng.integers(0, 6, n)
            xs = np.clip(cx[pick] + rng.integers(-9, 10, n), 0, w - 1)
            ys = np.clip(cy[pick] + rng.integers(-9, 10, n), 0, h - 1)
        else:
            xs, ys = rng.integers(0, w, n), rng.integers(0, h, n)
        pts = sorted(set(zip(ys.tolist(), xs.tolist())))  # distinct pixels, row-major like FAST
        resp = rng.integers(7, 12 if trial % 2 else 60, len(pts))
        keys = np.zeros(len(pts), orc.KP_DTYPE)
        keys["x"], keys["y"] = [p[1] for p in pts], [p[0] for p in pts]
        keys["response"] = resp
        N = int(rng.integers(0, 400))
        out = orc.distribute_oct_tree(keys, 16, 16 + w, 16, 16 + h, N)
        ref = cvp.distribute_oct_tree([(float(k["x"]), float(k["y"]), float(k["response"])) for k in keys], 16, 16 + w, 16, 16 + h, N)
        got = [(float(k["x"]), float(k["y"]), float(k["response"])) for k in out]
        assert got == ref, "trial %d (%d keys, N=%d, %dx%d)" % (trial, len(pts), N, w, h)


def test_max_keypoints_and_geometry_errors_without_gpu():
    g = api.host_level_geometry(1000, 1.2, 8, 20, 640, 480)  # reference mode untouched
    assert int(g["level_cols"][0]) == 5


# ---------------------------------------------------------------------------------------------------- GPU parity
def _assert_same(ok, od, gk, gd, what):
    assert len(ok) == len(gk), "%s: %d oracle vs %d gpu keypoints" % (what, len(ok), len(gk))
    for f in api.KP_DTYPE.names:
        bad = np.flatnonzero(ok[f] != gk[f])
        assert len(bad) == 0, "%s: kp.%s differs at %s (first: oracle %r gpu %r)" % (what, f, bad[:5], ok[f][bad[0]], gk[f][bad[0]])
    assert np.array_equal(od, gd), "%s: descriptors differ" % what


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_gpu_equals_golden(name):
    g, params = load_case(name)
    img = g["image"]
    ex = api.ORBextractor(*params[:4], minThFAST=params[4], max_width=img.shape[1], max_height=img.shape[0], max_batch=2)
    k, d, _ = ex(img)
    ex.close()
    _assert_same(g["kps"].astype(api.KP_DTYPE), g["desc"].reshape(-1, 32), k, d, name)


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,params", [(640, 480, (1000, 1.2, 8, 20, 7)), (752, 480, (1000, 1.2, 8, 20, 7)),
                                        (1241, 376, (2000, 1.2, 8, 20, 7)), (641, 479, (1500, 1.2, 8, 12, 12)),
                                        (320, 240, (500, 1.2, 4, 7, 20)), (1920, 1080, (4000, 1.2, 8, 20, 7))])
def test_gpu_batch_equals_oracle(w, h, params):
    """Batches of mixed content (noise-dense, sparse rectangles, constant, low contrast): the batch equals frame-by-frame
    oracle output; ini < min (the fallback can add nothing) and ini == min included."""
    nf = 3 if w >= 1920 else 6
    imgs = np.stack([synth.smooth_noise(70 + i, w, h) if i % 3 == 0 else synth.rects(70 + i, w, h) for i in range(nf)])
    imgs[-1] = 90
    rng = np.random.default_rng(1)
    imgs[-2] = (120 + rng.normal(0, 5, (h, w))).clip(0, 255).astype(np.uint8)
    ex = api.ORBextractor(*params[:4], minThFAST=params[4], max_width=w, max_height=h, max_batch=4)
    kps, desc, cnt = ex.extract_batch_host(imgs)
    o = orc.Extractor(*params[:4], min_th_fast=params[4])
    for f in range(nf):
        ok, od = o.extract(imgs[f])
        _assert_same(ok, od, kps[f, :cnt[f]], desc[f, :cnt[f]], "%dx%d frame %d" % (w, h, f))
    assert cnt[-1] == 0 and cnt[0] >= params[0]
    k1, d1, _ = ex(imgs[0])
    assert k1.tobytes() == kps[0, :cnt[0]].tobytes()
    assert ex.max_keypoints >= int(cnt.max())
    ex.close()


@pytest.mark.gpu
def test_gpu_mode_switch_keeps_reference_mode_intact():
    img = synth.smooth_noise(3)
    a = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=1)
    b = api.ORBextractor(1000, 1.2, 8, 20, minThFAST=7, max_width=640, max_height=480, max_batch=1)
    ka, da, _ = a(img)
    kb, db, _ = b(img)
    ok, od = orc.Extractor(1000, 1.2, 8, 20).extract(img)
    assert ka.tobytes() == ok.tobytes() and np.array_equal(da, od)
    assert kb.tobytes() != ka.tobytes() and len(kb) >= 1000
    a.close()
    b.close()
