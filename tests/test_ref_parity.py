"""The oracle restatement against the REFERENCE ITSELF: /root/reference/src/ORBextractor.cc compiled unmodified (and
ORBmatcher::DescriptorDistance) into oracle/_ref/libsdorb_ref.so by oracle/ref_build/Makefile, on the cv:: surface of
oracle/ref_compat whose pixel primitives are the cv2-4.13-pinned ones.  What this pins: every line of the path's control flow
and float arithmetic as g++ compiles the reference's own text with the reference's own flags -- the cell grid and its
float -> int ROI conversions, quota redistribution, both retainBest passes and their order, IC_Angle, the rBRIEF index math and
its FMA contraction, keypoint scaling, the chained pyramid and its borders, the constructor tables.  CPU only; skipped where
neither /root/reference nor a prebuilt oracle/_ref exists.
"""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import binding as orc
from oracle import ref_binding as ref
from sdslam_b200 import synth

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")) if "hamming" not in p)
C1 = (1000, 1.2, 8, 20)


def same_extraction(params, img, what, native=False):
    o = orc.Extractor(*params)
    r = ref.Extractor(*params, native=native)
    ok, od, st = o.extract(img, dump=True)
    rk, rd, lv, padded = r.extract(img, pyramid=True)
    assert len(ok) == len(rk), "%s: %d oracle vs %d reference keypoints" % (what, len(ok), len(rk))
    for f in orc.KP_DTYPE.names:
        bad = np.flatnonzero(ok[f] != rk[f])
        assert len(bad) == 0, "%s: kp.%s differs at %s" % (what, f, bad[:5])
    assert ok.tobytes() == rk.tobytes(), what
    assert np.array_equal(od, rd), what
    off = 0
    for l, g in enumerate(st["geometry"]):
        w, h = int(g["width"]), int(g["height"])
        assert lv[l].shape == (h, w), "%s: level %d size" % (what, l)
        assert np.array_equal(lv[l], st["pyramid"][off:off + w * h].reshape(h, w)), "%s: pyramid level %d" % (what, l)
        assert np.array_equal(padded[l], orc.border_reflect101(np.ascontiguousarray(lv[l]), 19)), "%s: border of level %d" % (what, l)
        off += w * h
    return len(ok)


def test_constructor_tables_equal_reference():
    """src/ORBextractor.cc:406-457 for a spread of parameter sets (float scale tables, per-level split, umax, pattern)."""
    pat = orc.pattern() if hasattr(orc, "pattern") else None
    for params in [C1, (1000, 2.0, 5, 20), (2000, 1.2, 8, 20), (4000, 1.2, 12, 20), (500, 1.1, 16, 7), (1, 1.2, 8, 20), (0, 1.2, 3, 20),
                   (7, 1.5, 2, 20), (50, 1.2, 1, 20), (300, 1.05, 12, 5), (1234, 1.3333, 9, 33), (100000, 1.2, 8, 1)]:
        t, r = orc.Extractor(*params).tables(), ref.Extractor(*params).tables()
        assert r["nlevels"] == params[2] and r["scale_factor"] == np.float32(params[1])
        for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "n_per_level", "umax"):
            assert t[k].tobytes() == r[k].tobytes(), (params, k)
        if pat is not None:
            assert np.array_equal(r["pattern"], pat)


@pytest.mark.parametrize("name", CASES)
def test_golden_fixture_equals_reference(name):
    """The committed fixtures are what the reference itself produces (they are regenerated from it: tests/golden/make_golden.py)."""
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = g["params"]
    params = (int(p[0]), float(p[1]), int(p[2]), int(p[3]))
    k, d, lv, _ = ref.Extractor(*params).extract(g["image"], pyramid=True)
    assert k.tobytes() == g["kps"].astype(orc.KP_DTYPE).tobytes()
    assert np.array_equal(d, g["desc"].reshape(-1, 32))
    for l, ((h, w), sha) in enumerate(zip(g["pyramid_shape"], g["pyramid_sha256"])):
        assert lv[l].shape == (int(h), int(w))
        assert hashlib.sha256(np.ascontiguousarray(lv[l]).tobytes()).hexdigest() == str(sha)
    same_extraction(params, g["image"], name)


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
def test_staged_configs_equal_reference(label, img, params):
    """The ten configurations the GPU stage-by-stage test uses (tests/test_gpu_parity.py::test_stages_match_oracle)."""
    same_extraction(params, img, label)
    same_extraction(params, img, label + " (-march=native build)", native=ref._native_runs_here())


def test_c3_frames_equal_reference():
    """256 frames of the bench workload (C3: 640x480, 1000 kp), frame-parallel; tests/tools/ref_sweep.py runs 4096 + the other
    shapes (log under profiles/)."""
    imgs = synth.frames(256)
    n = os.cpu_count() or 1
    ok, od, oc = orc.Extractor(*C1).extract_many(imgs, nthreads=n)
    rk, rd, rc = ref.Extractor(*C1).extract_many(imgs, nthreads=n)
    assert np.array_equal(oc, rc) and (rc == 1000).all()
    assert ok.tobytes() == rk.tobytes() and od.tobytes() == rd.tobytes()


def test_degenerate_and_edge_inputs_equal_reference():
    rng = np.random.default_rng(7)
    cases = [
        ("constant", np.full((120, 160), 77, np.uint8), (500, 1.2, 8, 20)),
        ("checker", (((np.mgrid[0:240, 0:320][1] // 9 + np.mgrid[0:240, 0:320][0] // 9) & 1) * 200 + 20).astype(np.uint8), (500, 1.2, 8, 20)),
        ("ties", (rng.integers(0, 2, (200, 260)) * 255).astype(np.uint8), (300, 1.2, 4, 20)),
        ("few_corners", synth.rects(5, 320, 240), (2000, 1.2, 6, 20)),
        ("nfeatures_1", synth.smooth_noise(9, 200, 160), (1, 1.2, 8, 20)),
        ("nfeatures_7", synth.smooth_noise(9, 200, 160), (7, 1.2, 2, 20)),
        ("one_level", synth.smooth_noise(10, 200, 160), (50, 1.2, 1, 20)),
        ("fine_scale", synth.smooth_noise(11, 260, 200), (300, 1.05, 12, 5)),
        ("wide", synth.smooth_noise(12, 1241, 376), (2000, 1.2, 8, 20)),
        ("strided_view", synth.smooth_noise(13, 400, 300)[7:247, 11:331], (600, 1.2, 6, 20)),
    ]
    for label, img, params in cases:
        same_extraction(params, img, label)


def test_random_sizes_and_parameters_equal_reference():
    """Fuzz over image sizes and parameter sets, including geometries where the reference throws (cell ROI outside a level):
    oracle and reference must agree on WHETHER it throws and on every byte otherwise."""
    rng = np.random.default_rng(2024)
    thrown = ran = 0
    for i in range(60):
        w, h = int(rng.integers(45, 420)), int(rng.integers(45, 330))
        params = (int(rng.integers(1, 1500)), float(rng.choice([1.1, 1.2, 1.25, 1.5, 2.0])), int(rng.integers(1, 9)), int(rng.integers(5, 60)))
        img = synth.smooth_noise(500 + i, w, h) if i % 3 else rng.integers(0, 256, (h, w), dtype=np.uint8)
        try:
            rk, rd = ref.Extractor(*params).extract(img)
        except RuntimeError:
            with pytest.raises(RuntimeError):
                orc.Extractor(*params).extract(img)
            thrown += 1
            continue
        try:
            ok, od = orc.Extractor(*params).extract(img)
        except RuntimeError:
            # the oracle also refuses geometries where the reference READS OUT OF BOUNDS without throwing (undefined behaviour:
            # a non-empty cell whose detectable rectangle passes maxBorder, oracle/sdorb_oracle.cc compute_keypoints)
            continue
        assert ok.tobytes() == rk.tobytes() and np.array_equal(od, rd), (w, h, params)
        ran += 1
    assert ran >= 20


def test_too_small_image_throws_in_both():
    with pytest.raises(RuntimeError):
        ref.Extractor(*C1).extract(np.zeros((30, 40), np.uint8))
    with pytest.raises(RuntimeError):
        orc.Extractor(*C1).extract(np.zeros((30, 40), np.uint8))


def test_empty_image_is_a_no_op_in_the_reference():
    """src/ORBextractor.cc:622-623: an empty image returns before anything is touched."""
    k, d = ref.Extractor(*C1).extract(np.zeros((0, 0), np.uint8))
    assert len(k) == 0 and len(d) == 0


def test_descriptor_distance_equals_reference():
    """src/ORBmatcher.cc:1459-1473, the reference's own SWAR loop, against the oracle, numpy bit counting and the fixture."""
    g = np.load(os.path.join(GOLD, "hamming_96x80.npz"))
    assert np.array_equal(ref.hamming_matrix(g["A"], g["B"]), g["dist"])
    rng = np.random.default_rng(11)
    A = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (150, 32), dtype=np.uint8)
    A[0] = 0
    B[0] = 255
    A[1] = B[1]
    m = ref.hamming_matrix(A, B)
    assert np.array_equal(m, orc.hamming_matrix(A, B))
    assert np.array_equal(m, np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2))
    assert m[0, 0] == 256 and m[1, 1] == 0
    assert ref.descriptor_distance(A[5], B[9]) == orc.descriptor_distance(A[5], B[9])
    if ref._native_runs_here():
        assert np.array_equal(ref.hamming_matrix(A, B, native=True), m)
