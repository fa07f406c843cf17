"""The whole oracle operator() against an independent second implementation assembled from real cv2 4.13 calls
(tests/cv2_pipeline.py): the committed fixtures under tests/golden/ (made by tests/golden/make_golden.py) and one
live run.  Follows /root/reference/src/ORBextractor.cc:620-700.  CPU only."""
import glob
import hashlib
import os

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden")
CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLD, "*.npz")) if "hamming" not in p)


def load_case(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    p = g["params"]
    return g, (int(p[0]), float(p[1]), int(p[2]), int(p[3]))


def test_fixture_inventory():
    assert len(CASES) >= 9


@pytest.mark.parametrize("name", CASES)
def test_oracle_equals_golden(name):
    g, params = load_case(name)
    k, d, st = orc.Extractor(*params).extract(g["image"], dump=True)
    assert k.tobytes() == g["kps"].astype(orc.KP_DTYPE).tobytes()
    assert np.array_equal(d, g["desc"].reshape(-1, 32))
    off = 0
    for (h, w), sha in zip(g["pyramid_shape"], g["pyramid_sha256"]):
        n = int(h) * int(w)
        assert hashlib.sha256(st["pyramid"][off:off + n].tobytes()).hexdigest() == str(sha)
        off += n


def test_golden_cover_reference_regimes():
    """The fixtures exercise: the reference defaults (scale 2.0 x 5 levels: a level with negative cell height), the
    2*nFeatures initialisation extractor, portrait aspect, zero-corner input, a level without corners / under quota."""
    g, p = load_case("c0_default_640x480")
    assert p == (1000, 2.0, 5, 20) and 0 < len(g["kps"]) <= 1000 and int(g["kps"]["octave"].max()) <= 3
    assert len(load_case("constant_160x120")[0]["kps"]) == 0
    assert len(load_case("c1_smooth_640x480")[0]["kps"]) == 1000
    assert len(load_case("c5_small_800x450")[0]["kps"]) == 4000
    chk = load_case("checker_320x240")[0]["kps"]
    # X-junctions are not FAST corners: level 0 is empty, and the total stays under the quota (391 < 500)
    assert 0 < len(chk) < 500 and (chk["octave"] > 0).all()


def test_oracle_equals_cv2_pipeline_live():
    cvp = pytest.importorskip("cv2_pipeline")
    img = synth.smooth_noise(21, 176, 144)
    params = (150, 1.2, 4, 15)
    k, d, pyr = cvp.extract(img, *params)
    ok, od, st = orc.Extractor(*params).extract(img, dump=True)
    assert len(k) == len(ok) > 100
    assert ok.tobytes() == k.tobytes() and np.array_equal(od, d)
    assert np.array_equal(st["pyramid"], np.concatenate([p.ravel() for p in pyr]))


def test_oracle_tables_c1():
    """the geometric split and the patch row ends."""
    t = orc.Extractor(1000, 1.2, 8, 20).tables()
    assert t["n_per_level"].tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert t["umax"].tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    g = orc.Extractor(1000, 1.2, 8, 20).geometry(640, 480)
    assert list(zip(g["width"].tolist(), g["height"].tolist())) == [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231),
                                                                     (257, 193), (214, 161), (179, 134)]
    assert list(zip(g["level_cols"].tolist(), g["level_rows"].tolist()))[:3] == [(5, 6), (5, 6), (4, 5)]
    assert g["scaled_patch_size"].tolist() == [31, 37, 44, 53, 64, 77, 92, 111]


def test_oracle_frame_parallel_driver_is_deterministic():
    imgs = synth.frames(5, 160, 120)
    e = orc.Extractor(200, 1.2, 4, 20)
    k1, d1, c1 = e.extract_many(imgs, nthreads=1)
    k4, d4, c4 = e.extract_many(imgs, nthreads=4)
    assert np.array_equal(c1, c4) and k1.tobytes() == k4.tobytes() and d1.tobytes() == d4.tobytes()
    for f in range(5):
        k, d = e.extract(imgs[f])
        assert c1[f] == len(k) and k1[f, :len(k)].tobytes() == k.tobytes() and np.array_equal(d1[f, :len(k)], d)


def test_oracle_geometry_error():
    """A level whose cell ROI leaves the image: the reference would throw cv::Exception (Mat::rowRange)."""
    with pytest.raises(RuntimeError):
        orc.Extractor(1000, 1.2, 8, 20).extract(np.zeros((30, 40), np.uint8))
