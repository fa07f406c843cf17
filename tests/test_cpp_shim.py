"""include/ORBextractor.h -- the C++ drop-in for SD_SLAM::ORBextractor (/root/reference/src/ORBextractor.h:34-90) -- is
compiled as C++11 (the reference's standard, CMakeLists.txt:40) against a minimal cv:: mock (OpenCV's C++ headers are
not in this image) and, on the GPU box, driven the way Frame.cc:195 drives the reference and compared with the oracle."""
import os
import struct
import subprocess

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import build as sbuild
from sdslam_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# The shim is compiled against TWO independently written cv:: surfaces (OpenCV's own C++ headers are not in this image): the
# 100-line mock of tests/helpers/cv_mock and oracle/ref_compat, the surface the reference's own sources are compiled against for
# the parity library (ref-counted Mat with ROI views, InputArray / OutputArray proxies, CV_Assert throwing cv::Exception).
CV_SURFACES = {"cv_mock": os.path.join(ROOT, "tests", "helpers", "cv_mock"), "ref_compat": os.path.join(ROOT, "oracle", "ref_compat")}


@pytest.fixture(scope="module", params=sorted(CV_SURFACES))
def driver(request, tmp_path_factory):
    so = sbuild.build()
    exe = str(tmp_path_factory.mktemp("shim_" + request.param) / "shim_driver")
    subprocess.check_call(["g++", "-std=c++11", "-Wall", "-Wextra", "-Werror", "-O1", "-I", os.path.join(ROOT, "include"),
                           "-I", CV_SURFACES[request.param], "-o", exe,
                           os.path.join(ROOT, "tests", "helpers", "shim_driver.cc"), so,
                           "-Wl,-rpath," + os.path.dirname(so)])
    return exe


def test_shim_compiles_and_keeps_the_reference_surface(driver):
    assert os.path.exists(driver)
    hdr = open(os.path.join(ROOT, "include", "ORBextractor.h")).read()
    for needle in ("namespace SD_SLAM", "class ORBextractor", "enum { HARRIS_SCORE = 0, FAST_SCORE = 1 }",
                   "ORBextractor(int _nfeatures, float _scaleFactor, int _nlevels, int _thFAST",
                   "std::vector<cv::Mat>& imagePyramid", "GetLevels()", "GetScaleFactor()", "GetScaleFactors()",
                   "GetInverseScaleFactors()", "GetScaleSigmaSquares()", "GetInverseScaleSigmaSquares()"):
        assert needle in hdr, needle


@pytest.mark.gpu
@pytest.mark.parametrize("w,h,params", [(640, 480, (1000, 1.2, 8, 20)), (320, 240, (500, 2.0, 3, 20)),
                                        (640, 480, (1000, 1.2, 8, 20, 7))])
def test_shim_driver_equals_oracle(driver, tmp_path, w, h, params):
    """Four constructor arguments: the reference; five (iniThFAST, minThFAST): the ORB-SLAM2-style mode (row f1)."""
    img = synth.smooth_noise(77, w, h)
    raw, out = str(tmp_path / "in.raw"), str(tmp_path / "out.bin")
    img.tofile(raw)
    subprocess.check_call([driver, raw, str(w), str(h)] + [str(p) for p in params[:4]] + [out] + [str(p) for p in params[4:]])
    buf = open(out, "rb").read()
    off = 0
    (n,) = struct.unpack_from("<i", buf, off)
    off += 4
    k = np.frombuffer(buf, orc.KP_DTYPE, n, off)
    off += 28 * n
    d = np.frombuffer(buf, np.uint8, 32 * n, off).reshape(n, 32)
    off += 32 * n
    o = orc.Extractor(*params[:4], min_th_fast=params[4] if len(params) > 4 else None)
    ok, od, st = o.extract(img, dump=True)
    assert n == len(ok) and k.tobytes() == ok.tobytes() and np.array_equal(d, od)
    (nl,) = struct.unpack_from("<i", buf, off)
    off += 4
    assert nl == params[2]
    poff = 0
    for l in range(nl):
        lw, lh = struct.unpack_from("<ii", buf, off)
        off += 8
        padded = np.frombuffer(buf, np.uint8, (lw + 38) * (lh + 38), off).reshape(lh + 38, lw + 38)
        off += padded.size
        inner = st["pyramid"][poff:poff + lw * lh].reshape(lh, lw)
        poff += lw * lh
        assert np.array_equal(padded, orc.border_reflect101(inner, 19)), "pyramid level %d" % l
    t = o.tables()
    for name in ("scale", "inv_scale", "sigma2", "inv_sigma2"):
        assert buf[off:off + 4 * nl] == t[name].tobytes(), name
        off += 4 * nl
    (dd,) = struct.unpack_from("<i", buf, off)
    off += 4
    assert dd == orc.descriptor_distance(od[0], od[1])
    m = np.frombuffer(buf, orc.MATCH_DTYPE, 1, off)
    assert m.tobytes() == orc.match_best2(od[0:1], od).tobytes()
