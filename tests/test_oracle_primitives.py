"""Pins every primitive of the CPU oracle (oracle/sdorb_oracle.cc) against the arithmetic of record:
Python cv2 4.13 (the only importable implementation of the OpenCV calls the reference makes,
/root/reference/src/ORBextractor.cc:101,536,586,602,660,690,692), glibc's sincosf and libstdc++'s nth_element.
CPU only.
"""
import ctypes
import os
import subprocess
import tempfile

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import synth

cv2 = pytest.importorskip("cv2")


def _images():
    rng = np.random.default_rng(11)
    yield "noise", rng.integers(0, 256, (97, 131), dtype=np.uint8)
    yield "smooth", synth.smooth_noise(4, 320, 240)
    yield "rects", synth.rects(2, 213, 167)
    yield "const", np.full((64, 80), 200, np.uint8)
    y, x = np.mgrid[0:120, 0:160]
    yield "checker", (((x // 7 + y // 5) & 1) * 220 + 10).astype(np.uint8)
    yield "ramp", ((x * 3 + y * 2) % 256).astype(np.uint8)


# ------------------------------------------------------------------ cv::resize INTER_LINEAR
@pytest.mark.parametrize("src_wh,dst_wh", [((640, 480), (533, 400)), ((533, 400), (444, 333)), ((752, 480), (627, 400)),
                                           ((179, 134), (149, 112)), ((640, 480), (320, 240)), ((80, 60), (40, 30)),
                                           ((101, 77), (84, 64)), ((50, 40), (49, 39)), ((31, 9), (26, 8))])
def test_resize_matches_cv2(src_wh, dst_wh):
    rng = np.random.default_rng(src_wh[0] * 7 + dst_wh[0])
    img = rng.integers(0, 256, (src_wh[1], src_wh[0]), dtype=np.uint8)
    ref = cv2.resize(img, dst_wh, interpolation=cv2.INTER_LINEAR)
    assert np.array_equal(orc.resize_linear(img, *dst_wh), ref)


def test_resize_reads_roi_only():
    """resize of an ROI view must not see the pixels around it (the pyramid levels are ROIs of padded buffers)."""
    rng = np.random.default_rng(5)
    big = rng.integers(0, 256, (140, 180), dtype=np.uint8)
    roi = big[19:-19, 19:-19]
    assert np.array_equal(orc.resize_linear(roi, 118, 85), cv2.resize(np.ascontiguousarray(roi), (118, 85), interpolation=cv2.INTER_LINEAR))


# ------------------------------------------------------------------ copyMakeBorder
@pytest.mark.parametrize("name,img", list(_images()))
def test_border_matches_cv2(name, img):
    ref = cv2.copyMakeBorder(img, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
    assert np.array_equal(orc.border_reflect101(img, 19), ref)


# ------------------------------------------------------------------ cv::GaussianBlur 7x7 sigma 2
@pytest.mark.parametrize("name,img", list(_images()))
def test_blur_matches_cv2(name, img):
    ref = cv2.GaussianBlur(img.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(orc.gaussian_blur(img), ref)


@pytest.mark.parametrize("wh", [(7, 7), (8, 3), (3, 9), (1, 1), (2, 5), (640, 4)])
def test_blur_small_images(wh):
    rng = np.random.default_rng(wh[0] * 31 + wh[1])
    img = rng.integers(0, 256, (wh[1], wh[0]), dtype=np.uint8)
    ref = cv2.GaussianBlur(img.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
    assert np.array_equal(orc.gaussian_blur(img), ref)


# ------------------------------------------------------------------ cv::FAST 9/16 + NMS
def _cv_fast(img, th, nonmax=True):
    det = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=nonmax, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    return [(k.pt[0], k.pt[1], k.response) for k in det.detect(np.ascontiguousarray(img))]


@pytest.mark.parametrize("name,img", list(_images()))
@pytest.mark.parametrize("th", [7, 20, 60])
def test_fast_matches_cv2(name, img, th):
    got = orc.fast(img, th)
    exp = _cv_fast(img, th)
    assert [(float(a), float(b), float(c)) for a, b, c in zip(got["x"], got["y"], got["response"])] == exp
    assert (got["size"] == 7).all() and (got["angle"] == -1).all() and (got["octave"] == 0).all() and (got["class_id"] == -1).all()


def test_fast_cell_roi_equals_view():
    """The reference calls FAST on overlapping cell ROIs (src/ORBextractor.cc:532-536): a strided view must give
    the result of a contiguous copy, in cell-local coordinates."""
    img = synth.smooth_noise(9, 320, 240)
    roi = img[16:16 + 80, 16:16 + 127]
    got = orc.fast(roi, 20)
    exp = _cv_fast(roi.copy(), 20)
    assert [(float(a), float(b), float(c)) for a, b, c in zip(got["x"], got["y"], got["response"])] == exp


@pytest.mark.parametrize("wh", [(6, 20), (20, 6), (7, 7), (8, 7), (2, 7)])
def test_fast_tiny(wh):
    rng = np.random.default_rng(3)
    img = rng.integers(0, 256, (wh[1], wh[0]), dtype=np.uint8)
    got = orc.fast(img, 5)
    exp = _cv_fast(img, 5)
    assert [(float(a), float(b), float(c)) for a, b, c in zip(got["x"], got["y"], got["response"])] == exp


def test_fast_score_map_consistent_with_keypoints():
    img = synth.smooth_noise(6, 200, 150)
    sm = orc.fast_score_map(img, 20)
    k = orc.fast(img, 20, nonmax=False)
    assert (sm[:3] == 0).all() and (sm[-3:] == 0).all() and (sm[:, :3] == 0).all() and (sm[:, -3:] == 0).all()
    assert int((sm > 0).sum()) == len(k)
    kn = orc.fast(img, 20, nonmax=True)
    assert np.array_equal(sm[kn["y"].astype(int), kn["x"].astype(int)].astype(np.float32), kn["response"])


# ------------------------------------------------------------------ cv::fastAtan2
def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(1)
    ys = np.concatenate([rng.integers(-2_700_000, 2_700_000, 20000), [0, 0, 1, -1, 5, 0, -7]]).astype(np.float32)
    xs = np.concatenate([rng.integers(-2_700_000, 2_700_000, 20000), [0, 1, 0, 0, 5, -3, -7]]).astype(np.float32)
    for y, x in zip(ys, xs):
        assert np.float32(orc.fast_atan2(y, x)) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)
    assert orc.fast_atan2(0.0, 0.0) == 0.0


# ------------------------------------------------------------------ KeyPointsFilter::retainBest
@pytest.mark.parametrize("n", [1, 2, 3, 17, 50, 200, 377, 1000])
@pytest.mark.parametrize("seed", [3, 8])
def test_retain_best_matches_cv2_orb(n, seed):
    """cv2.ORB with one level and FAST_SCORE is FAST -> runByImageBorder -> retainBest(n): its output order is the
    order std::nth_element + std::partition leave behind, which is what the reference's per-cell / per-level
    trims depend on."""
    img = synth.smooth_noise(seed, 320, 240)
    orb = cv2.ORB_create(nfeatures=n, nlevels=1, scoreType=cv2.ORB_FAST_SCORE, edgeThreshold=31, patchSize=31,
                         fastThreshold=20)
    exp = [(k.pt[0], k.pt[1], k.response) for k in orb.detect(img)]
    f = orc.fast(img, 20)
    e = 31
    f = f[(f["x"] >= e) & (f["x"] < img.shape[1] - e) & (f["y"] >= e) & (f["y"] < img.shape[0] - e)]
    sel = f[orc.retain_best_order(f["response"], n)]
    assert [(float(a), float(b), float(c)) for a, b, c in zip(sel["x"], sel["y"], sel["response"])] == exp


def test_retain_best_edge_cases():
    r = np.array([5, 5, 5, 5], np.float32)
    assert list(orc.retain_best_order(r, 10)) == [0, 1, 2, 3]      # n >= size: untouched
    assert list(orc.retain_best_order(r, 4)) == [0, 1, 2, 3]
    assert len(orc.retain_best_order(r, 0)) == 0                    # n == 0 clears
    assert len(orc.retain_best_order(r, 2)) == 4                    # all tied with the cut: all are kept (before resize)
    assert len(orc.retain_best_order(np.zeros(0, np.float32), 3)) == 0
    kp = np.zeros(6, orc.KP_DTYPE)
    kp["response"] = [1, 9, 3, 9, 2, 7]
    kp["x"] = np.arange(6)
    m = orc.lib().orc_retain_best(kp.ctypes.data_as(ctypes.c_void_p), 6, 3)
    assert m == 3 and sorted(kp["response"][:3].tolist()) == [7, 9, 9]


# ------------------------------------------------------------------ glibc sincosf restated
def test_sincosf_restated_sampled():
    libm = ctypes.CDLL("libm.so.6")
    libm.sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
    rng = np.random.default_rng(2)
    xs = np.concatenate([rng.uniform(0, 2 * np.pi, 5000), [0.0, 1e-6, 2 ** -13, np.pi / 4, np.pi / 2, np.pi, 2 * np.pi]]).astype(np.float32)
    for x in xs:
        s, c = ctypes.c_float(), ctypes.c_float()
        libm.sincosf(float(x), ctypes.byref(s), ctypes.byref(c))
        assert orc.sincosf_restated(x) == (s.value, c.value), x


def test_sincosf_restated_exhaustive():
    """Every float in [0, 2*pi*1.0001] (about 1.09e9 values): the restated algorithm equals this libm bit for bit."""
    assert orc.sincosf_mismatches(0.0, 2 * np.pi * 1.0001, nthreads=os.cpu_count() or 4) == 0


# ------------------------------------------------------------------ descriptor index math
_VERBATIM = r"""
#include <cmath>
extern "C" int cvRound_(float v) { return (int)lrintf(v); }
// the expression of src/ORBextractor.cc:115-117, compiled with the reference's flags (CMakeLists.txt:39-40)
extern "C" void rot(int px, int py, float a, float b, int* r, int* c) {
  *r = cvRound_(px * b + py * a);
  *c = cvRound_(px * a - py * b);
}
"""


def test_pattern_rotation_matches_reference_flags():
    """g++ -O3 -march=native contracts x*b + y*a into fma(x, b, y*a) on this FMA host; the oracle freezes exactly
    that form.  Compile the verbatim expression with the reference flags and compare on a sweep of angles."""
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "v.cc"), os.path.join(d, "v.so")
        open(src, "w").write(_VERBATIM)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++11", "-shared", "-fPIC", "-o", so, src])
        flags = open("/proc/cpuinfo").read()
        if " fma" not in flags:
            pytest.skip("host CPU has no FMA: the reference build would not contract here")
        lib = ctypes.CDLL(so)
        lib.rot.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_float, ctypes.c_float, ctypes.POINTER(ctypes.c_int), ctypes.POINTER(ctypes.c_int)]
        rng = np.random.default_rng(4)
        angles = np.concatenate([rng.uniform(0, 360, 300), [0, 45, 90, 180, 270, 359.99]]).astype(np.float32)
        pts = [(-13, 12), (13, -13), (7, 0), (0, -9), (1, 1), (-5, -12), (12, 13), (-13, -13), (4, 7), (-2, 9)]
        for ang in angles:
            s, c = orc.sincosf_restated(np.float32(ang) * np.float32(np.pi / 180.0))
            for px, py in pts:
                r, cc = ctypes.c_int(), ctypes.c_int()
                lib.rot(px, py, c, s, ctypes.byref(r), ctypes.byref(cc))
                assert orc.pattern_rotate(px, py, c, s) == (r.value, cc.value)


# ------------------------------------------------------------------ DescriptorDistance
def test_descriptor_distance_against_bit_count():
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_96x80.npz"))
    A, B, dist = g["A"], g["B"], g["dist"]
    assert np.array_equal(orc.hamming_matrix(A, B), dist)
    assert orc.descriptor_distance(A[7], B[5]) == 0
    assert orc.descriptor_distance(np.zeros(32, np.uint8), np.full(32, 255, np.uint8)) == 256


def test_match_best2_rule():
    """src/ORBmatcher.cc:1239-1262: strict <, ascending j -> first minimal index; second = second smallest value."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "hamming_96x80.npz"))
    A, B, dist = g["A"], g["B"], g["dist"].astype(np.int32)
    m = orc.match_best2(A, B, ratio=0.75, th_low=50)
    assert np.array_equal(m["best_idx"], dist.argmin(axis=1))
    assert np.array_equal(m["best_dist"], dist.min(axis=1))
    assert np.array_equal(m["second_dist"], np.sort(dist, axis=1)[:, 1])
    assert m["best_idx"][7] == 5 and m["best_dist"][7] == 0 and m["second_dist"][7] == 0 and m["accepted"][7] == 0
    acc = (m["best_dist"] < 50) & (m["best_dist"].astype(np.float32) < np.float32(0.75) * m["second_dist"].astype(np.float32))
    assert np.array_equal(m["accepted"].astype(bool), acc)
    empty = orc.match_best2(A, np.zeros((0, 32), np.uint8))
    assert (empty["best_idx"] == -1).all() and (empty["best_dist"] == 256).all() and (empty["second_dist"] == 256).all()


# ------------------------------------------------------------------ MapPoint::ComputeDistinctiveDescriptors
def _distinctive_numpy(d):
    """Independent restatement of /root/reference/src/MapPoint.cc:252-275 with numpy bit counting."""
    n = len(d)
    if n == 0:
        return -1, 2 ** 31 - 1
    dist = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
    med = np.sort(dist, axis=1)[:, int(0.5 * (n - 1))]
    return int(np.argmin(med)), int(med.min())


def test_distinctive_descriptor_rule():
    rng = np.random.default_rng(12)
    sizes = [1, 2, 3, 4, 5, 8, 17, 40, 0, 100, 2]
    sets = []
    for n in sizes:
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if n >= 5:  # a cluster of near-duplicates plus outliers, and exact duplicates (first index must win)
            d[1:n // 2] = d[0] ^ (rng.integers(0, 256, (n // 2 - 1, 32), dtype=np.uint8) & 0x11)
            d[n - 1] = d[1]
        sets.append(d)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    best, med = orc.distinctive_many(np.concatenate(sets), offsets, nthreads=2)
    for s, d in enumerate(sets):
        bi, bm = _distinctive_numpy(d)
        assert (int(best[s]), int(med[s])) == (bi, bm), "set %d (n=%d)" % (s, len(d))


# ------------------------------------------------------------------ Frame::AssignFeaturesToGrid / ComputeStereoFromRGBD
def test_assign_grid_rule():
    """src/Frame.cc:179-192, 323-332 against a direct Python restatement (round half away from zero, x-major cells)."""
    import math
    rng = np.random.default_rng(21)
    k = np.zeros(700, orc.KP_DTYPE)
    k["x"] = rng.uniform(-20, 660, 700).astype(np.float32)
    k["y"] = rng.uniform(-20, 500, 700).astype(np.float32)
    k["x"][:8] = [0, 5, 15, 634.9, 635, 639.99, 4.99999, 645]   # cell borders: 640 / 64 = 10 px per cell
    k["y"][:8] = [0, 5, 15, 474.9, 475, 479.99, 4.99999, 100]
    inv_w, inv_h = np.float32(64) / np.float32(640), np.float32(48) / np.float32(480)
    cs, idx = orc.assign_grid(k, 0.0, 0.0, float(inv_w), float(inv_h))
    grid = {}
    for i in range(len(k)):
        fx, fy = np.float32(k["x"][i] - np.float32(0)) * inv_w, np.float32(k["y"][i] - np.float32(0)) * inv_h
        px = int(math.copysign(math.floor(abs(float(fx)) + 0.5), float(fx)))
        py = int(math.copysign(math.floor(abs(float(fy)) + 0.5), float(fy)))
        if 0 <= px < 64 and 0 <= py < 48:
            grid.setdefault(px * 48 + py, []).append(i)
    assert cs[-1] == sum(len(v) for v in grid.values()) == len(idx)
    for c in range(64 * 48):
        assert idx[cs[c]:cs[c + 1]].tolist() == grid.get(c, [])


def test_stereo_from_rgbd_rule():
    rng = np.random.default_rng(22)
    depth = rng.uniform(-1, 8, (48, 64)).astype(np.float32)
    depth[depth < 0.5] = 0
    k = np.zeros(50, orc.KP_DTYPE)
    k["x"] = rng.uniform(0, 63.9, 50).astype(np.float32)
    k["y"] = rng.uniform(0, 47.9, 50).astype(np.float32)
    ku = k.copy()
    ku["x"] += np.float32(0.25)
    ur, z = orc.stereo_from_rgbd(k, ku, depth, 40.0)
    for i in range(50):
        d = depth[int(k["y"][i]), int(k["x"][i])]
        if d > 0:
            assert z[i] == d and ur[i] == np.float32(ku["x"][i] - np.float32(40.0) / d)
        else:
            assert z[i] == -1 and ur[i] == -1


# ------------------------------------------------------------------ Frame::UndistortKeyPoints / ComputeImageBounds vs real cv2
CAMERAS = {  # (fx, fy, cx, cy), distortion: the TUM1 / TUM2 / EuRoC-like models SD-SLAM is run with (README.md:41-105)
    "tum1": ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    "tum2": ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
    "euroc": ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    "none": ((500.0, 500.0, 320.0, 240.0), (0.0, 0.0, 0.0, 0.0)),
}


@pytest.mark.parametrize("cam", sorted(CAMERAS))
def test_undistort_keypoints_matches_cv2(cam):
    K4, dist = CAMERAS[cam]
    K4, dist = np.array(K4, np.float32), np.array(dist, np.float32)
    rng = np.random.default_rng(5)
    k = np.zeros(4000, orc.KP_DTYPE)
    k["x"] = rng.uniform(0, 752, 4000).astype(np.float32)
    k["y"] = rng.uniform(0, 480, 4000).astype(np.float32)
    k["x"][:4], k["y"][:4] = [0, 752, 0, 752], [0, 0, 480, 480]
    k["octave"] = rng.integers(0, 8, 4000)
    out = orc.undistort_keypoints(k, K4, dist)
    Km = np.array([[K4[0], 0, K4[2]], [0, K4[1], K4[3]], [0, 0, 1]], np.float32)
    if dist[0] != 0:
        pts = np.stack([k["x"], k["y"]], 1).reshape(-1, 1, 2)
        ref = cv2.undistortPoints(pts, Km, dist, None, Km).reshape(-1, 2)
    else:
        ref = np.stack([k["x"], k["y"]], 1)  # src/Frame.cc:336-339: mvKeysUn = mvKeys
    assert np.array_equal(out["x"], ref[:, 0]) and np.array_equal(out["y"], ref[:, 1])
    for name in ("size", "angle", "response", "octave", "class_id"):
        assert np.array_equal(out[name], k[name])
    b = orc.image_bounds(752, 480, K4, dist)
    if dist[0] != 0:
        c = cv2.undistortPoints(np.array([[[0, 0]], [[752, 0]], [[0, 480]], [[752, 480]]], np.float32), Km, dist, None, Km).reshape(4, 2)
        exp = [min(c[0, 0], c[2, 0]), max(c[1, 0], c[3, 0]), min(c[0, 1], c[1, 1]), max(c[2, 1], c[3, 1])]
    else:
        exp = [0, 752, 0, 480]
    assert b.tolist() == [float(np.float32(v)) for v in exp]


def test_oracle_under_asan_ubsan(tmp_path):
    """SURVEY.md section 4, row "Sanitizers": the oracle compiled with -fsanitize=address,undefined runs both extractor modes
    (incl. the reference defaults with their negative cell height, a cornerless and a tiny image), the primitives, every matcher
    (incl. empty sides) and the Frame post-processing without a report (tests/helpers/oracle_sanitize_driver.cc)."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "orc_san")
    cmd = ["g++", "-O1", "-g", "-std=c++17", "-ffp-contract=off", "-pthread", "-fsanitize=address,undefined", "-fno-sanitize-recover=all",
           "-o", exe, os.path.join(root, "tests", "helpers", "oracle_sanitize_driver.cc"), os.path.join(root, "oracle", "sdorb_oracle.cc"), "-lm"]
    try:
        subprocess.check_call(cmd)
    except (OSError, subprocess.CalledProcessError):
        pytest.skip("no sanitizer runtime for g++ on this host")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "oracle sanitize run ok" in r.stdout, r.stderr[-4000:]
