"""Pins the oracle's guided matchers (oracle/sdorb_oracle.cc: orc_features_in_area, orc_three_maxima,
orc_search_for_initialization, orc_search_by_projection) against a second, independent restatement of
/root/reference/src/ORBmatcher.cc:256-357, 946-1075, 1423-1454 and src/Frame.cc:271-321 (tests/search_cases.py).
(Round 2: the same oracle functions are also compared with the reference's own compiled text, tests/test_ref_matchers.py.)
The reference has no tests or fixtures for these functions and its own build cannot run here (Eigen / OpenCV C++ headers are
absent), so two independently written restatements agreeing is the pin; a committed fixture (tests/golden/matcher/search_pairs.npz,
made by tests/golden/make_search_golden.py from the Python restatement) freezes it.  CPU only."""
import os

import numpy as np
import pytest

import search_cases as sc
from oracle import binding as orc

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _orc_grid(k, gp):
    cs, idx = orc.assign_grid(k, *gp)
    return (cs, idx) + tuple(gp)


@pytest.mark.parametrize("seed", range(4))
def test_features_in_area_equals_restatement(seed):
    k1, d1, k2, d2 = sc.frame_pair(seed, 300, 500)
    gp = sc.grid_params(640, 480, -3.5 if seed & 1 else 0.0, 2.25 if seed & 1 else 0.0)
    g = _orc_grid(k2, gp)
    pg = sc.py_grid(k2, gp)
    rng = np.random.default_rng(seed)
    for q in range(300):
        x, y = rng.uniform(-80, 720), rng.uniform(-80, 560)
        r = float(rng.choice([0.5, 3, 15, 40, 100, 900]))
        lo, hi = [(-1, -1), (0, -1), (2, -1), (0, 0), (1, 3), (0, 5), (3, 2)][q % 7]
        assert orc.features_in_area(k2, g, x, y, r, lo, hi).tolist() == sc.py_features_in_area(k2, pg, gp, x, y, r, lo, hi)


def test_three_maxima_equals_restatement():
    rng = np.random.default_rng(5)
    cases = [rng.integers(0, 40, 30) for _ in range(200)] + [np.zeros(30, int), np.full(30, 7), np.arange(30), np.arange(30)[::-1]]
    cases += [np.array([100] + [9] * 29), np.array([100, 10] + [9] * 28), np.array([100, 10, 10] + [0] * 27), np.array([0] * 29 + [1])]
    for s in cases:
        assert orc.three_maxima(s) == sc.py_three_maxima([int(v) for v in s])


@pytest.mark.parametrize("seed,n1,n2,window,ratio,orient,dup", [
    (0, 600, 640, 100, 0.9, True, 0.0), (1, 400, 380, 30, 0.9, True, 0.3), (2, 500, 500, 100, 0.6, False, 0.5),
    (3, 64, 700, 15, 1.0, True, 0.0), (4, 300, 1, 100, 0.9, True, 0.0), (5, 0, 50, 100, 0.9, True, 0.0),
    (6, 50, 0, 100, 0.9, True, 0.0), (7, 700, 700, 400, 0.9, True, 0.8)])
def test_search_for_initialization_equals_restatement(seed, n1, n2, window, ratio, orient, dup):
    k1, d1, k2, d2 = sc.frame_pair(seed, n1, n2, dup=dup, flips=20)
    gp = sc.grid_params()
    prev = np.stack([k1["x"], k1["y"]], 1) if n1 else np.zeros((0, 2), np.float32)  # Initializer: vbPrevMatched = F1 keypoints
    n, m12, pm = orc.search_for_initialization(k1, d1, k2, d2, _orc_grid(k2, gp), prev, window, ratio, orient)
    pn, pm12, ppm = sc.py_search_for_initialization(k1, d1, k2, d2, gp, prev, window, ratio, orient)
    assert n == pn and np.array_equal(m12, pm12) and pm.tobytes() == ppm.tobytes()
    assert n == int((m12 >= 0).sum())
    if seed == 0:
        assert n > 100  # the case exercises real matches
    if n:  # a second call continues from the updated vbPrevMatched, as Tracking::MonocularInitialization does
        n2_, m12b, _ = orc.search_for_initialization(k1, d1, k2, d2, _orc_grid(k2, gp), pm, window, ratio, orient)
        pn2, pm12b, _ = sc.py_search_for_initialization(k1, d1, k2, d2, gp, ppm, window, ratio, orient)
        assert n2_ == pn2 and np.array_equal(m12b, pm12b)


@pytest.mark.parametrize("seed,nl,nc,th,mode,orient,stereo", [
    (0, 600, 640, 15.0, 0, True, False), (1, 500, 450, 7.0, 0, True, True), (2, 400, 500, 15.0, 1, True, True),
    (3, 400, 500, 15.0, 2, False, True), (4, 300, 0, 15.0, 0, True, False), (5, 0, 300, 15.0, 0, True, False),
    (6, 700, 700, 30.0, 0, True, True)])
def test_search_by_projection_equals_restatement(seed, nl, nc, th, mode, orient, stereo):
    kl, dl, kc, dc = sc.frame_pair(seed + 20, nl, nc, jitter=4.0, dup=0.2 if seed == 6 else 0.0, level0=0.3)
    proj, flags, dmp = sc.projection_inputs(seed, kl, dl)
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(nc) < 0.6, kc["x"] - rng.uniform(0, 30, nc), -1).astype(np.float32) if stereo else np.full(nc, -1, np.float32)
    occ = (rng.random(nc) < 0.1).astype(np.uint8)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp = sc.grid_params()
    bounds = (0.0, 640.0, 0.0, 480.0)
    klu = kl.copy()
    klu["angle"] = kl["angle"]
    n, asg = orc.search_by_projection(kl, klu, proj, flags, dmp, kc, dc, ur, occ, _orc_grid(kc, gp), sf, bounds, th, 40.0, mode, orient)
    pn, pasg = sc.py_search_by_projection(kl, klu, proj, flags, dmp, kc, dc, ur, occ, gp, sf, bounds, th, 40.0, mode, orient)
    assert n == pn and np.array_equal(asg, pasg)
    if seed == 0:
        assert n > 100


def test_search_golden_fixture():
    """The committed vectors (inputs + results of the Python restatement) against the oracle."""
    z = np.load(os.path.join(GOLD, "matcher", "search_pairs.npz"))
    gp = tuple(float(v) for v in z["gp"])
    k1, k2 = z["k1"].view(orc.KP_DTYPE).reshape(-1), z["k2"].view(orc.KP_DTYPE).reshape(-1)
    n, m12, pm = orc.search_for_initialization(k1, z["d1"], k2, z["d2"], _orc_grid(k2, gp), z["prev"], 100, 0.9, True)
    assert n == int(z["init_n"]) and np.array_equal(m12, z["init_m12"]) and pm.tobytes() == z["init_prev"].tobytes()
    n, asg = orc.search_by_projection(k1, k1, z["proj"], z["flags"], z["d1"], k2, z["d2"], z["ur"], z["occ"], _orc_grid(k2, gp),
                                      z["sf"], z["bounds"], 15.0, 40.0, 0, True)
    assert n == int(z["proj_n"]) and np.array_equal(asg, z["proj_assigned"])


# ------------------------------------------------------------------ SearchForTriangulation (row f4, second half)
_VERBATIM_EPIPOLAR = r"""
// the arithmetic of ORBmatcher::CheckDistEpipolarLine (src/ORBmatcher.cc:128-144), F12(i, j) as F[3 * i + j]
extern "C" int check(float x1, float y1, float x2, float y2, const double* F, float sigma2) {
  const float a = x1*F[0*3+0]+y1*F[1*3+0]+F[2*3+0];
  const float b = x1*F[0*3+1]+y1*F[1*3+1]+F[2*3+1];
  const float c = x1*F[0*3+2]+y1*F[1*3+2]+F[2*3+2];
  const float num = a*x2+b*y2+c;
  const float den = a*a+b*b;
  if (den == 0)
    return 0;
  const float dsqr = num*num/den;
  return dsqr<3.84*sigma2;
}
"""


def test_epipolar_gate_matches_reference_flags():
    """g++ -O3 -march=native (the reference's flags, CMakeLists.txt:40) contracts p*q + r*s into fma(p, q, r*s); the oracle
    freezes that.  Compile the verbatim expression with those flags and compare on points near and far from the line."""
    import ctypes
    import subprocess
    import tempfile
    if " fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA: the reference build would not contract here")
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "v.cc"), os.path.join(d, "v.so")
        open(src, "w").write(_VERBATIM_EPIPOLAR)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++11", "-shared", "-fPIC", "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.check.argtypes = [ctypes.c_float] * 4 + [ctypes.c_void_p, ctypes.c_float]
        case = sc.triangulation_case(1, 400, 400)
        k1, k2, F12, sigma2 = case[0], case[4], case[8], case[12]
        F = np.ascontiguousarray(F12, np.float64).reshape(9)
        rng = np.random.default_rng(8)
        n_true = 0
        for _ in range(20000):
            i1, i2 = int(rng.integers(0, len(k1))), int(rng.integers(0, len(k2)))
            s2 = float(sigma2[int(k2["octave"][i2])])
            args = (float(k1["x"][i1]), float(k1["y"][i1]), float(k2["x"][i2]), float(k2["y"][i2]))
            ref = bool(lib.check(*args, F.ctypes.data, s2))
            assert orc.check_dist_epipolar_line(*args, F12, s2) == ref
            n_true += ref
        assert 50 < n_true < 19000
        Z = np.zeros(9)
        assert not orc.check_dist_epipolar_line(1.0, 2.0, 3.0, 4.0, Z, 1.0) and not lib.check(1.0, 2.0, 3.0, 4.0, Z.ctypes.data, 1.0)


@pytest.mark.parametrize("seed,n1,n2,orient", [(0, 140, 150, True), (1, 120, 90, False), (2, 60, 200, True), (3, 0, 40, True),
                                               (4, 40, 0, True), (5, 150, 150, True)])
def test_search_for_triangulation_equals_restatement(seed, n1, n2, orient):
    case = sc.triangulation_case(seed, n1, n2, dup=0.3 if seed == 5 else 0.1)
    n, m12 = orc.search_for_triangulation(*case, orient)
    pn, pm12 = sc.py_search_for_triangulation(*case, orient)
    assert n == pn and np.array_equal(m12, pm12)
    assert n <= int((m12 >= 0).sum())  # every removal of the orientation check is counted, also none twice here
    if seed == 0:
        assert n > 10


def _load_triangulation_fixture():
    z = np.load(os.path.join(GOLD, "matcher", "triangulation_pair.npz"))
    k1, k2 = z["k1"].view(orc.KP_DTYPE).reshape(-1), z["k2"].view(orc.KP_DTYPE).reshape(-1)
    case = (k1, z["d1"], z["mp1"], z["ur1"], k2, z["d2"], z["mp2"], z["ur2"], z["F12"], float(z["epipole"][0]), float(z["epipole"][1]),
            z["sf"], z["sigma2"])
    return case, int(z["n"]), z["m12"]


def test_triangulation_golden_fixture():
    """The committed vectors made by the Python restatement (tests/golden/make_search_golden.py) against the oracle."""
    case, n, m12 = _load_triangulation_fixture()
    on, om12 = orc.search_for_triangulation(*case, True)
    assert on == n and np.array_equal(om12, m12) and n > 10


# ------------------------------------------------------------------ SearchByProjection(Frame, map points): Tracking::SearchLocalPoints
@pytest.mark.parametrize("seed,nf,nmp,th,ratio,stereo", [(0, 600, 500, 1.0, 0.8, False), (1, 500, 700, 3.0, 0.8, True),
                                                         (2, 300, 200, 5.0, 0.6, True), (3, 0, 50, 1.0, 0.8, False),
                                                         (4, 200, 0, 1.0, 0.8, False), (5, 700, 900, 1.0, 1.0, True)])
def test_search_map_points_equals_restatement(seed, nf, nmp, th, ratio, stereo):
    _, _, kf, df = sc.frame_pair(seed + 60, 10, nf, dup=0.3 if seed == 5 else 0.0, level0=0.3)
    proj, vc, lvl, fl, dmp = sc.map_point_inputs(seed, kf, df, nmp)
    rng = np.random.default_rng(seed)
    ur = (np.where(rng.random(nf) < 0.6, kf["x"] - rng.uniform(0, 30, nf), -1) if stereo else np.full(nf, -1)).astype(np.float32)
    occ = (rng.random(nf) < 0.1).astype(np.uint8)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp = sc.grid_params()
    n, asg = orc.search_map_points(proj, vc, lvl, fl, dmp, kf, df, ur, occ, _orc_grid(kf, gp), sf, th, ratio)
    pn, pasg = sc.py_search_map_points(proj, vc, lvl, fl, dmp, kf, df, ur, occ, gp, sf, th, ratio)
    assert n == pn and np.array_equal(asg, pasg)
    if seed == 0:
        assert n > 100


# ------------------------------------------------------------------ SearchByPoints (loop detection)
@pytest.mark.parametrize("seed,n1,n2,ratio,orient,dup", [(0, 300, 320, 0.75, True, 0.0), (1, 200, 150, 0.9, False, 0.4),
                                                         (2, 0, 40, 0.75, True, 0.0), (3, 40, 0, 0.75, True, 0.0), (4, 250, 250, 0.75, True, 0.6)])
def test_search_by_points_equals_restatement(seed, n1, n2, ratio, orient, dup):
    k1, d1, k2, d2 = sc.frame_pair(seed + 80, n1, n2, dup=dup, flips=25)
    rng = np.random.default_rng(seed)
    v1, v2 = (rng.random(n1) < 0.8).astype(np.uint8), (rng.random(n2) < 0.8).astype(np.uint8)
    n, m12 = orc.search_by_points(k1, d1, v1, k2, d2, v2, ratio, orient)
    pn, pm12 = sc.py_search_by_points(k1, d1, v1, k2, d2, v2, ratio, orient)
    assert n == pn and np.array_equal(m12, pm12)
    if seed == 0:
        assert n > 50


# ------------------------------------------------------------------ SearchByProjection(Frame, KeyFrame, ...) through the Frame / Frame form
@pytest.mark.parametrize("seed,nk,nc,th,orb_dist,orient", [(0, 500, 520, 10.0, 100, True), (1, 400, 300, 3.0, 64, True),
                                                           (2, 300, 400, 10.0, 64, False), (3, 0, 100, 10.0, 100, True)])
def test_keyframe_projection_overload_maps_onto_frame_frame_form(seed, nk, nc, th, orb_dist, orient):
    """The relocalisation overload (src/ORBmatcher.cc:1306-1421) restated directly in Python equals the oracle's Frame / Frame
    function driven with the mapping documented at sdorb_projection_search::orb_dist."""
    a = sc.kf_projection_args(seed, nk, nc)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp, bounds = sc.grid_params(), (0.0, 640.0, 0.0, 480.0)
    pn, pasg = sc.py_search_by_projection_kf(a["kk"], a["proj"], a["valid"], a["pred"], a["dmp"], a["kc"], a["dc"], a["has_mp"], gp, sf,
                                             bounds, th, orb_dist, orient)
    n, asg = orc.search_by_projection(a["k_level"], a["kk"], a["proj"], a["flags"], a["dmp"], a["kc"], a["dc"], a["ur"], a["has_mp"],
                                      _orc_grid(a["kc"], gp), sf, bounds, th, 0.0, 0, orient, orb_dist=orb_dist)
    assert n == pn and np.array_equal(asg, pasg)
    if seed == 0:
        assert n > 100


# ------------------------------------------------------------------ Fuse: the keypoint search (src/ORBmatcher.cc:535-586)
_VERBATIM_FUSE_E2 = r"""
extern "C" int gate(float u, float v, float ur, float kpx, float kpy, float kpr, float inv, int stereo) {
  if (stereo) {
    const float ex = u-kpx;
    const float ey = v-kpy;
    const float er = ur-kpr;
    const float e2 = ex*ex+ey*ey+er*er;
    if (e2*inv>7.8) return 0;
  } else {
    const float ex = u-kpx;
    const float ey = v-kpy;
    const float e2 = ex*ex+ey*ey;
    if (e2*inv>5.99) return 0;
  }
  return 1;
}
"""


def test_fuse_e2_matches_reference_flags():
    """The chi-square gate of Fuse (src/ORBmatcher.cc:560-580) compiled verbatim with the reference's flags (-O3 -march=native,
    CMakeLists.txt:40) against the oracle's frozen contraction, on single-keypoint frames with errors around both limits."""
    import ctypes
    import subprocess
    import tempfile
    if " fma" not in open("/proc/cpuinfo").read():
        pytest.skip("host CPU has no FMA: the reference build would not contract here")
    with tempfile.TemporaryDirectory() as d:
        src, so = os.path.join(d, "v.cc"), os.path.join(d, "v.so")
        open(src, "w").write(_VERBATIM_FUSE_E2)
        subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++11", "-shared", "-fPIC", "-o", so, src])
        lib = ctypes.CDLL(so)
        lib.gate.argtypes = [ctypes.c_float] * 7 + [ctypes.c_int]
        rng = np.random.default_rng(21)
        sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
        inv = (np.float32(1) / (sf * sf)).astype(np.float32)
        gp = sc.grid_params()
        kf = np.zeros(1, orc.KP_DTYPE)
        desc = np.zeros((1, 32), np.uint8)
        n_pass = 0
        for t in range(4000):
            lvl = int(rng.integers(0, 8))
            stereo = bool(t & 1)
            lim = (7.8 if stereo else 5.99) / float(inv[lvl])
            rad = np.sqrt(lim * rng.uniform(0.9, 1.1))  # |e| around the limit
            dirn = rng.normal(size=3 if stereo else 2)
            e = dirn / np.linalg.norm(dirn) * rad
            kf["x"], kf["y"], kf["octave"] = np.float32(rng.uniform(100, 500)), np.float32(rng.uniform(100, 400)), lvl
            kpr = np.float32(kf["x"][0] - 12.5) if stereo else np.float32(-1)
            u, v = np.float32(kf["x"][0] + e[0]), np.float32(kf["y"][0] + e[1])
            ur = np.float32(kpr + e[2]) if stereo else np.float32(0)
            ref = lib.gate(float(u), float(v), float(ur), float(kf["x"][0]), float(kf["y"][0]), float(kpr), float(inv[lvl]), int(stereo))
            bi, bd = orc.fuse_search(np.array([[u, v, ur]], np.float32), [lvl], [1], desc, kf, desc, [kpr], _orc_grid(kf, gp), sf, inv, 30.0)
            assert (bd[0] == 0) == bool(ref) and (bi[0] == 0) == bool(ref)
            n_pass += ref
        assert 1000 < n_pass < 3000


@pytest.mark.parametrize("seed,nf,nmp,th,stereo,dup", [(0, 600, 500, 3.0, True, 0.0), (1, 500, 700, 2.5, False, 0.0), (2, 0, 50, 3.0, True, 0.0),
                                                       (3, 200, 0, 3.0, True, 0.0), (4, 700, 900, 4.0, True, 0.3), (5, 1, 1, 3.0, False, 0.0),
                                                       (6, 600, 600, 4.0, True, 0.0)])
def test_fuse_search_equals_restatement(seed, nf, nmp, th, stereo, dup):
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=dup, level0=0.3)
    proj, lvl, fl, ur = sc.fuse_inputs(seed, kf, nmp, stereo=stereo)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    inv = (np.float32(1) / (sf * sf)).astype(np.float32)
    gp = sc.grid_params() if seed != 6 else sc.grid_params(667.75, 497.0, -12.25, -7.5)  # image bounds of a distorted camera
    bi, bd = orc.fuse_search(proj, lvl, fl, dmp, kf, df, ur, _orc_grid(kf, gp), sf, inv, th)
    pbi, pbd = sc.py_fuse_search(proj, lvl, fl, dmp, kf, df, ur, gp, sf, inv, th)
    assert np.array_equal(bi, pbi) and np.array_equal(bd, pbd)
    if seed == 0:
        assert (bi >= 0).sum() > 100 and ((bd > 50) & (bd < 256)).sum() > 20


@pytest.mark.parametrize("seed,nf,nmp,th,th_dist", [(0, 600, 500, 3.0, 50), (1, 500, 700, 4.0, 100), (2, 0, 50, 3.0, 50), (4, 700, 900, 4.0, 50)])
def test_fuse_search_sim3_form_equals_restatement(seed, nf, nmp, th, th_dist):
    """Without the reprojection gate: Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:682-708; TH_LOW) and one
    direction of SearchBySim3 (TH_HIGH)."""
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=0.3 if seed == 4 else 0.0, level0=0.3)
    proj, lvl, fl, _ = sc.fuse_inputs(seed, kf, nmp, stereo=False)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp = sc.grid_params()
    bi, bd = orc.fuse_search(proj, lvl, fl, dmp, kf, df, None, _orc_grid(kf, gp), sf, None, th, False, th_dist)
    pbi, pbd = sc.py_fuse_search(proj, lvl, fl, dmp, kf, df, None, gp, sf, None, th, False, th_dist)
    assert np.array_equal(bi, pbi) and np.array_equal(bd, pbd)
    if seed < 2:
        assert (bi >= 0).sum() > 100


@pytest.mark.parametrize("seed,n1,n2,th,dup", [(0, 500, 520, 7.5, 0.0), (1, 300, 400, 7.5, 0.3), (2, 0, 100, 7.5, 0.0), (3, 100, 0, 7.5, 0.0),
                                               (4, 640, 600, 3.0, 0.1)])
def test_search_by_sim3_equals_restatement(seed, n1, n2, th, dup):
    s1, s2 = sc.sim3_case(seed, n1, n2, dup=dup)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp = sc.grid_params()
    n, m12, m1, m2 = orc.search_by_sim3(s1 + (_orc_grid(s1[4], gp),), s2 + (_orc_grid(s2[4], gp),), sf, th)
    pn, pm12, pm1, pm2 = sc.py_search_by_sim3(s1, s2, gp, sf, th)
    assert n == pn and np.array_equal(m12, pm12) and np.array_equal(m1, pm1) and np.array_equal(m2, pm2)
    if seed == 0:
        assert n > 100 and (m1 >= 0).sum() > n  # some one-way matches fail the agreement check


@pytest.mark.parametrize("seed,nf,nmp,th,dup", [(0, 600, 500, 10, 0.0), (1, 500, 700, 4, 0.0), (2, 0, 50, 10, 0.0), (3, 200, 0, 10, 0.0),
                                                (4, 700, 900, 10, 0.3)])
def test_search_by_projection_sim3_equals_restatement(seed, nf, nmp, th, dup):
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=dup, level0=0.3)
    proj, lvl, fl, _ = sc.fuse_inputs(seed, kf, nmp, stereo=False)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    matched = (np.random.default_rng(seed).random(nf) < 0.15).astype(np.uint8)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp = sc.grid_params()
    n, asg = orc.search_by_projection_sim3(proj, lvl, fl, dmp, kf, df, matched, _orc_grid(kf, gp), sf, th)
    pn, pasg = sc.py_search_by_projection_sim3(proj, lvl, fl, dmp, kf, df, matched, gp, sf, th)
    assert n == pn and np.array_equal(asg, pasg)
    assert not (asg[matched.astype(bool)] >= 0).any()
    if seed == 0:
        assert n > 100


def test_local_map_loop_golden_fixture():
    """tests/golden/matcher/local_map_loop.npz (inputs + results of the Python restatements, tests/golden/make_loop_golden.py) against
    the oracle: local-map search, SearchByPoints, both Fuse searches, SearchBySim3, the Sim3 SearchByProjection."""
    z = np.load(os.path.join(GOLD, "matcher", "local_map_loop.npz"))
    gp = tuple(float(v) for v in z["gp"])
    sf, inv = z["sf"], z["inv_sigma2"]
    kf = z["kf"].view(orc.KP_DTYPE).reshape(-1)
    g = _orc_grid(kf, gp)
    bi, bd = orc.fuse_search(z["proj"], z["level"], z["flags"], z["dmp"], kf, z["df"], z["ur"], g, sf, inv, 3.0)
    assert np.array_equal(bi, z["fuse_idx"]) and np.array_equal(bd, z["fuse_dist"])
    bi, bd = orc.fuse_search(z["proj"], z["level"], z["flags"], z["dmp"], kf, z["df"], None, g, sf, None, 4.0, False, 50)
    assert np.array_equal(bi, z["fuse_sim3_idx"]) and np.array_equal(bd, z["fuse_sim3_dist"])
    n, asg = orc.search_by_projection_sim3(z["proj"], z["level"], z["flags"], z["dmp"], kf, z["df"], z["occ"], g, sf, 10)
    assert n == int(z["proj_sim3_n"]) and np.array_equal(asg, z["proj_sim3_assigned"])
    n, asg = orc.search_map_points(z["lm_proj"], z["lm_view_cos"], z["lm_level"], z["lm_flags"], z["lm_dmp"], kf, z["df"], z["ur"], z["occ"],
                                   g, sf, 3.0, 0.8)
    assert n == int(z["lm_n"]) and np.array_equal(asg, z["lm_assigned"])
    ka, kb = z["sim3_k_a"].view(orc.KP_DTYPE).reshape(-1), z["sim3_k_b"].view(orc.KP_DTYPE).reshape(-1)
    side = lambda t, k: (z["sim3_proj_" + t], z["sim3_level_" + t], z["sim3_flags_" + t], z["sim3_dmp_" + t], k, z["sim3_d_" + t], _orc_grid(k, gp))
    n, m12, m1, m2 = orc.search_by_sim3(side("a", ka), side("b", kb), sf, 7.5)
    assert n == int(z["sim3_n"]) and np.array_equal(m12, z["sim3_m12"]) and np.array_equal(m1, z["sim3_m1"]) and np.array_equal(m2, z["sim3_m2"])
    k1, k2 = z["bp_k1"].view(orc.KP_DTYPE).reshape(-1), z["bp_k2"].view(orc.KP_DTYPE).reshape(-1)
    n, m12 = orc.search_by_points(k1, z["bp_d1"], z["bp_v1"], k2, z["bp_d2"], z["bp_v2"], 0.75, True)
    assert n == int(z["bp_n"]) and np.array_equal(m12, z["bp_m12"])
