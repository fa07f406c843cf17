"""The oracle's matcher / Frame restatements against the REFERENCE ITSELF: /root/reference/src/{ORBmatcher,Frame,KeyFrame,MapPoint,
Map}.cc compiled unmodified into oracle/_ref/libsdorb_ref.so (oracle/ref_build/Makefile; cv:: surface, Eigen sliver and Converter
stand-in from oracle/ref_compat) and driven through oracle/ref_build/ref_matcher_shim.cc, which builds real Frame / KeyFrame /
MapPoint objects from the same flat arrays the oracle takes.  Rows a11 / a12 / f2 / f3 / f4 of SURVEY.md section 8.

Routines that project map points before they search (the SearchByProjection overloads) are driven with poses that make the
projection exact -- identity rotation, unit depth, fx = fy = 1: the projection of (u, v, 1) is (u, v), invz = 1 -- so the
comparison starts at the (u, v, invz) the oracle is given.  CPU only; skipped where oracle/_ref is not available.
"""
import numpy as np
import pytest

import search_cases as sc
from oracle import binding as orc
from oracle import ref_binding as ref

pytestmark = pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built and /root/reference absent")
SF = (np.float32(1.2) ** np.arange(8)).astype(np.float32)


def _orc_grid(k, gp):
    cs, idx = orc.assign_grid(k, *gp)
    return (cs, idx) + tuple(gp)


# ------------------------------------------------------------------ Frame (row f3)
@pytest.mark.parametrize("seed", range(3))
def test_assign_grid_and_features_in_area_equal_reference(seed):
    """Frame::AssignFeaturesToGrid / PosInGrid / GetFeaturesInArea (src/Frame.cc:179-192, 271-332)."""
    k1, d1, k2, d2 = sc.frame_pair(seed, 300, 700)
    k2["x"][:8] = [0, 5, 15, 634.9, 635, 639.99, 4.99999, 645]  # cell borders: 10 px per cell
    k2["y"][:8] = [0, 5, 15, 474.9, 475, 479.99, 4.99999, 100]
    gp = sc.grid_params(640, 480, -3.5 if seed & 1 else 0.0, 2.25 if seed & 1 else 0.0)
    cs, idx = orc.assign_grid(k2, *gp)
    rcs, ridx = ref.assign_grid(k2, *gp)
    assert np.array_equal(cs, rcs) and np.array_equal(idx, ridx)
    g = (cs, idx) + tuple(gp)
    rng = np.random.default_rng(seed)
    for q in range(200):
        x, y = float(rng.uniform(-80, 720)), float(rng.uniform(-80, 560))
        r = float(rng.choice([0.5, 3, 15, 40, 100, 900]))
        lo, hi = [(-1, -1), (0, -1), (2, -1), (0, 0), (1, 3), (0, 5), (3, 2)][q % 7]
        assert orc.features_in_area(k2, g, x, y, r, lo, hi).tolist() == ref.features_in_area(k2, gp, x, y, r, lo, hi).tolist()


CAMERAS = {  # the TUM1 / TUM2 / EuRoC-like models SD-SLAM is run with
    "tum1": ((517.306408, 516.469215, 318.643040, 255.313989), (0.262383, -0.953104, -0.005358, 0.002628, 1.163314)),
    "tum2": ((520.908620, 521.007327, 325.141442, 249.701764), (0.231222, -0.784899, -0.003257, -0.000105, 0.917205)),
    "euroc": ((458.654, 457.296, 367.215, 248.375), (-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05)),
    "none": ((500.0, 500.0, 320.0, 240.0), (0.0, 0.0, 0.0, 0.0)),
}


@pytest.mark.parametrize("cam", sorted(CAMERAS))
def test_undistort_and_bounds_equal_reference(cam):
    """Frame::UndistortKeyPoints / ComputeImageBounds (src/Frame.cc:335-397): the control flow around cv::undistortPoints --
    the dist[0] == 0 shortcut, the N x 2 -> 2-channel reshape, which fields are kept, which corner feeds which bound."""
    K4, dist = (np.array(v, np.float32) for v in CAMERAS[cam])
    rng = np.random.default_rng(5)
    k = np.zeros(1500, orc.KP_DTYPE)
    k["x"] = rng.uniform(0, 752, len(k)).astype(np.float32)
    k["y"] = rng.uniform(0, 480, len(k)).astype(np.float32)
    k["octave"] = rng.integers(0, 8, len(k))
    k["angle"] = rng.uniform(0, 360, len(k)).astype(np.float32)
    k["response"], k["size"], k["class_id"] = 33, 31, -1
    assert orc.undistort_keypoints(k, K4, dist).tobytes() == ref.undistort_keypoints(k, K4, dist).tobytes()
    assert orc.image_bounds(752, 480, K4, dist).tobytes() == ref.image_bounds(752, 480, K4, dist).tobytes()


def test_stereo_from_rgbd_equals_reference():
    """Frame::ComputeStereoFromRGBD (src/Frame.cc:399-417), including Mat::at<float>(float, float) truncating its arguments."""
    rng = np.random.default_rng(22)
    depth = rng.uniform(-1, 8, (48, 64)).astype(np.float32)
    depth[depth < 0.5] = 0
    k = np.zeros(200, orc.KP_DTYPE)
    k["x"] = rng.uniform(0, 63.9, len(k)).astype(np.float32)
    k["y"] = rng.uniform(0, 47.9, len(k)).astype(np.float32)
    ku = k.copy()
    ku["x"] += rng.uniform(-1, 1, len(k)).astype(np.float32)
    for mbf in (40.0, 386.1448):
        our, oz = orc.stereo_from_rgbd(k, ku, depth, mbf)
        rur, rz = ref.stereo_from_rgbd(k, ku, depth, mbf)
        assert our.tobytes() == rur.tobytes() and oz.tobytes() == rz.tobytes()


# ------------------------------------------------------------------ ORBmatcher (rows a12, f2)
def test_three_maxima_equals_reference():
    rng = np.random.default_rng(5)
    cases = [rng.integers(0, 40, 30) for _ in range(200)] + [np.zeros(30, int), np.full(30, 7), np.arange(30), np.arange(30)[::-1]]
    cases += [np.array([100] + [9] * 29), np.array([100, 10] + [9] * 28), np.array([100, 10, 10] + [0] * 27), np.array([0] * 29 + [1])]
    for s in cases:
        assert orc.three_maxima(s) == ref.three_maxima(s)


@pytest.mark.parametrize("seed,n1,n2,ratio,orient,dup", [(0, 300, 320, 0.75, True, 0.0), (1, 200, 150, 0.9, False, 0.4), (2, 0, 40, 0.75, True, 0.0),
                                                         (3, 40, 0, 0.75, True, 0.0), (4, 250, 250, 0.75, True, 0.6), (5, 1000, 1000, 0.75, True, 0.1)])
def test_search_by_points_equals_reference(seed, n1, n2, ratio, orient, dup):
    """ORBmatcher::SearchByPoints (src/ORBmatcher.cc:1209-1304): the best / second-best scan that row a12 batches, with its
    vbMatched2 bookkeeping, the ratio test and the rotation histogram."""
    k1, d1, k2, d2 = sc.frame_pair(seed + 80, n1, n2, dup=dup, flips=25)
    rng = np.random.default_rng(seed)
    v1, v2 = (rng.random(n1) < 0.8).astype(np.uint8), (rng.random(n2) < 0.8).astype(np.uint8)
    n, m12 = orc.search_by_points(k1, d1, v1, k2, d2, v2, ratio, orient)
    rn, rm12 = ref.search_by_points(k1, d1, v1, k2, d2, v2, ratio, orient)
    assert n == rn and np.array_equal(m12, rm12)
    if seed in (0, 5):
        assert n > 50


def test_best_two_scan_equals_reference_loop():
    """The batched best / second-best (orc_match_greedy: what sdorb_match_greedy_batch computes) IS SearchByPoints without the
    orientation check: accepted matches of the scan == matches12 of the reference."""
    k1, d1, k2, d2 = sc.frame_pair(91, 400, 420, dup=0.3, flips=25)
    v1, v2 = np.ones(len(k1), np.uint8), np.ones(len(k2), np.uint8)
    rn, rm12 = ref.search_by_points(k1, d1, v1, k2, d2, v2, 0.75, False)
    m = orc.match_best2(d1, d2, ratio=0.75, th_low=50, greedy=True)
    got = np.where(m["accepted"] != 0, m["best_idx"], -1)
    assert np.array_equal(got, rm12) and rn == int((got >= 0).sum()) and rn > 50


@pytest.mark.parametrize("seed,n1,n2,window,ratio,orient,dup", [
    (0, 600, 640, 100, 0.9, True, 0.0), (1, 400, 380, 30, 0.9, True, 0.3), (2, 500, 500, 100, 0.6, False, 0.5),
    (3, 64, 700, 15, 1.0, True, 0.0), (4, 300, 1, 100, 0.9, True, 0.0), (5, 0, 50, 100, 0.9, True, 0.0),
    (6, 50, 0, 100, 0.9, True, 0.0), (7, 700, 700, 400, 0.9, True, 0.8)])
def test_search_for_initialization_equals_reference(seed, n1, n2, window, ratio, orient, dup):
    """ORBmatcher::SearchForInitialization (src/ORBmatcher.cc:256-357) over Frame::GetFeaturesInArea, twice (the second call
    continues from the updated vbPrevMatched, as Tracking::MonocularInitialization does)."""
    k1, d1, k2, d2 = sc.frame_pair(seed, n1, n2, dup=dup, flips=20)
    gp = sc.grid_params()
    prev = np.stack([k1["x"], k1["y"]], 1) if n1 else np.zeros((0, 2), np.float32)
    n, m12, pm = orc.search_for_initialization(k1, d1, k2, d2, _orc_grid(k2, gp), prev, window, ratio, orient)
    rn, rm12, rpm = ref.search_for_initialization(k1, d1, k2, d2, gp, prev, window, ratio, orient)
    assert n == rn and np.array_equal(m12, rm12) and pm.tobytes() == rpm.tobytes()
    if seed == 0:
        assert n > 100
    n2_, m12b, _ = orc.search_for_initialization(k1, d1, k2, d2, _orc_grid(k2, gp), pm, window, ratio, orient)
    rn2, rm12b, _ = ref.search_for_initialization(k1, d1, k2, d2, gp, rpm, window, ratio, orient)
    assert n2_ == rn2 and np.array_equal(m12b, rm12b)


@pytest.mark.parametrize("seed,nl,nc,th,mode,orient,stereo", [
    (0, 600, 640, 15.0, 0, True, False), (1, 500, 450, 7.0, 0, True, True), (2, 400, 500, 15.0, 1, True, True),
    (3, 400, 500, 15.0, 2, False, True), (4, 300, 0, 15.0, 0, True, False), (5, 0, 300, 15.0, 0, True, False),
    (6, 700, 700, 30.0, 0, True, True)])
def test_search_by_projection_equals_reference(seed, nl, nc, th, mode, orient, stereo):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) (src/ORBmatcher.cc:946-1075): bounds test, level
    windows of the three modes, occupancy (also of keypoints taken earlier in the same call), the stereo gate, the histogram."""
    kl, dl, kc, dc = sc.frame_pair(seed + 20, nl, nc, jitter=4.0, dup=0.2 if seed == 6 else 0.0, level0=0.3)
    proj, flags, dmp = sc.projection_inputs(seed, kl, dl)
    proj[:, 2] = 1.0  # unit depth: the reference's own projection of (u, v, 1) is exact
    rng = np.random.default_rng(seed)
    ur = np.where(rng.random(nc) < 0.6, kc["x"] - rng.uniform(0, 30, nc), -1).astype(np.float32) if stereo else np.full(nc, -1, np.float32)
    occ = (rng.random(nc) < 0.1).astype(np.uint8)
    gp = sc.grid_params()
    bounds = (0.0, 640.0, 0.0, 480.0)
    klu = kl.copy()
    n, asg = orc.search_by_projection(kl, klu, proj, flags, dmp, kc, dc, ur, occ, _orc_grid(kc, gp), SF, bounds, th, 40.0, mode, orient)
    rn, rasg = ref.search_by_projection(kl, klu, proj, flags, dmp, kc, dc, ur, occ, gp, SF, bounds, th, 40.0, mode, orient)
    assert n == rn and np.array_equal(asg, rasg)
    # the KeyFrame twin SearchByProjection(Frame&, KeyFrame*, th, bMono) (:1077-1207) gives the same result on the same data
    kn, kasg = ref.search_by_projection(kl, klu, proj, flags, dmp, kc, dc, ur, occ, gp, SF, bounds, th, 40.0, mode, orient, keyframe_overload=True)
    assert n == kn and np.array_equal(asg, kasg)
    if seed == 0:
        assert n > 100


@pytest.mark.parametrize("seed,nf,nmp,th,ratio,stereo", [(0, 600, 500, 1.0, 0.8, False), (1, 500, 700, 3.0, 0.8, True), (2, 300, 200, 5.0, 0.6, True),
                                                         (3, 0, 50, 1.0, 0.8, False), (4, 200, 0, 1.0, 0.8, False), (5, 700, 900, 1.0, 1.0, True)])
def test_search_map_points_equals_reference(seed, nf, nmp, th, ratio, stereo):
    """ORBmatcher::SearchByProjection(Frame&, vpMapPoints, th) (src/ORBmatcher.cc:43-126), the local-map search: RadiusByViewingCos,
    the (level - 1, level) window, best / second-best with their levels, the ratio rule that only applies on equal levels."""
    _, _, kf, df = sc.frame_pair(seed + 60, 10, nf, dup=0.3 if seed == 5 else 0.0, level0=0.3)
    proj, vc, lvl, fl, dmp = sc.map_point_inputs(seed, kf, df, nmp)
    rng = np.random.default_rng(seed)
    ur = (np.where(rng.random(nf) < 0.6, kf["x"] - rng.uniform(0, 30, nf), -1) if stereo else np.full(nf, -1)).astype(np.float32)
    occ = (rng.random(nf) < 0.1).astype(np.uint8)
    gp = sc.grid_params()
    n, asg = orc.search_map_points(proj, vc, lvl, fl, dmp, kf, df, ur, occ, _orc_grid(kf, gp), SF, th, ratio)
    rn, rasg = ref.search_map_points(proj, vc, lvl, fl, dmp, kf, df, ur, occ, gp, SF, (0.0, 640.0, 0.0, 480.0), th, ratio)
    assert n == rn and np.array_equal(asg, rasg)
    if seed == 0:
        assert n > 100


# ------------------------------------------------------------------ MapPoint (row f4)
def test_distinctive_descriptors_equal_reference():
    """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:225-284) on real MapPoint / KeyFrame objects: the observation
    map is keyed by KeyFrame POINTER, so the order in which the reference visits the descriptors is an address order; the
    median rule itself does not depend on it except through ties -- the winner's DESCRIPTOR must be one the oracle's rule allows."""
    rng = np.random.default_rng(12)
    for n in [1, 2, 3, 4, 5, 8, 17, 40]:
        d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
        if n >= 5:
            d[1:n // 2] = d[0] ^ (rng.integers(0, 256, (n // 2 - 1, 32), dtype=np.uint8) & 0x11)
        r = ref.distinctive(d)
        best, med = orc.distinctive_many(d, np.array([0, n], np.int32))
        D = np.unpackbits(d[:, None, :] ^ d[None, :, :], axis=2).sum(axis=2)
        medians = np.sort(D, axis=1)[:, int(0.5 * (n - 1))]
        assert medians[r] == medians.min() == med[0], "n = %d: the reference's pick does not have the least median" % n
        assert medians[best[0]] == medians.min()


# ------------------------------------------------------------------ the keyframe-side searches (rows f2 / f4)
@pytest.mark.parametrize("seed,n1,n2,orient", [(0, 140, 150, True), (1, 120, 90, False), (2, 60, 200, True), (3, 0, 40, True), (4, 40, 0, True),
                                               (5, 150, 150, True), (6, 300, 320, True)])
def test_search_for_triangulation_equals_reference(seed, n1, n2, orient):
    """ORBmatcher::SearchForTriangulation + CheckDistEpipolarLine (src/ORBmatcher.cc:128-144, 359-462): the epipolar gate in the
    float / double mix and with the FMA contraction g++ applies to the reference's own expression under its own flags."""
    case = sc.triangulation_case(seed, n1, n2, dup=0.3 if seed == 5 else 0.1)
    n, m12 = orc.search_for_triangulation(*case, orient)
    rn, rm12 = ref.search_for_triangulation(*case, orient)
    assert n == rn and np.array_equal(m12, rm12)
    if seed in (0, 6):
        assert n > 10


def _fuse_case(seed, nf, nmp, stereo, dup, gp, bf=40.0):
    """fuse_inputs of tests/search_cases.py, with the right coordinate in the form the reference can produce: ur = u - bf * invz
    for a depth that is a power of two, the stereo keypoints' own right coordinates placed around it."""
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=dup, level0=0.3)
    proj, lvl, fl, ur_kf = sc.fuse_inputs(seed, kf, nmp, stereo=stereo)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    rng = np.random.default_rng(seed + 77)
    invz = rng.choice(np.array([0.25, 0.5, 1.0, 2.0], np.float32), nmp)
    proj[:, 0] = np.maximum(proj[:, 0], np.float32(gp[0] + 1))  # KeyFrame::IsInImage: inside the (integer) lower bounds
    proj[:, 1] = np.maximum(proj[:, 1], np.float32(gp[1] + 1))
    proj[:, 2] = proj[:, 0] - np.float32(bf) * invz                # one rounding, like the reference's (contracted) u - bf * invz
    if nf and stereo:  # every stereo keypoint's right coordinate near the ur of some point projected next to it
        r0 = np.random.default_rng(seed + 1300)
        r0.integers(0, 8, nmp)
        r0.random(nf), r0.uniform(0, 30, nf)
        src = r0.integers(0, nf, nmp)
        ur_kf = np.where(ur_kf >= 0, ur_kf, -1).astype(np.float32)
        for i in range(nmp):
            if ur_kf[src[i]] >= 0:
                ur_kf[src[i]] = max(np.float32(proj[i, 2] + rng.uniform(-1.5, 1.5)), np.float32(0))
    return kf, df, proj, invz, lvl, fl, dmp, ur_kf.astype(np.float32)


@pytest.mark.parametrize("seed,nf,nmp,th,stereo,dup,origin", [(0, 600, 500, 3.0, True, 0.0, (0.0, 0.0)), (1, 500, 700, 2.5, False, 0.0, (0.0, 0.0)),
                                                              (2, 0, 50, 3.0, True, 0.0, (0.0, 0.0)), (3, 200, 0, 3.0, True, 0.0, (0.0, 0.0)),
                                                              (4, 700, 900, 4.0, True, 0.3, (0.0, 0.0)), (6, 600, 600, 4.0, True, 0.0, (-12.0, -7.0))])
def test_fuse_search_equals_reference(seed, nf, nmp, th, stereo, dup, origin):
    """The keypoint search of ORBmatcher::Fuse (src/ORBmatcher.cc:477-615) on real KeyFrame / MapPoint objects: projection,
    PredictScale, KeyFrame::GetFeaturesInArea, the level window, the chi-square gate as the reference build contracts it, the
    first-minimum rule; read back through the map surgery Fuse performs (AddMapPoint / Replace)."""
    gp = sc.grid_params(640, 480, *origin)
    kf, df, proj, invz, lvl, fl, dmp, ur = _fuse_case(seed, nf, nmp, stereo, dup, gp)
    inv = (np.float32(1) / (SF * SF)).astype(np.float32)
    bi, bd = orc.fuse_search(proj, lvl, fl, dmp, kf, df, ur, _orc_grid(kf, gp), SF, inv, th)
    rbi = ref.fuse_search(proj, invz, lvl, fl, dmp, kf, df, ur, gp, SF, th, 40.0)
    assert np.array_equal(bi, rbi)
    if seed == 0:
        assert (bi >= 0).sum() > 100


@pytest.mark.parametrize("seed,n1,n2,th,dup", [(0, 500, 520, 7.5, 0.0), (1, 300, 400, 7.5, 0.3), (2, 0, 100, 7.5, 0.0), (3, 100, 0, 7.5, 0.0),
                                               (4, 640, 600, 3.0, 0.1)])
def test_search_by_sim3_equals_reference(seed, n1, n2, th, dup):
    """ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:734-944) at the identity transform: both directed searches and the agreement check."""
    s1, s2 = sc.sim3_case(seed, n1, n2, dup=dup)
    gp = sc.grid_params()
    for s in (s1, s2):  # inside the lower image bounds (KeyFrame::IsInImage)
        s[0][:, 0] = np.maximum(s[0][:, 0], np.float32(1))
        s[0][:, 1] = np.maximum(s[0][:, 1], np.float32(1))
        s[0][:, 2] = 1.0
    n, m12, m1, m2 = orc.search_by_sim3(s1 + (_orc_grid(s1[4], gp),), s2 + (_orc_grid(s2[4], gp),), SF, th)
    rn, rm12 = ref.search_by_sim3(s1, s2, gp, SF, th)
    assert n == rn and np.array_equal(m12, rm12)
    if seed == 0:
        assert n > 100


def _inside(proj):
    proj[:, 0] = np.maximum(proj[:, 0], np.float32(1))  # KeyFrame::IsInImage: inside the lower bounds
    proj[:, 1] = np.maximum(proj[:, 1], np.float32(1))
    return proj


@pytest.mark.parametrize("seed,nf,nmp,th", [(0, 600, 500, 3.0), (1, 500, 700, 4.0), (2, 0, 50, 3.0), (4, 700, 900, 4.0)])
def test_fuse_sim3_form_equals_reference(seed, nf, nmp, th):
    """The keypoint search of the Sim3 overload Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:617-732): no
    reprojection gate, TH_LOW; read back through AddMapPoint / vpReplacePoint."""
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=0.3 if seed == 4 else 0.0, level0=0.3)
    proj, lvl, fl, _ = sc.fuse_inputs(seed, kf, nmp, stereo=False)
    proj = _inside(proj)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    gp = sc.grid_params()
    bi, bd = orc.fuse_search(proj, lvl, fl, dmp, kf, df, None, _orc_grid(kf, gp), SF, None, th, False, 50)
    rbi = ref.fuse_sim3_search(proj, lvl, fl, dmp, kf, df, gp, SF, th)
    assert np.array_equal(bi, rbi)
    if seed < 2:
        assert (bi >= 0).sum() > 100


@pytest.mark.parametrize("seed,nf,nmp,th,dup", [(0, 600, 500, 10, 0.0), (1, 500, 700, 4, 0.0), (2, 0, 50, 10, 0.0), (3, 200, 0, 10, 0.0),
                                                (4, 700, 900, 10, 0.3)])
def test_search_by_projection_sim3_equals_reference(seed, nf, nmp, th, dup):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) (src/ORBmatcher.cc:146-254): keypoints taken on entry or
    earlier in the call are skipped."""
    _, _, kf, df = sc.frame_pair(seed + 140, 10, nf, dup=dup, level0=0.3)
    proj, lvl, fl, _ = sc.fuse_inputs(seed, kf, nmp, stereo=False)
    proj = _inside(proj)
    dmp = sc.fuse_descriptors(seed, df, nf, nmp)
    matched = (np.random.default_rng(seed).random(nf) < 0.15).astype(np.uint8)
    gp = sc.grid_params()
    n, asg = orc.search_by_projection_sim3(proj, lvl, fl, dmp, kf, df, matched, _orc_grid(kf, gp), SF, th)
    rn, rasg = ref.search_by_projection_sim3(proj, lvl, fl, dmp, kf, df, matched, gp, SF, th)
    assert n == rn and np.array_equal(asg, rasg)
    if seed == 0:
        assert n > 100


@pytest.mark.parametrize("seed,nk,nc,th,orb_dist,orient", [(0, 500, 520, 10.0, 100, True), (1, 400, 300, 3.0, 64, True), (2, 300, 400, 10.0, 64, False),
                                                           (3, 0, 100, 10.0, 100, True)])
def test_relocalisation_projection_search_equals_reference(seed, nk, nc, th, orb_dist, orient):
    """The relocalisation overload SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1306-1421)
    against the oracle's Frame / Frame function driven with the mapping of sdorb_projection_search::orb_dist."""
    a = sc.kf_projection_args(seed, nk, nc)
    gp, bounds = sc.grid_params(), (0.0, 640.0, 0.0, 480.0)
    n, asg = orc.search_by_projection(a["k_level"], a["kk"], a["proj"], a["flags"], a["dmp"], a["kc"], a["dc"], a["ur"], a["has_mp"],
                                      _orc_grid(a["kc"], gp), SF, bounds, th, 0.0, 0, orient, orb_dist=orb_dist)
    rn, rasg = ref.search_by_projection_reloc(a["kk"], a["proj"], a["valid"], a["pred"], a["dmp"], a["kc"], a["dc"], a["has_mp"], gp, SF, bounds, th,
                                              orb_dist, orient)
    assert n == rn and np.array_equal(asg, rasg)
    if seed == 0:
        assert n > 100
