"""The guided matchers on the GPU (sdorb_search_for_initialization_batch / sdorb_search_by_projection_batch, through the
C ABI) against the CPU oracle: ORBmatcher::SearchForInitialization (/root/reference/src/ORBmatcher.cc:256-357) and
ORBmatcher::SearchByProjection(Frame&, const Frame&, th, bMono) (:946-1075) over Frame::GetFeaturesInArea
(src/Frame.cc:271-321).  Bit-exact: vnMatches12, vbPrevMatched, the map-point assignment and the return values.
Needs a B200: run with  pytest -m gpu."""
import ctypes as C
import os

import numpy as np
import pytest

import search_cases as sc
from oracle import binding as orc
from sdslam_b200 import api, synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def ex():
    e = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=8)
    yield e
    e.close()


def _slab(arrs, cap, dtype, tail=()):
    out = np.zeros((len(arrs), cap) + tuple(tail), dtype)
    for p, a in enumerate(arrs):
        out[p, :len(a)] = a
    return out


def _init_batch(ex, pairs, gp, window, ratio, orient, cap):
    P = len(pairs)
    k1 = _slab([p[0] for p in pairs], cap, api.KP_DTYPE)
    d1 = _slab([p[1] for p in pairs], cap, np.uint8, (32,))
    k2 = _slab([p[2] for p in pairs], cap, api.KP_DTYPE)
    d2 = _slab([p[3] for p in pairs], cap, np.uint8, (32,))
    n1 = np.array([len(p[0]) for p in pairs], np.int32)
    n2 = np.array([len(p[2]) for p in pairs], np.int32)
    cs, idx = ex.assign_grid_batch(k2, n2, *gp)
    prev = np.stack([k1["x"], k1["y"]], 2)
    return (k1, d1, n1, k2, d2, n2, (cs, idx) + tuple(gp), prev), ex.search_for_initialization_batch(
        k1, d1, n1, k2, d2, n2, (cs, idx) + tuple(gp), prev, window, ratio, orient)


@pytest.mark.parametrize("window,ratio,orient", [(100, 0.9, True), (30, 0.9, True), (100, 0.6, False), (400, 1.0, True)])
def test_search_for_initialization_equals_oracle(ex, window, ratio, orient):
    sizes = [(600, 640, 0.0), (400, 380, 0.3), (500, 500, 0.5), (64, 700, 0.0), (300, 1, 0.0), (0, 50, 0.0), (50, 0, 0.0),
             (700, 700, 0.8), (33, 31, 0.0), (1, 1, 0.0)]
    pairs = [sc.frame_pair(40 + s, a, b, dup=d, flips=20) for s, (a, b, d) in enumerate(sizes)]
    gp = sc.grid_params()
    cap = 704
    (k1, d1, n1, k2, d2, n2, grid, prev), (nm, m12, pm) = _init_batch(ex, pairs, gp, window, ratio, orient, cap)
    total = 0
    for p, (a1, b1, a2, b2) in enumerate(pairs):
        ocs, oidx = orc.assign_grid(a2, *gp)
        on, om12, opm = orc.search_for_initialization(a1, b1, a2, b2, (ocs, oidx) + tuple(gp), prev[p, :len(a1)], window, ratio, orient)
        assert nm[p] == on, "pair %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(m12[p, :len(a1)], om12), "pair %d: vnMatches12" % p
        assert (m12[p, len(a1):] == -1).all()
        assert pm[p, :len(a1)].tobytes() == opm.tobytes(), "pair %d: vbPrevMatched" % p
        assert pm[p, len(a1):].tobytes() == prev[p, len(a1):].tobytes()
        total += on
    assert total > 300
    # second round from the updated vbPrevMatched (Tracking::MonocularInitialization keeps calling it)
    nm2, m12b, pm2 = ex.search_for_initialization_batch(k1, d1, n1, k2, d2, n2, grid, pm, window, ratio, orient)
    for p, (a1, b1, a2, b2) in enumerate(pairs):
        on, om12, opm = orc.search_for_initialization(a1, b1, a2, b2, (grid[0][p], grid[1][p]) + tuple(gp), pm[p, :len(a1)], window,
                                                      ratio, orient)
        assert nm2[p] == on and np.array_equal(m12b[p, :len(a1)], om12) and pm2[p, :len(a1)].tobytes() == opm.tobytes()


def _projection_case(seed, nl, nc, stereo, dup=0.0):
    kl, dl, kc, dc = sc.frame_pair(seed + 20, nl, nc, jitter=4.0, dup=dup, level0=0.3)
    proj, flags, dmp = sc.projection_inputs(seed, kl, dl)
    rng = np.random.default_rng(seed)
    ur = (np.where(rng.random(nc) < 0.6, kc["x"] - rng.uniform(0, 30, nc), -1) if stereo else np.full(nc, -1)).astype(np.float32)
    occ = (rng.random(nc) < 0.1).astype(np.uint8)
    klu = kl.copy()
    klu["angle"] = np.mod(kl["angle"] + np.float32(0.5), 360)  # mvKeysUn carries the angle, mvKeys the octave
    return kl, klu, proj, flags, dmp, kc, dc, ur, occ


@pytest.mark.parametrize("th,mode,orient,stereo", [(15.0, 0, True, False), (7.0, 0, True, True), (15.0, 1, True, True),
                                                   (15.0, 2, False, True), (30.0, 0, True, True)])
def test_search_by_projection_equals_oracle(ex, th, mode, orient, stereo):
    sizes = [(600, 640), (500, 450), (400, 500), (300, 0), (0, 300), (700, 700), (1, 1), (65, 33)]
    cases = [_projection_case(s, a, b, stereo, dup=0.2 if s == 5 else 0.0) for s, (a, b) in enumerate(sizes)]
    cap = 704
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    bounds = (0.0, 640.0, 0.0, 480.0)
    col = lambda j, dt, tail=(): _slab([c[j] for c in cases], cap, dt, tail)
    kl, klu, proj, flags, dmp = col(0, api.KP_DTYPE), col(1, api.KP_DTYPE), col(2, np.float32, (3,)), col(3, np.uint8), col(4, np.uint8, (32,))
    kc, dc, ur, occ = col(5, api.KP_DTYPE), col(6, np.uint8, (32,)), col(7, np.float32), col(8, np.uint8)
    nl = np.array([len(c[0]) for c in cases], np.int32)
    nc = np.array([len(c[5]) for c in cases], np.int32)
    cs, idx = ex.assign_grid_batch(kc, nc, *gp)
    nm, asg = ex.search_by_projection_batch(kl, klu, proj, flags, dmp, nl, kc, dc, ur, occ, nc, (cs, idx) + tuple(gp), sf, bounds, th,
                                            40.0, mode, orient)
    total = 0
    for p, c in enumerate(cases):
        ocs, oidx = orc.assign_grid(c[5], *gp)
        on, oasg = orc.search_by_projection(c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7], c[8], (ocs, oidx) + tuple(gp), sf, bounds,
                                            th, 40.0, mode, orient)
        assert nm[p] == on, "pair %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(asg[p, :len(c[5])], oasg), "pair %d: assignment" % p
        assert (asg[p, len(c[5]):] == -1).all()
        total += on
    assert total > 300


def test_search_golden_fixture(ex):
    """The committed vectors made by the independent Python restatement (tests/golden/make_search_golden.py)."""
    z = np.load(os.path.join(GOLD, "matcher", "search_pairs.npz"))
    gp = tuple(float(v) for v in z["gp"])
    k1, k2 = z["k1"].view(api.KP_DTYPE).reshape(1, -1), z["k2"].view(api.KP_DTYPE).reshape(1, -1)
    cap = max(k1.shape[1], k2.shape[1])
    pad = lambda a, tail=(): _slab([a.reshape((-1,) + tuple(tail))], cap, a.dtype, tail)
    K1, K2, D1, D2 = pad(k1[0]), pad(k2[0]), pad(z["d1"], (32,)), pad(z["d2"], (32,))
    n1, n2 = np.array([k1.shape[1]], np.int32), np.array([k2.shape[1]], np.int32)
    cs, idx = ex.assign_grid_batch(K2, n2, *gp)
    grid = (cs, idx) + gp
    nm, m12, pm = ex.search_for_initialization_batch(K1, D1, n1, K2, D2, n2, grid, pad(z["prev"], (2,)), 100, 0.9, True)
    assert nm[0] == int(z["init_n"]) and np.array_equal(m12[0, :n1[0]], z["init_m12"])
    assert pm[0, :n1[0]].tobytes() == z["init_prev"].tobytes()
    nm, asg = ex.search_by_projection_batch(K1, K1, pad(z["proj"], (3,)), pad(z["flags"]), D1, n1, K2, D2, pad(z["ur"]), pad(z["occ"]),
                                            n2, grid, z["sf"], z["bounds"], 15.0, 40.0, 0, True)
    assert nm[0] == int(z["proj_n"]) and np.array_equal(asg[0, :n2[0]], z["proj_assigned"])


def test_extract_to_initialization_search_on_device_memory(ex):
    """The monocular-initialisation chain on real ORB output with everything resident on the device: extract two shifted
    views of one scene, grid the second, SearchForInitialization (windowSize 100, nnratio 0.9) -- device pointers in, device
    pointers out -- equals the oracle run on the extractor's host output, and finds the shift."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    big = synth.smooth_noise(3, 700, 520)
    imgs = np.stack([big[20:500, 30:670], big[14:494, 19:659]])  # second view shifted by (+11, +6) px
    kps, desc, cnt = ex.extract_batch_host(np.ascontiguousarray(imgs))
    gp = sc.grid_params()
    cap = kps.shape[1]
    prev = np.stack([kps["x"][0], kps["y"][0]], 1)[None].astype(np.float32)
    def t(a):  # keypoint slabs travel as [P, cap, 7] float32 (28-byte cv::KeyPoint rows)
        a = np.ascontiguousarray(a)
        return torch.from_numpy(a.view(np.float32).reshape(a.shape + (7,)) if a.dtype == api.KP_DTYPE else a).to(dev)

    tk1, tk2, td1, td2 = t(kps[0:1]), t(kps[1:2]), t(desc[0:1]), t(desc[1:2])
    tn1, tn2 = t(cnt[0:1].astype(np.int32)), t(cnt[1:2].astype(np.int32))
    tprev = t(prev)
    tcs = torch.zeros((1, 64 * 48 + 1), dtype=torch.int32, device=dev)
    tidx = torch.zeros((1, cap), dtype=torch.int32, device=dev)
    ex.assign_grid_batch(tk2, tn2, *gp, cell_start=tcs, indices=tidx, device=True,
                         stream=torch.cuda.current_stream().cuda_stream)
    tm12 = torch.zeros((1, cap), dtype=torch.int32, device=dev)
    tnm = torch.zeros(1, dtype=torch.int32, device=dev)
    ex.search_for_initialization_batch(tk1, td1, tn1, tk2, td2, tn2, (tcs, tidx) + gp, tprev, 100, 0.9, True, matches12=tm12,
                                       nmatches=tnm, device=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    n1, n2 = int(cnt[0]), int(cnt[1])
    ocs, oidx = orc.assign_grid(kps[1, :n2], *gp)
    on, om12, opm = orc.search_for_initialization(kps[0, :n1], desc[0, :n1], kps[1, :n2], desc[1, :n2], (ocs, oidx) + gp, prev[0, :n1],
                                                  100, 0.9, True)
    assert int(tnm.cpu()[0]) == on and np.array_equal(tm12.cpu().numpy()[0, :n1], om12)
    assert tprev.cpu().numpy()[0, :n1].tobytes() == opm.tobytes()
    m = np.flatnonzero(om12 >= 0)
    assert len(m) > 40
    dx = kps["x"][1][om12[m]] - kps["x"][0][m]
    dy = kps["y"][1][om12[m]] - kps["y"][0][m]
    assert abs(float(np.median(dx)) - 11) <= 1.5 and abs(float(np.median(dy)) - 6) <= 1.5


def test_search_argument_errors(ex):
    k = np.zeros((1, 8), api.KP_DTYPE)
    d = np.zeros((1, 8, 32), np.uint8)
    n = np.zeros(1, np.int32)
    cs, idx = np.zeros((1, 64 * 48 + 1), np.int32), np.zeros((1, 8), np.int32)
    g = api._FrameGrid(api._ptr(cs), api._ptr(idx), 0.0, 0.0, 0.1, 0.1)
    prev, m12, nm = np.zeros((1, 8, 2), np.float32), np.zeros((1, 8), np.int32), np.zeros(1, np.int32)
    L = api.lib()
    call = lambda cap, mem, kp1=k: L.sdorb_search_for_initialization_batch(
        ex._h, api._ptr(kp1), api._ptr(d), api._ptr(n), api._ptr(k), api._ptr(d), api._ptr(n), C.byref(g), 1, cap, api._ptr(prev), 100,
        0.9, 1, api._ptr(m12), api._ptr(nm), mem, None)
    assert call(8, api.MEM_HOST) == 0 and nm[0] == 0
    assert call(0, api.MEM_HOST) < 0 and call(16385, api.MEM_HOST) < 0 and call(8, 7) < 0 and call(8, api.MEM_HOST, None) < 0


# ------------------------------------------------------------------ SearchForTriangulation (row f4, second half)
@pytest.mark.parametrize("orient", [True, False])
def test_search_for_triangulation_equals_oracle(ex, orient):
    """sdorb_search_for_triangulation_batch = ORBmatcher::SearchForTriangulation (src/ORBmatcher.cc:359-462) from the epipole
    on: a ragged batch of keyframe pairs (more keypoints than one CTA's threads, empty frames, duplicate descriptors), each with
    its own F12 and epipole."""
    sizes = [(600, 640, 0.1), (300, 280, 0.3), (1000, 900, 0.0), (0, 50, 0.0), (50, 0, 0.0), (257, 513, 0.5), (1, 1, 0.0)]
    cases = [sc.triangulation_case(s, a, b, dup=d) for s, (a, b, d) in enumerate(sizes)]
    cap = 1024
    col = lambda j, dt, tail=(): _slab([c[j] for c in cases], cap, dt, tail)
    k1, d1, mp1, ur1 = col(0, api.KP_DTYPE), col(1, np.uint8, (32,)), col(2, np.uint8), col(3, np.float32)
    k2, d2, mp2, ur2 = col(4, api.KP_DTYPE), col(5, np.uint8, (32,)), col(6, np.uint8), col(7, np.float32)
    n1 = np.array([len(c[0]) for c in cases], np.int32)
    n2 = np.array([len(c[4]) for c in cases], np.int32)
    F = np.stack([c[8].reshape(9) for c in cases])
    ep = np.array([[c[9], c[10]] for c in cases], np.float32)
    sf, s2 = cases[0][11], cases[0][12]
    nm, m12 = ex.search_for_triangulation_batch(k1, d1, mp1, ur1, n1, k2, d2, mp2, ur2, n2, F, ep, sf, s2, orient)
    total = 0
    for p, c in enumerate(cases):
        on, om12 = orc.search_for_triangulation(*c, orient)
        assert nm[p] == on, "pair %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(m12[p, :len(c[0])], om12), "pair %d: vMatches12" % p
        assert (m12[p, len(c[0]):] == -1).all()
        total += on
    assert total > 150
    # device memory in, device memory out
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a.view(np.float32).reshape(a.shape + (7,)) if a.dtype == api.KP_DTYPE else a).to(dev)
    tm12 = torch.zeros((len(cases), cap), dtype=torch.int32, device=dev)
    tnm = torch.zeros(len(cases), dtype=torch.int32, device=dev)
    ex.search_for_triangulation_batch(t(k1), t(d1), t(mp1), t(ur1), t(n1), t(k2), t(d2), t(mp2), t(ur2), t(n2), t(F), t(ep), sf, s2, orient,
                                      matches12=tm12, nmatches=tnm, device=True, stream=torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    assert np.array_equal(tm12.cpu().numpy(), m12) and np.array_equal(tnm.cpu().numpy(), nm)


def test_triangulation_golden_fixture(ex):
    from test_oracle_search import _load_triangulation_fixture
    c, n, m12 = _load_triangulation_fixture()
    cap = max(len(c[0]), len(c[4]))
    one = lambda a, tail=(): _slab([a], cap, a.dtype, tail)
    nm, gm = ex.search_for_triangulation_batch(one(c[0]), one(c[1], (32,)), one(c[2]), one(c[3]), [len(c[0])], one(c[4]), one(c[5], (32,)),
                                               one(c[6]), one(c[7]), [len(c[4])], c[8].reshape(1, 9), np.array([[c[9], c[10]]], np.float32),
                                               c[11], c[12], True)
    assert nm[0] == n and np.array_equal(gm[0, :len(c[0])], m12)


# ------------------------------------------------------------------ SearchByProjection(Frame, map points): Tracking::SearchLocalPoints
@pytest.mark.parametrize("th,ratio,stereo", [(1.0, 0.8, False), (3.0, 0.8, True), (5.0, 0.6, True)])
def test_search_map_points_equals_oracle(ex, th, ratio, stereo):
    """sdorb_search_map_points_batch = ORBmatcher::SearchByProjection(Frame&, const vector<MapPoint*>&, th)
    (src/ORBmatcher.cc:43-119): ragged batch incl. empty frames / no map points, duplicate descriptors, view cosines around the
    0.998 switch of RadiusByViewingCos."""
    sizes = [(600, 500, 0.0), (500, 700, 0.0), (300, 200, 0.0), (0, 50, 0.0), (200, 0, 0.0), (700, 900, 0.3), (1, 1, 0.0)]
    frames = []
    for s, (nf, nmp, dup) in enumerate(sizes):
        _, _, kf, df = sc.frame_pair(s + 60, 10, nf, dup=dup, level0=0.3)
        proj, vc, lvl, fl, dmp = sc.map_point_inputs(s, kf, df, nmp)
        rng = np.random.default_rng(s)
        ur = (np.where(rng.random(nf) < 0.6, kf["x"] - rng.uniform(0, 30, nf), -1) if stereo else np.full(nf, -1)).astype(np.float32)
        occ = (rng.random(nf) < 0.1).astype(np.uint8)
        frames.append((proj, vc, lvl, fl, dmp, kf, df, ur, occ))
    cap, capmp = 704, 912
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    col = lambda j, c, dt, tail=(): _slab([f[j] for f in frames], c, dt, tail)
    proj, vc, lvl, fl, dmp = col(0, capmp, np.float32, (3,)), col(1, capmp, np.float32), col(2, capmp, np.int32), col(3, capmp, np.uint8), col(4, capmp, np.uint8, (32,))
    kf, df, ur, occ = col(5, cap, api.KP_DTYPE), col(6, cap, np.uint8, (32,)), col(7, cap, np.float32), col(8, cap, np.uint8)
    nmp = np.array([len(f[0]) for f in frames], np.int32)
    nf = np.array([len(f[5]) for f in frames], np.int32)
    cs, idx = ex.assign_grid_batch(kf, nf, *gp)
    nm, asg = ex.search_map_points_batch(proj, vc, lvl, fl, dmp, nmp, kf, df, ur, occ, nf, (cs, idx) + tuple(gp), sf, th, ratio)
    total = 0
    for p, f in enumerate(frames):
        ocs, oidx = orc.assign_grid(f[5], *gp)
        on, oasg = orc.search_map_points(f[0], f[1], f[2], f[3], f[4], f[5], f[6], f[7], f[8], (ocs, oidx) + tuple(gp), sf, th, ratio)
        assert nm[p] == on, "frame %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(asg[p, :len(f[5])], oasg), "frame %d: assignment" % p
        assert (asg[p, len(f[5]):] == -1).all()
        total += on
    assert total > 300


@pytest.mark.parametrize("ratio,orient", [(0.75, True), (0.9, False)])
def test_search_by_points_equals_oracle(ex, ratio, orient):
    """sdorb_search_by_points_batch = ORBmatcher::SearchByPoints (src/ORBmatcher.cc:1209-1304): ragged batch of keyframe pairs,
    keypoints without (good) map points on both sides, duplicate descriptors (contested rows of the second keyframe)."""
    sizes = [(600, 640, 0.0), (400, 380, 0.4), (1000, 1000, 0.1), (0, 50, 0.0), (50, 0, 0.0), (33, 65, 0.6), (1, 1, 0.0)]
    pairs = [sc.frame_pair(80 + s, a, b, dup=d, flips=25) for s, (a, b, d) in enumerate(sizes)]
    rng = np.random.default_rng(3)
    valid = [((rng.random(len(p[0])) < 0.8).astype(np.uint8), (rng.random(len(p[2])) < 0.8).astype(np.uint8)) for p in pairs]
    cap = 1024
    k1, d1 = _slab([p[0] for p in pairs], cap, api.KP_DTYPE), _slab([p[1] for p in pairs], cap, np.uint8, (32,))
    k2, d2 = _slab([p[2] for p in pairs], cap, api.KP_DTYPE), _slab([p[3] for p in pairs], cap, np.uint8, (32,))
    v1, v2 = _slab([v[0] for v in valid], cap, np.uint8), _slab([v[1] for v in valid], cap, np.uint8)
    n1 = np.array([len(p[0]) for p in pairs], np.int32)
    n2 = np.array([len(p[2]) for p in pairs], np.int32)
    nm, m12 = ex.search_by_points_batch(k1, d1, v1, n1, k2, d2, v2, n2, ratio, orient)
    total = 0
    for p, (a1, b1, a2, b2) in enumerate(pairs):
        on, om12 = orc.search_by_points(a1, b1, valid[p][0], a2, b2, valid[p][1], ratio, orient)
        assert nm[p] == on, "pair %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(m12[p, :len(a1)], om12), "pair %d: matches" % p
        assert (m12[p, len(a1):] == -1).all()
        total += on
    assert total > 300


def test_keyframe_projection_overload_on_gpu(ex):
    """SearchByProjection(Frame&, KeyFrame*, sAlreadyFound, th, ORBdist) (src/ORBmatcher.cc:1306-1421) through
    sdorb_search_by_projection_batch with orb_dist = 64, against the direct Python restatement of that overload."""
    cases = [sc.kf_projection_args(s, nk, nc) for s, (nk, nc) in enumerate([(500, 520), (400, 300), (0, 100), (300, 0)])]
    cap = 544
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    gp, bounds = sc.grid_params(), (0.0, 640.0, 0.0, 480.0)
    col = lambda key, dt, tail=(): _slab([c[key] for c in cases], cap, dt, tail)
    kc = col("kc", api.KP_DTYPE)
    nk = np.array([len(c["kk"]) for c in cases], np.int32)
    nc = np.array([len(c["kc"]) for c in cases], np.int32)
    cs, idx = ex.assign_grid_batch(kc, nc, *gp)
    nm, asg = ex.search_by_projection_batch(col("k_level", api.KP_DTYPE), col("kk", api.KP_DTYPE), col("proj", np.float32, (3,)),
                                            col("flags", np.uint8), col("dmp", np.uint8, (32,)), nk, kc, col("dc", np.uint8, (32,)),
                                            col("ur", np.float32), col("has_mp", np.uint8), nc, (cs, idx) + tuple(gp), sf, bounds, 10.0,
                                            0.0, 0, True, orb_dist=64)
    for p, a in enumerate(cases):
        pn, pasg = sc.py_search_by_projection_kf(a["kk"], a["proj"], a["valid"], a["pred"], a["dmp"], a["kc"], a["dc"], a["has_mp"], gp,
                                                 sf, bounds, 10.0, 64, True)
        assert nm[p] == pn and np.array_equal(asg[p, :len(a["kc"])], pasg), "case %d" % p
    assert int(nm.sum()) > 100


@pytest.mark.parametrize("th,stereo,origin", [(3.0, True, None), (2.5, False, None), (4.0, True, (-12.25, -7.5))])
def test_fuse_search_equals_oracle(ex, th, stereo, origin):
    """sdorb_fuse_search_batch = the keypoint search of ORBmatcher::Fuse(KeyFrame*, vpMapPoints, th) (src/ORBmatcher.cc:535-586):
    ragged batch of keyframes incl. empty ones / no map points, duplicate descriptors, reprojection errors on the chi-square
    limits, descriptor distances on both sides of TH_LOW; origin: image bounds of a distorted camera (mnMinX, mnMinY < 0)."""
    sizes = [(600, 500, 0.0), (500, 700, 0.0), (300, 200, 0.0), (0, 50, 0.0), (200, 0, 0.0), (700, 900, 0.3), (1, 1, 0.0)]
    frames = []
    for s, (nf, nmp, dup) in enumerate(sizes):
        _, _, kf, df = sc.frame_pair(s + 140, 10, nf, dup=dup, level0=0.3)
        proj, lvl, fl, ur = sc.fuse_inputs(s, kf, nmp, stereo=stereo)
        frames.append((proj, lvl, fl, sc.fuse_descriptors(s, df, nf, nmp), kf, df, ur))
    cap, capmp = 704, 912
    gp = sc.grid_params() if origin is None else sc.grid_params(667.75, 497.0, *origin)
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    inv = (np.float32(1) / (sf * sf)).astype(np.float32)
    col = lambda j, c, dt, tail=(): _slab([f[j] for f in frames], c, dt, tail)
    kf = col(4, cap, api.KP_DTYPE)
    nmp = np.array([len(f[0]) for f in frames], np.int32)
    nf = np.array([len(f[4]) for f in frames], np.int32)
    cs, idx = ex.assign_grid_batch(kf, nf, *gp)
    bi, bd = ex.fuse_search_batch(col(0, capmp, np.float32, (3,)), col(1, capmp, np.int32), col(2, capmp, np.uint8),
                                  col(3, capmp, np.uint8, (32,)), nmp, kf, col(5, cap, np.uint8, (32,)), col(6, cap, np.float32),
                                  (cs, idx) + tuple(gp), sf, inv, th)
    fused = rejected = 0
    for p, f in enumerate(frames):
        ocs, oidx = orc.assign_grid(f[4], *gp)
        obi, obd = orc.fuse_search(f[0], f[1], f[2], f[3], f[4], f[5], f[6], (ocs, oidx) + tuple(gp), sf, inv, th)
        n = len(f[0])
        assert np.array_equal(bi[p, :n], obi), "keyframe %d: best_idx" % p
        assert np.array_equal(bd[p, :n], obd), "keyframe %d: best_dist" % p
        assert (bi[p, n:] == -1).all() and (bd[p, n:] == 256).all()
        fused += int((obi >= 0).sum())
        rejected += int(((obd > 50) & (obd < 256)).sum())
    assert fused > 500 and rejected > 50


@pytest.mark.parametrize("th,th_dist", [(3.0, 50), (4.0, 100)])
def test_fuse_search_sim3_form_equals_oracle(ex, th, th_dist):
    """check_reprojection = 0: the search of Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) (src/ORBmatcher.cc:682-708); u_right and
    inv_level_sigma2 are NULL."""
    sizes = [(600, 500, 0.0), (0, 50, 0.0), (200, 0, 0.0), (700, 900, 0.3), (1, 1, 0.0)]
    frames = []
    for s, (nf, nmp, dup) in enumerate(sizes):
        _, _, kf, df = sc.frame_pair(s + 140, 10, nf, dup=dup, level0=0.3)
        proj, lvl, fl, _ = sc.fuse_inputs(s, kf, nmp, stereo=False)
        frames.append((proj, lvl, fl, sc.fuse_descriptors(s, df, nf, nmp), kf, df))
    cap, capmp = 704, 912
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    col = lambda j, c, dt, tail=(): _slab([f[j] for f in frames], c, dt, tail)
    kf = col(4, cap, api.KP_DTYPE)
    nmp = np.array([len(f[0]) for f in frames], np.int32)
    nf = np.array([len(f[4]) for f in frames], np.int32)
    cs, idx = ex.assign_grid_batch(kf, nf, *gp)
    bi, bd = ex.fuse_search_batch(col(0, capmp, np.float32, (3,)), col(1, capmp, np.int32), col(2, capmp, np.uint8),
                                  col(3, capmp, np.uint8, (32,)), nmp, kf, col(5, cap, np.uint8, (32,)), None,
                                  (cs, idx) + tuple(gp), sf, None, th, check_reprojection=False, th_dist=th_dist)
    found = 0
    for p, f in enumerate(frames):
        ocs, oidx = orc.assign_grid(f[4], *gp)
        obi, obd = orc.fuse_search(f[0], f[1], f[2], f[3], f[4], f[5], None, (ocs, oidx) + tuple(gp), sf, None, th, False, th_dist)
        n = len(f[0])
        assert np.array_equal(bi[p, :n], obi) and np.array_equal(bd[p, :n], obd), "keyframe %d" % p
        assert (bi[p, n:] == -1).all() and (bd[p, n:] == 256).all()
        found += int((obi >= 0).sum())
    assert found > 300


def test_search_by_sim3_equals_oracle(ex):
    """sdorb_search_by_sim3_batch = ORBmatcher::SearchBySim3 (src/ORBmatcher.cc:734-944) from the projections on: vnMatch1, vnMatch2,
    the agreement check and nFound, on a ragged batch incl. empty keyframes and duplicate descriptors."""
    sizes = [(500, 520, 0.0), (300, 400, 0.3), (0, 100, 0.0), (100, 0, 0.0), (640, 600, 0.1), (1, 1, 0.0)]
    cases = [sc.sim3_case(s, a, b, dup=d) for s, (a, b, d) in enumerate(sizes)]
    cap = 640
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    th = 7.5

    def side(j):
        ks = _slab([c[j][4] for c in cases], cap, api.KP_DTYPE)
        n = np.array([len(c[j][4]) for c in cases], np.int32)
        cs, idx = ex.assign_grid_batch(ks, n, *gp)
        return (_slab([c[j][0] for c in cases], cap, np.float32, (3,)), _slab([c[j][1] for c in cases], cap, np.int32),
                _slab([c[j][2] for c in cases], cap, np.uint8), _slab([c[j][3] for c in cases], cap, np.uint8, (32,)), n, ks,
                _slab([c[j][5] for c in cases], cap, np.uint8, (32,)), (cs, idx) + tuple(gp))
    nf, m12, m1, m2 = ex.search_by_sim3_batch(side(0), side(1), sf, th)
    total = 0
    for p, (s1, s2) in enumerate(cases):
        g1, g2 = orc.assign_grid(s1[4], *gp) + tuple(gp), orc.assign_grid(s2[4], *gp) + tuple(gp)
        on, om12, om1, om2 = orc.search_by_sim3(s1 + (g1,), s2 + (g2,), sf, th)
        n1, n2 = len(s1[4]), len(s2[4])
        assert nf[p] == on, "pair %d: nFound %d vs oracle %d" % (p, nf[p], on)
        assert np.array_equal(m12[p, :n1], om12) and np.array_equal(m1[p, :n1], om1) and np.array_equal(m2[p, :n2], om2), "pair %d" % p
        assert (m12[p, n1:] == -1).all() and (m1[p, n1:] == -1).all() and (m2[p, n2:] == -1).all()
        total += on
    assert total > 300


@pytest.mark.parametrize("th", [10, 4])
def test_search_by_projection_sim3_equals_oracle(ex, th):
    """sdorb_search_map_points_batch with sim3_form = 1 = ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th)
    (src/ORBmatcher.cc:146-254): keypoints already matched on entry, matches occupying their keypoint for the later points."""
    sizes = [(600, 500, 0.0), (500, 700, 0.0), (0, 50, 0.0), (200, 0, 0.0), (700, 900, 0.3), (1, 1, 0.0)]
    frames = []
    for s, (nf, nmp, dup) in enumerate(sizes):
        _, _, kf, df = sc.frame_pair(s + 140, 10, nf, dup=dup, level0=0.3)
        proj, lvl, fl, _ = sc.fuse_inputs(s, kf, nmp, stereo=False)
        matched = (np.random.default_rng(s).random(nf) < 0.15).astype(np.uint8)
        frames.append((proj, lvl, fl, sc.fuse_descriptors(s, df, nf, nmp), kf, df, matched))
    cap, capmp = 704, 912
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    col = lambda j, c, dt, tail=(): _slab([f[j] for f in frames], c, dt, tail)
    kf = col(4, cap, api.KP_DTYPE)
    nmp = np.array([len(f[0]) for f in frames], np.int32)
    nf = np.array([len(f[4]) for f in frames], np.int32)
    cs, idx = ex.assign_grid_batch(kf, nf, *gp)
    nm, asg = ex.search_by_projection_sim3_batch(col(0, capmp, np.float32, (3,)), col(1, capmp, np.int32), col(2, capmp, np.uint8),
                                                 col(3, capmp, np.uint8, (32,)), nmp, kf, col(5, cap, np.uint8, (32,)),
                                                 col(6, cap, np.uint8), nf, (cs, idx) + tuple(gp), sf, th)
    total = 0
    for p, f in enumerate(frames):
        ocs, oidx = orc.assign_grid(f[4], *gp)
        on, oasg = orc.search_by_projection_sim3(f[0], f[1], f[2], f[3], f[4], f[5], f[6], (ocs, oidx) + tuple(gp), sf, th)
        assert nm[p] == on, "keyframe %d: nmatches %d vs oracle %d" % (p, nm[p], on)
        assert np.array_equal(asg[p, :len(f[4])], oasg), "keyframe %d: assignment" % p
        assert (asg[p, len(f[4]):] == -1).all()
        total += on
    assert total > 300


def test_local_map_loop_golden_fixture(ex):
    """tests/golden/matcher/local_map_loop.npz (made by the Python restatements, tests/golden/make_loop_golden.py) through the C ABI:
    local-map search, SearchByPoints, both Fuse searches, SearchBySim3, the Sim3 SearchByProjection."""
    z = np.load(os.path.join(GOLD, "matcher", "local_map_loop.npz"))
    gp = tuple(float(v) for v in z["gp"])
    sf, inv = z["sf"], z["inv_sigma2"]
    one = lambda a, dt=None: np.ascontiguousarray(a if dt is None else a.view(dt).reshape(-1))[None]
    kf = one(z["kf"], api.KP_DTYPE)
    nf, nmp = np.array([kf.shape[1]], np.int32), np.array([len(z["proj"])], np.int32)
    grid = ex.assign_grid_batch(kf, nf, *gp) + gp
    args = (one(z["proj"]), one(z["level"]), one(z["flags"]), one(z["dmp"]), nmp, kf, one(z["df"]))
    bi, bd = ex.fuse_search_batch(*args, one(z["ur"]), grid, sf, inv, 3.0)
    assert np.array_equal(bi[0], z["fuse_idx"]) and np.array_equal(bd[0], z["fuse_dist"])
    bi, bd = ex.fuse_search_batch(*args, None, grid, sf, None, 4.0, check_reprojection=False)
    assert np.array_equal(bi[0], z["fuse_sim3_idx"]) and np.array_equal(bd[0], z["fuse_sim3_dist"])
    nm, asg = ex.search_by_projection_sim3_batch(*args, one(z["occ"]), nf, grid, sf, 10)
    assert nm[0] == int(z["proj_sim3_n"]) and np.array_equal(asg[0], z["proj_sim3_assigned"])
    nm, asg = ex.search_map_points_batch(one(z["lm_proj"]), one(z["lm_view_cos"]), one(z["lm_level"]), one(z["lm_flags"]), one(z["lm_dmp"]),
                                         nmp, kf, one(z["df"]), one(z["ur"]), one(z["occ"]), nf, grid, sf, 3.0, 0.8)
    assert nm[0] == int(z["lm_n"]) and np.array_equal(asg[0], z["lm_assigned"])
    ka, kb = z["sim3_k_a"].view(api.KP_DTYPE).reshape(-1), z["sim3_k_b"].view(api.KP_DTYPE).reshape(-1)
    cap = max(len(ka), len(kb))
    pad = lambda a, tail=(): _slab([a], cap, a.dtype, tail)

    def side(t, k):
        ks, n = pad(k), np.array([len(k)], np.int32)
        return (pad(z["sim3_proj_" + t], (3,)), pad(z["sim3_level_" + t]), pad(z["sim3_flags_" + t]), pad(z["sim3_dmp_" + t], (32,)), n, ks,
                pad(z["sim3_d_" + t], (32,)), ex.assign_grid_batch(ks, n, *gp) + gp)
    nfound, m12, m1, m2 = ex.search_by_sim3_batch(side("a", ka), side("b", kb), sf, 7.5)
    assert nfound[0] == int(z["sim3_n"]) and np.array_equal(m12[0, :len(ka)], z["sim3_m12"])
    assert np.array_equal(m1[0, :len(ka)], z["sim3_m1"]) and np.array_equal(m2[0, :len(kb)], z["sim3_m2"])
    k1, k2 = z["bp_k1"].view(api.KP_DTYPE).reshape(-1), z["bp_k2"].view(api.KP_DTYPE).reshape(-1)
    cap = max(len(k1), len(k2))
    nm, m12 = ex.search_by_points_batch(pad(k1), pad(z["bp_d1"], (32,)), pad(z["bp_v1"]), [len(k1)], pad(k2), pad(z["bp_d2"], (32,)),
                                        pad(z["bp_v2"]), [len(k2)], 0.75, True)
    assert nm[0] == int(z["bp_n"]) and np.array_equal(m12[0, :len(k1)], z["bp_m12"])


def test_fuse_and_sim3_on_device_memory(ex):
    """sdorb_fuse_search_batch / sdorb_search_by_sim3_batch with SDORB_MEM_DEVICE (device pointers in and out, caller's stream)
    equal the host-memory calls."""
    torch = pytest.importorskip("torch")
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a.view(np.float32).reshape(a.shape + (7,)) if a.dtype == api.KP_DTYPE else a)).to(dev)
    stream = torch.cuda.current_stream().cuda_stream
    gp = sc.grid_params()
    sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
    inv = (np.float32(1) / (sf * sf)).astype(np.float32)
    frames = []
    for s, (nf, nmp) in enumerate([(600, 500), (0, 50), (300, 400)]):
        _, _, kf, df = sc.frame_pair(s + 140, 10, nf, level0=0.3)
        proj, lvl, fl, ur = sc.fuse_inputs(s, kf, nmp, stereo=True)
        frames.append((proj, lvl, fl, sc.fuse_descriptors(s, df, nf, nmp), kf, df, ur))
    cap, capmp = 608, 512
    col = lambda j, c, dt, tail=(): _slab([f[j] for f in frames], c, dt, tail)
    kf = col(4, cap, api.KP_DTYPE)
    nmp = np.array([len(f[0]) for f in frames], np.int32)
    nf = np.array([len(f[4]) for f in frames], np.int32)
    cs, idx = ex.assign_grid_batch(kf, nf, *gp)
    host_args = (col(0, capmp, np.float32, (3,)), col(1, capmp, np.int32), col(2, capmp, np.uint8), col(3, capmp, np.uint8, (32,)), nmp, kf,
                 col(5, cap, np.uint8, (32,)), col(6, cap, np.float32))
    bi, bd = ex.fuse_search_batch(*host_args, (cs, idx) + tuple(gp), sf, inv, 3.0)
    tbi = torch.full((len(frames), capmp), -7, dtype=torch.int32, device=dev)
    tbd = torch.full((len(frames), capmp), -7, dtype=torch.int32, device=dev)
    ex.fuse_search_batch(*[t(a) for a in host_args], (t(cs), t(idx)) + tuple(gp), sf, inv, 3.0, best_idx=tbi, best_dist=tbd, device=True,
                         stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(tbi.cpu().numpy(), bi) and np.array_equal(tbd.cpu().numpy(), bd) and (bi >= 0).sum() > 100

    cases = [sc.sim3_case(s, a, b) for s, (a, b) in enumerate([(500, 520), (0, 100), (300, 200)])]
    cap = 544

    def side(j, dev_side):
        ks = _slab([c[j][4] for c in cases], cap, api.KP_DTYPE)
        n = np.array([len(c[j][4]) for c in cases], np.int32)
        cs, idx = ex.assign_grid_batch(ks, n, *gp)
        arrs = [_slab([c[j][0] for c in cases], cap, np.float32, (3,)), _slab([c[j][1] for c in cases], cap, np.int32),
                _slab([c[j][2] for c in cases], cap, np.uint8), _slab([c[j][3] for c in cases], cap, np.uint8, (32,)), n, ks,
                _slab([c[j][5] for c in cases], cap, np.uint8, (32,))]
        if dev_side:
            return tuple(t(a) for a in arrs) + ((t(cs), t(idx)) + tuple(gp),)
        return tuple(arrs) + ((cs, idx) + tuple(gp),)
    nfound, m12, m1, m2 = ex.search_by_sim3_batch(side(0, False), side(1, False), sf, 7.5)
    out = (torch.zeros(len(cases), dtype=torch.int32, device=dev),) + tuple(
        torch.full((len(cases), cap), -7, dtype=torch.int32, device=dev) for _ in range(3))
    ex.search_by_sim3_batch(side(0, True), side(1, True), sf, 7.5, out=out, device=True, stream=stream)
    torch.cuda.synchronize()
    assert np.array_equal(out[0].cpu().numpy(), nfound) and np.array_equal(out[1].cpu().numpy(), m12)
    assert np.array_equal(out[2].cpu().numpy(), m1) and np.array_equal(out[3].cpu().numpy(), m2) and int(nfound.sum()) > 100
