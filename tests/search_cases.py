"""Synthetic frame pairs for the guided matchers (ORBmatcher::SearchForInitialization / SearchByProjection(Frame, Frame),
/root/reference/src/ORBmatcher.cc:256-357, 946-1075) and a direct Python restatement of both, written from the reference
text on Python lists and a dict-of-lists grid (Frame::mGrid) -- deliberately sharing nothing with oracle/sdorb_oracle.cc.
Used by tests/test_oracle_search.py (pins the oracle) and tests/test_gpu_search.py (kernels against the oracle)."""
import math

import numpy as np

from oracle import binding as orc

F32 = np.float32
COLS, ROWS = 64, 48
TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # src/ORBmatcher.cc:36-38
INT_MAX = 2 ** 31 - 1


def frame_pair(seed, n1=600, n2=640, width=640, height=480, nlevels=8, jitter=6.0, flips=12, dup=0.0, level0=0.6):
    """Frame 1: random keypoints; frame 2: the same points moved by up to `jitter` px with `flips` descriptor bits
    flipped, shuffled, padded with unrelated ones.  `dup` = fraction of frame-2 descriptors that are exact copies of
    another one (distance ties and contested matches)."""
    rng = np.random.default_rng(seed)
    k1 = np.zeros(n1, orc.KP_DTYPE)
    k1["x"] = rng.uniform(-5, width + 5, n1).astype(F32)
    k1["y"] = rng.uniform(-5, height + 5, n1).astype(F32)
    k1["octave"] = np.where(rng.random(n1) < level0, 0, rng.integers(0, nlevels, n1))
    k1["angle"] = rng.uniform(0, 360, n1).astype(F32)
    k1["size"], k1["class_id"] = 31, -1
    d1 = rng.integers(0, 256, (n1, 32), dtype=np.uint8)
    k2 = np.zeros(n2, orc.KP_DTYPE)
    d2 = rng.integers(0, 256, (n2, 32), dtype=np.uint8)
    k2["x"] = rng.uniform(-5, width + 5, n2).astype(F32)
    k2["y"] = rng.uniform(-5, height + 5, n2).astype(F32)
    k2["octave"] = np.where(rng.random(n2) < level0, 0, rng.integers(0, nlevels, n2))
    k2["angle"] = rng.uniform(0, 360, n2).astype(F32)
    k2["size"], k2["class_id"] = 31, -1
    m = min(n1, n2)
    src = rng.permutation(n1)[:m]
    dst = rng.permutation(n2)[:m]
    k2["x"][dst] = k1["x"][src] + rng.uniform(-jitter, jitter, m).astype(F32)
    k2["y"][dst] = k1["y"][src] + rng.uniform(-jitter, jitter, m).astype(F32)
    lvl = k1["octave"][src] + np.where(rng.random(m) < 0.15, rng.integers(-1, 2, m), 0)
    k2["octave"][dst] = np.clip(lvl, 0, nlevels - 1)
    rot = F32(rng.uniform(0, 360))  # one dominant rotation plus outliers: the histogram check has something to cut
    noise = np.where(rng.random(m) < 0.2, rng.uniform(0, 360, m), rng.normal(0, 4, m))
    k2["angle"][dst] = np.mod(k1["angle"][src] - rot + noise, 360).astype(F32)
    dd = d1[src].copy()
    for r in range(m):
        bits = rng.choice(256, int(rng.integers(0, flips + 1)), replace=False)
        for b in bits:
            dd[r, b >> 3] ^= 1 << (b & 7)
    d2[dst] = dd
    ndup = int(dup * n2)
    if ndup:
        a, b = rng.integers(0, n2, ndup), rng.integers(0, n2, ndup)
        d2[a] = d2[b]
        k2["x"][a] = k2["x"][b] + rng.integers(-2, 3, ndup).astype(F32)
        k2["y"][a] = k2["y"][b] + rng.integers(-2, 3, ndup).astype(F32)
        k2["octave"][a] = k2["octave"][b]
    return k1, d1, k2, d2


def grid_params(width=640, height=480, min_x=0.0, min_y=0.0):
    """mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv as Frame's constructor computes them
    (src/Frame.cc:106-107, 161-162) for bounds [min, min + size]."""
    inv_w = F32(COLS) / F32(F32(min_x + width) - F32(min_x))
    inv_h = F32(ROWS) / F32(F32(min_y + height) - F32(min_y))
    return float(F32(min_x)), float(F32(min_y)), float(inv_w), float(inv_h)


def _round_away(v):
    v = float(v)
    return int(math.copysign(math.floor(abs(v) + 0.5), v))


def py_grid(kps, gp):
    """Frame::AssignFeaturesToGrid with PosInGrid (src/Frame.cc:179-192, 323-332): {(ix, iy): [i, ...]}."""
    min_x, min_y, inv_w, inv_h = (F32(v) for v in gp)
    grid = {}
    for i in range(len(kps)):
        px = _round_away(F32(F32(kps["x"][i]) - min_x) * inv_w)
        py = _round_away(F32(F32(kps["y"][i]) - min_y) * inv_h)
        if 0 <= px < COLS and 0 <= py < ROWS:
            grid.setdefault((px, py), []).append(i)
    return grid


def py_features_in_area(kps, grid, gp, x, y, r, min_level, max_level):
    """Frame::GetFeaturesInArea, src/Frame.cc:271-321."""
    min_x, min_y, inv_w, inv_h = (F32(v) for v in gp)
    x, y, r = F32(x), F32(y), F32(r)
    out = []
    c0 = max(0, int(math.floor(F32(F32(F32(x - min_x) - r) * inv_w))))
    if c0 >= COLS:
        return out
    c1 = min(COLS - 1, int(math.ceil(F32(F32(F32(x - min_x) + r) * inv_w))))
    if c1 < 0:
        return out
    r0 = max(0, int(math.floor(F32(F32(F32(y - min_y) - r) * inv_h))))
    if r0 >= ROWS:
        return out
    r1 = min(ROWS - 1, int(math.ceil(F32(F32(F32(y - min_y) + r) * inv_h))))
    if r1 < 0:
        return out
    check = min_level > 0 or max_level >= 0
    for ix in range(c0, c1 + 1):
        for iy in range(r0, r1 + 1):
            for i in grid.get((ix, iy), ()):
                o = int(kps["octave"][i])
                if check:
                    if o < min_level:
                        continue
                    if max_level >= 0 and o > max_level:
                        continue
                if abs(F32(kps["x"][i]) - x) < r and abs(F32(kps["y"][i]) - y) < r:
                    out.append(i)
    return out


def py_distance(a, b):
    return int(np.unpackbits(np.bitwise_xor(a, b)).sum())


def py_three_maxima(sizes):
    """ORBmatcher::ComputeThreeMaxima, src/ORBmatcher.cc:1423-1454."""
    m1 = m2 = m3 = 0
    i1 = i2 = i3 = -1
    for i, s in enumerate(sizes):
        if s > m1:
            m3, m2, m1 = m2, m1, s
            i3, i2, i1 = i2, i1, i
        elif s > m2:
            m3, m2 = m2, s
            i3, i2 = i2, i
        elif s > m3:
            m3, i3 = s, i
    if F32(m2) < F32(0.1) * F32(m1):
        i2 = i3 = -1
    elif F32(m3) < F32(0.1) * F32(m1):
        i3 = -1
    return i1, i2, i3


def _bin(a1, a2):
    rot = F32(a1) - F32(a2)
    if rot < 0:
        rot = F32(rot + F32(360))
    b = _round_away(F32(rot * F32(F32(1) / F32(HISTO_LENGTH))))
    return 0 if b == HISTO_LENGTH else b


def py_search_for_initialization(k1, d1, k2, d2, gp, prev, window, nnratio, check_orientation):
    """ORBmatcher::SearchForInitialization, src/ORBmatcher.cc:256-357: (nmatches, vnMatches12, vbPrevMatched)."""
    grid = py_grid(k2, gp)
    prev = np.array(prev, F32).reshape(-1, 2).copy()
    n = 0
    m12 = [-1] * len(k1)
    hist = [[] for _ in range(HISTO_LENGTH)]
    mdist = [INT_MAX] * len(k2)
    m21 = [-1] * len(k2)
    for i1 in range(len(k1)):
        lvl = int(k1["octave"][i1])
        if lvl > 0:
            continue
        cand = py_features_in_area(k2, grid, gp, prev[i1, 0], prev[i1, 1], window, lvl, lvl)
        if not cand:
            continue
        best, best2, bidx = INT_MAX, INT_MAX, -1
        for i2 in cand:
            d = py_distance(d1[i1], d2[i2])
            if mdist[i2] <= d:
                continue
            if d < best:
                best2, best, bidx = best, d, i2
            elif d < best2:
                best2 = d
        if best <= TH_LOW and F32(best) < F32(F32(best2) * F32(nnratio)):
            if m21[bidx] >= 0:
                m12[m21[bidx]] = -1
                n -= 1
            m12[i1], m21[bidx], mdist[bidx] = bidx, i1, best
            n += 1
            if check_orientation:
                hist[_bin(k1["angle"][i1], k2["angle"][bidx])].append(i1)
    if check_orientation:
        keep = py_three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                if m12[i1] >= 0:
                    m12[i1] = -1
                    n -= 1
    for i1 in range(len(k1)):
        if m12[i1] >= 0:
            prev[i1] = (k2["x"][m12[i1]], k2["y"][m12[i1]])
    return n, np.array(m12, np.int32), prev


def projection_inputs(seed, k1, d1, nlevels=8, width=640, height=480):
    """What the caller of SearchByProjection supplies for the last frame: the projected (u, v, invzc), the flags and
    the map points' descriptors -- here frame 1's keypoints moved a little, some without map point, some behind the camera,
    some outside the bounds."""
    rng = np.random.default_rng(seed + 77)
    n = len(k1)
    proj = np.zeros((n, 3), F32)
    proj[:, 0] = k1["x"] + rng.uniform(-3, 3, n).astype(F32)
    proj[:, 1] = k1["y"] + rng.uniform(-3, 3, n).astype(F32)
    proj[:, 2] = (1.0 / rng.uniform(0.5, 12, n)).astype(F32)
    proj[rng.random(n) < 0.05, 2] *= -1
    flags = (rng.random(n) < 0.8).astype(np.uint8) | ((rng.random(n) < 0.7).astype(np.uint8) << 1)
    desc_mp = d1.copy()
    return proj, flags, desc_mp


def py_search_by_projection(kl, klu, proj, flags, desc_mp, kc, dc, ur_c, occ_c, gp, scale_factors, bounds, th, mbf, mode,
                            check_orientation):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono), src/ORBmatcher.cc:946-1075, from :983 on:
    (nmatches, assigned)."""
    grid = py_grid(kc, gp)
    n = 0
    assigned = [-1] * len(kc)
    occ = [bool(v) for v in occ_c]
    hist = [[] for _ in range(HISTO_LENGTH)]
    for i in range(len(kl)):
        if not flags[i] & 1:
            continue
        u, v, invzc = (F32(t) for t in proj[i])
        if invzc < 0:
            continue
        if u < F32(bounds[0]) or u > F32(bounds[1]) or v < F32(bounds[2]) or v > F32(bounds[3]):
            continue
        o = int(kl["octave"][i])
        radius = F32(F32(th) * F32(scale_factors[o]))
        if mode == 1:
            cand = py_features_in_area(kc, grid, gp, u, v, radius, o, -1)
        elif mode == 2:
            cand = py_features_in_area(kc, grid, gp, u, v, radius, 0, o)
        else:
            cand = py_features_in_area(kc, grid, gp, u, v, radius, o - 1, o + 1)
        if not cand:
            continue
        best, bidx = 256, -1
        for i2 in cand:
            if occ[i2]:
                continue
            if ur_c[i2] > 0:
                # u - mbf*invzc, fused under the reference's -O3 -march=native: one rounding of the exact value
                ur = F32(np.float64(u) - np.float64(F32(mbf)) * np.float64(invzc))
                if abs(F32(ur - F32(ur_c[i2]))) > radius:
                    continue
            d = py_distance(desc_mp[i], dc[i2])
            if d < best:
                best, bidx = d, i2
        if best <= TH_HIGH:
            assigned[bidx] = i
            occ[bidx] = bool(flags[i] & 2)
            n += 1
            if check_orientation:
                hist[_bin(klu["angle"][i], kc["angle"][bidx])].append(bidx)
    if check_orientation:
        keep = py_three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for i2 in hist[b]:
                    assigned[i2] = -1
                    n -= 1
    return n, np.array(assigned, np.int32)


# ------------------------------------------------------------------ SearchForTriangulation (src/ORBmatcher.cc:359-462, 128-144)
def _round_f32(fr):
    """Fraction -> nearest float32 (ties to even), without the double rounding of float32(float(fr))."""
    from fractions import Fraction
    f = F32(float(fr))
    best = None
    for c in (np.nextafter(f, F32(-np.inf)), f, np.nextafter(f, F32(np.inf))):
        err = abs(Fraction(float(c)) - fr)
        even = (int(np.float32(c).view(np.uint32)) & 1) == 0
        key = (err, 0 if even else 1)
        if best is None or key < best[0]:
            best = (key, c)
    return F32(best[1])


def _fmaf(a, b, c):
    from fractions import Fraction
    return _round_f32(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))


def _fma(a, b, c):
    from fractions import Fraction
    return float(Fraction(float(a)) * Fraction(float(b)) + Fraction(float(c)))  # Fraction -> float rounds to nearest even


def py_check_dist_epipolar_line(x1, y1, x2, y2, F12, sigma2):
    """ORBmatcher::CheckDistEpipolarLine as the reference build evaluates it: a, b, c in double with the first product fused
    (fma(x1, F(0,j), y1 * F(1,j)) + F(2,j)), num / den in float with the first product fused."""
    F12 = np.asarray(F12, np.float64).reshape(3, 3)
    x1, y1, x2, y2 = F32(x1), F32(y1), F32(x2), F32(y2)
    abc = [F32(_fma(float(x1), F12[0, j], float(y1) * F12[1, j]) + F12[2, j]) for j in range(3)]
    a, b, c = abc
    num = F32(_fmaf(a, x2, F32(b * y2)) + c)
    den = _fmaf(a, a, F32(b * b))
    if den == 0:
        return False
    dsqr = F32(F32(num * num) / den)
    return float(dsqr) < 3.84 * float(F32(sigma2))


def py_search_for_triangulation(k1, d1, mp1, ur1, k2, d2, mp2, ur2, F12, ex, ey, sf, sigma2, check_orientation):
    """ORBmatcher::SearchForTriangulation from the epipole on: (nmatches, vMatches12)."""
    n = 0
    m12 = [-1] * len(k1)
    matched2 = [False] * len(k2)  # never set by this reference
    hist = [[] for _ in range(HISTO_LENGTH)]
    ex, ey = F32(ex), F32(ey)
    for i1 in range(len(k1)):
        if mp1[i1]:
            continue
        stereo1 = ur1[i1] >= 0
        best, bidx = TH_LOW, -1
        for i2 in range(len(k2)):
            if matched2[i2] or mp2[i2]:
                continue
            stereo2 = ur2[i2] >= 0
            if not py_check_dist_epipolar_line(k1["x"][i1], k1["y"][i1], k2["x"][i2], k2["y"][i2], F12, sigma2[int(k2["octave"][i2])]):
                continue
            d = py_distance(d1[i1], d2[i2])
            if d > TH_LOW or d > best:
                continue
            if not stereo1 and not stereo2:
                dx, dy = F32(ex - F32(k2["x"][i2])), F32(ey - F32(k2["y"][i2]))
                if _fmaf(dx, dx, F32(dy * dy)) < F32(F32(100) * F32(sf[int(k2["octave"][i2])])):
                    continue
            bidx, best = i2, d
        if bidx >= 0:
            m12[i1] = bidx
            n += 1
            if check_orientation:
                hist[_bin(k1["angle"][i1], k2["angle"][bidx])].append(i1)
    if check_orientation:
        keep = py_three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                m12[i1] = -1
                n -= 1
    return n, np.array(m12, np.int32)


def triangulation_case(seed, n1=300, n2=320, width=640, height=480, nlevels=8, dup=0.1):
    """Two keyframes related by an epipolar geometry F12: the second frame's copies of frame-1 keypoints are put on (or a
    pixel or two off) their epipolar lines; some keypoints already have map points, some are stereo; the epipole lies inside
    the image so that the mono / mono epipole test bites."""
    rng = np.random.default_rng(seed + 500)
    k1, d1, k2, d2 = frame_pair(seed + 300, n1, n2, width, height, nlevels, dup=dup, level0=0.4)
    F12 = np.array([[1.1e-6, 2.3e-5, -0.011], [-2.1e-5, 0.9e-6, 0.023], [0.0093, -0.031, 1.0]], np.float64)
    F12 = F12 * rng.uniform(0.5, 2.0) + rng.normal(0, 1e-7, (3, 3))
    m = min(n1, n2)
    src = rng.permutation(n1)[:m]
    dst = rng.permutation(n2)[:m]
    for s_, t_ in zip(src, dst):
        x1, y1 = float(k1["x"][s_]), float(k1["y"][s_])
        a, b, c = (x1 * F12[0, j] + y1 * F12[1, j] + F12[2, j] for j in range(3))
        x2 = rng.uniform(0, width)
        y2 = -(a * x2 + c) / b if abs(b) > 1e-12 else rng.uniform(0, height)
        off = rng.choice([0.0, 0.0, 0.7, 1.5, 2.5, 6.0]) * rng.choice([-1, 1]) * (1.2 ** int(k2["octave"][t_]))
        k2["x"][t_], k2["y"][t_] = F32(x2), F32(y2 + off)
        dd = d1[s_].copy()
        for bit in rng.choice(256, int(rng.integers(0, 30)), replace=False):
            dd[bit >> 3] ^= 1 << (bit & 7)
        d2[t_] = dd
    mp1 = (rng.random(n1) < 0.3).astype(np.uint8)
    mp2 = (rng.random(n2) < 0.3).astype(np.uint8)
    ur1 = np.where(rng.random(n1) < 0.3, k1["x"] - 5, -1).astype(F32)
    ur2 = np.where(rng.random(n2) < 0.3, k2["x"] - 5, -1).astype(F32)
    ex, ey = F32(rng.uniform(100, width - 100)), F32(rng.uniform(100, height - 100))
    sf = (F32(1.2) ** np.arange(nlevels)).astype(F32)
    return k1, d1, mp1, ur1, k2, d2, mp2, ur2, F12, ex, ey, sf, (sf * sf).astype(F32)


# ------------------------------------------------------------------ SearchByProjection(Frame, map points) (src/ORBmatcher.cc:43-126)
def map_point_inputs(seed, kf, df, n_mp=400, nlevels=8):
    """Map points of the local map as Tracking::SearchLocalPoints hands them over: most project near a keypoint of the frame
    (descriptor a few bits away, predicted level the keypoint's or one above), some are not in view, some have no observations."""
    rng = np.random.default_rng(seed + 900)
    nF = len(kf)
    proj = np.zeros((n_mp, 3), F32)
    level = np.zeros(n_mp, np.int32)
    desc_mp = rng.integers(0, 256, (n_mp, 32), dtype=np.uint8)
    if nF:
        src = rng.integers(0, nF, n_mp)
        proj[:, 0] = kf["x"][src] + rng.uniform(-3, 3, n_mp).astype(F32)
        proj[:, 1] = kf["y"][src] + rng.uniform(-3, 3, n_mp).astype(F32)
        proj[:, 2] = proj[:, 0] - rng.uniform(0, 30, n_mp).astype(F32)
        level[:] = np.clip(kf["octave"][src] + rng.integers(0, 2, n_mp), 0, nlevels - 1)
        dd = df[src].copy()
        for r in range(n_mp):
            for b in rng.choice(256, int(rng.integers(0, 40)), replace=False):
                dd[r, b >> 3] ^= 1 << (b & 7)
        desc_mp = dd
    else:
        proj[:, 0], proj[:, 1] = rng.uniform(0, 640, n_mp), rng.uniform(0, 480, n_mp)
        level[:] = rng.integers(0, nlevels, n_mp)
    view_cos = np.where(rng.random(n_mp) < 0.5, rng.uniform(0.9981, 1.0, n_mp), rng.uniform(0.5, 0.998, n_mp)).astype(F32)
    view_cos[:4] = [0.998, np.nextafter(F32(0.998), F32(1)), np.nextafter(F32(0.998), F32(0)), 1.0][:min(4, n_mp)] if n_mp >= 4 else view_cos[:4]
    flags = (rng.random(n_mp) < 0.85).astype(np.uint8) | ((rng.random(n_mp) < 0.7).astype(np.uint8) << 1)
    return proj, view_cos, level, flags, desc_mp


def py_search_map_points(proj, view_cos, level, flags, desc_mp, kf, df, ur, occ_in, gp, sf, th, nnratio):
    """ORBmatcher::SearchByProjection(Frame, vpMapPoints, th): (nmatches, assigned)."""
    grid = py_grid(kf, gp)
    n = 0
    assigned = [-1] * len(kf)
    occ = [bool(v) for v in occ_in]
    factor = float(F32(th)) != 1.0
    for i in range(len(proj)):
        if not flags[i] & 1:
            continue
        lvl = int(level[i])
        r = F32(2.5) if float(view_cos[i]) > 0.998 else F32(4.0)
        if factor:
            r = F32(r * F32(th))
        radius = F32(r * F32(sf[lvl]))
        x, y, xr = (F32(v) for v in proj[i])
        cand = py_features_in_area(kf, grid, gp, x, y, radius, lvl - 1, lvl)
        if not cand:
            continue
        best = best2 = 256
        blevel = blevel2 = bidx = -1
        for idx in cand:
            if occ[idx]:
                continue
            if ur[idx] > 0 and abs(F32(xr - F32(ur[idx]))) > radius:
                continue
            d = py_distance(desc_mp[i], df[idx])
            if d < best:
                best2, best, blevel2, blevel, bidx = best, d, blevel, int(kf["octave"][idx]), idx
            elif d < best2:
                blevel2, best2 = int(kf["octave"][idx]), d
        if best <= TH_HIGH:
            if blevel == blevel2 and F32(best) > F32(F32(nnratio) * F32(best2)):
                continue
            assigned[bidx] = i
            occ[bidx] = bool(flags[i] & 2)
            n += 1
    return n, np.array(assigned, np.int32)


# ------------------------------------------------------------------ SearchByPoints (src/ORBmatcher.cc:1209-1304)
def py_search_by_points(k1, d1, v1, k2, d2, v2, nnratio, check_orientation):
    """ORBmatcher::SearchByPoints: (nmatches, matches12)."""
    n = 0
    hist = [[] for _ in range(HISTO_LENGTH)]
    m12 = [-1] * len(k1)
    matched2 = [False] * len(k2)
    for i1 in range(len(k1)):
        if not v1[i1]:
            continue
        best1 = best2 = 256
        bidx = -1
        for i2 in range(len(k2)):
            if not v2[i2] or matched2[i2]:
                continue
            d = py_distance(d1[i1], d2[i2])
            if d < best1:
                best2, best1, bidx = best1, d, i2
            elif d < best2:
                best2 = d
        if best1 < TH_LOW and F32(best1) < F32(F32(nnratio) * F32(best2)):
            m12[i1] = bidx
            matched2[bidx] = True
            if check_orientation:
                hist[_bin(k1["angle"][i1], k2["angle"][bidx])].append(i1)
            n += 1
    if check_orientation:
        keep = py_three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b in keep:
                continue
            for i1 in hist[b]:
                m12[i1] = -1
                n -= 1
    return n, np.array(m12, np.int32)


# ------------------------------------------------------------------ SearchByProjection(Frame, KeyFrame, sAlreadyFound, th, ORBdist)
def py_search_by_projection_kf(kkf_un, proj_uv, valid, pred_level, desc_mp, kc, dc, has_mp_cur, gp, sf, bounds, th, orb_dist,
                               check_orientation):
    """src/ORBmatcher.cc:1306-1421 from the projection on: valid[i] = pMP && !isBad() && !sAlreadyFound.count(pMP) && the depth test
    of :1345-1350; pred_level = pMP->PredictScale(dist3D, &CurrentFrame).  (nmatches, assigned)."""
    grid = py_grid(kc, gp)
    n = 0
    assigned = [-1] * len(kc)
    has_mp = [bool(v) for v in has_mp_cur]
    hist = [[] for _ in range(HISTO_LENGTH)]
    for i in range(len(kkf_un)):
        if not valid[i]:
            continue
        u, v = F32(proj_uv[i][0]), F32(proj_uv[i][1])
        if u < F32(bounds[0]) or u > F32(bounds[1]) or v < F32(bounds[2]) or v > F32(bounds[3]):
            continue
        lvl = int(pred_level[i])
        radius = F32(F32(th) * F32(sf[lvl]))
        cand = py_features_in_area(kc, grid, gp, u, v, radius, lvl - 1, lvl + 1)
        if not cand:
            continue
        best, bidx = 256, -1
        for i2 in cand:
            if has_mp[i2]:
                continue
            d = py_distance(desc_mp[i], dc[i2])
            if d < best:
                best, bidx = d, i2
        if best <= orb_dist:
            assigned[bidx] = i
            has_mp[bidx] = True
            n += 1
            if check_orientation:
                hist[_bin(kkf_un["angle"][i], kc["angle"][bidx])].append(bidx)
    if check_orientation:
        keep = py_three_maxima([len(h) for h in hist])
        for b in range(HISTO_LENGTH):
            if b not in keep:
                for i2 in hist[b]:
                    assigned[i2] = -1
                    n -= 1
    return n, np.array(assigned, np.int32)


def kf_projection_args(seed, nk, nc):
    """Inputs of the KeyFrame overload and their mapping onto the Frame / Frame entry point (include/sdorb.h,
    sdorb_projection_search::orb_dist): octave of kps_last = predicted level, flags = valid | 2, proj[2] = 1, mvuRight = -1."""
    kk, dk, kc, dc = frame_pair(seed + 120, nk, nc, jitter=4.0, level0=0.3, flips=30)
    rng = np.random.default_rng(seed + 7)
    proj = np.zeros((nk, 3), F32)
    proj[:, 0] = kk["x"] + rng.uniform(-3, 3, nk).astype(F32)
    proj[:, 1] = kk["y"] + rng.uniform(-3, 3, nk).astype(F32)
    proj[:, 2] = 1
    valid = (rng.random(nk) < 0.8).astype(np.uint8)
    pred = np.clip(kk["octave"] + rng.integers(-1, 2, nk), 0, 7).astype(np.int32)
    has_mp = (rng.random(nc) < 0.15).astype(np.uint8)
    k_level = kk.copy()
    k_level["octave"] = pred
    return dict(kk=kk, proj=proj, valid=valid, pred=pred, dmp=dk, kc=kc, dc=dc, has_mp=has_mp, k_level=k_level,
                flags=(valid | 2).astype(np.uint8), ur=np.full(nc, -1, F32))


# ------------------------------------------------------------------ Fuse: the keypoint search (src/ORBmatcher.cc:535-586)
def fuse_inputs(seed, kf, n_mp=400, nlevels=8, stereo=True):
    """Map points as ORBmatcher::Fuse sees them after its visibility checks: most project close to a keypoint of the keyframe
    (some right on the chi-square limit), predicted level the keypoint's or one above / below, some fail the checks (flag 0)."""
    rng = np.random.default_rng(seed + 1300)
    nF = len(kf)
    proj = np.zeros((n_mp, 3), F32)
    level = rng.integers(0, nlevels, n_mp).astype(np.int32)
    ur = np.full(nF, -1, F32)
    if nF:
        ur_all = np.where(rng.random(nF) < 0.6, kf["x"] - rng.uniform(0, 30, nF), -1).astype(F32)
        if stereo:
            ur = ur_all
        src = rng.integers(0, nF, n_mp)
        spread = rng.choice([0.3, 1.5, 4.0], n_mp)
        proj[:, 0] = kf["x"][src] + (rng.uniform(-1, 1, n_mp) * spread).astype(F32)
        proj[:, 1] = kf["y"][src] + (rng.uniform(-1, 1, n_mp) * spread).astype(F32)
        proj[:, 2] = np.where(ur[src] >= 0, ur[src] + rng.uniform(-1.5, 1.5, n_mp), proj[:, 0] - 10).astype(F32)
        level[:] = np.clip(kf["octave"][src] + rng.integers(-1, 2, n_mp), 0, nlevels - 1)
    else:
        proj[:, 0], proj[:, 1] = rng.uniform(0, 640, n_mp), rng.uniform(0, 480, n_mp)
    flags = (rng.random(n_mp) < 0.85).astype(np.uint8)
    return proj, level, flags, ur


def fuse_descriptors(seed, df, kf_src_n, n_mp):
    """Map-point descriptors: the descriptor of the keypoint the point was projected next to (same draw of src as fuse_inputs)
    with 0..70 bits flipped, so distances straddle TH_LOW = 50."""
    rng = np.random.default_rng(seed + 1400)
    if kf_src_n == 0:
        return rng.integers(0, 256, (n_mp, 32), dtype=np.uint8)
    r0 = np.random.default_rng(seed + 1300)
    r0.integers(0, 8, n_mp)
    r0.random(kf_src_n), r0.uniform(0, 30, kf_src_n)
    dd = df[r0.integers(0, kf_src_n, n_mp)].copy()
    for r in range(n_mp):
        for b in rng.choice(256, int(rng.integers(0, 71)), replace=False):
            dd[r, b >> 3] ^= 1 << (b & 7)
    return dd


def py_fuse_search(proj, level, flags, desc_mp, kf, df, ur, gp, sf, inv_sigma2, th, check_reprojection=True, th_dist=TH_LOW):
    """(best_idx, best_dist) per map point; e2 in the contracted form of the reference build (see
    test_fuse_e2_matches_reference_flags): fma(er, er, fma(ex, ex, ey * ey)).  check_reprojection=False: the search of the Sim3
    overload (src/ORBmatcher.cc:682-708) and, with th_dist = TH_HIGH, of either direction of SearchBySim3 (:812-845, :890-923)."""
    grid = py_grid(kf, gp)
    bi, bd = [-1] * len(proj), [256] * len(proj)
    for i in range(len(proj)):
        if not flags[i] & 1:
            continue
        lvl = int(level[i])
        u, v, r = (F32(x) for x in proj[i])
        radius = F32(F32(th) * F32(sf[lvl]))
        best, bidx = 256, -1
        for idx in py_features_in_area(kf, grid, gp, u, v, radius, -1, -1):
            kl = int(kf["octave"][idx])
            if kl < lvl - 1 or kl > lvl:
                continue
            if check_reprojection:
                ex, ey = F32(u - F32(kf["x"][idx])), F32(v - F32(kf["y"][idx]))
                e2 = _fmaf(ex, ex, F32(ey * ey))
                lim = 5.99
                if ur[idx] >= 0:
                    er = F32(r - F32(ur[idx]))
                    e2 = _fmaf(er, er, e2)
                    lim = 7.8
                if float(F32(e2 * F32(inv_sigma2[kl]))) > lim:
                    continue
            d = py_distance(desc_mp[i], df[idx])
            if d < best:
                best, bidx = d, idx
        bd[i] = best
        if best <= th_dist:
            bi[i] = bidx
    return np.array(bi, np.int32), np.array(bd, np.int32)


# ------------------------------------------------------------------ SearchBySim3 (src/ORBmatcher.cc:734-944)
def sim3_case(seed, n1, n2, nlevels=8, dup=0.0):
    """Two keyframes seeing the same scene: keypoint i of either keyframe may carry a map point (flag), which projects into the
    other keyframe next to its true partner (or somewhere else), with the map point's own descriptor a few bits off."""
    k1, d1, k2, d2 = frame_pair(seed + 160, n1, n2, jitter=5.0, level0=0.3, flips=30, dup=dup)
    rng = np.random.default_rng(seed + 1600)

    def side(ka, da, kb, db):
        na, nb = len(ka), len(kb)
        proj = np.zeros((na, 3), F32)
        level = rng.integers(0, nlevels, na).astype(np.int32)
        dmp = da.copy()
        if na and nb:
            src = np.where(rng.random(na) < 0.8, np.arange(na) % nb, rng.integers(0, nb, na))
            proj[:, 0] = kb["x"][src] + rng.uniform(-4, 4, na).astype(F32)
            proj[:, 1] = kb["y"][src] + rng.uniform(-4, 4, na).astype(F32)
            level[:] = np.clip(kb["octave"][src] + rng.integers(0, 2, na), 0, nlevels - 1)
            dmp = db[src].copy()
            for r in range(na):
                for b in rng.choice(256, int(rng.integers(0, 60)), replace=False):
                    dmp[r, b >> 3] ^= 1 << (b & 7)
        elif na:
            proj[:, 0], proj[:, 1] = rng.uniform(0, 640, na), rng.uniform(0, 480, na)
        flags = (rng.random(na) < 0.8).astype(np.uint8)
        return proj, level, flags, dmp
    return (side(k1, d1, k2, d2) + (k1, d1)), (side(k2, d2, k1, d1) + (k2, d2))


def py_search_by_sim3(s1, s2, gp, sf, th):
    """(nFound, matches12, vnMatch1, vnMatch2); s = (proj, level, flags, desc_mp, kps, desc) of a keyframe."""
    def one_way(s, kb, db):
        proj, level, flags, dmp = s[:4]
        grid = py_grid(kb, gp)
        out = [-1] * len(proj)
        for i in range(len(proj)):
            if not flags[i] & 1:
                continue
            lvl = int(level[i])
            radius = F32(F32(th) * F32(sf[lvl]))
            best, bidx = INT_MAX, -1
            for idx in py_features_in_area(kb, grid, gp, F32(proj[i][0]), F32(proj[i][1]), radius, -1, -1):
                if kb["octave"][idx] < lvl - 1 or kb["octave"][idx] > lvl:
                    continue
                d = py_distance(dmp[i], db[idx])
                if d < best:
                    best, bidx = d, idx
            if best <= TH_HIGH:
                out[i] = bidx
        return out
    m1, m2 = one_way(s1, s2[4], s2[5]), one_way(s2, s1[4], s1[5])
    m12 = [-1] * len(m1)
    n = 0
    for i1, idx2 in enumerate(m1):
        if idx2 >= 0 and m2[idx2] == i1:
            m12[i1] = idx2
            n += 1
    return n, np.array(m12, np.int32), np.array(m1, np.int32), np.array(m2, np.int32)


# ------------------------------------------------------------------ SearchByProjection(KeyFrame, Scw, vpPoints, vpMatched, th) (:146-254)
def py_search_by_projection_sim3(proj, level, flags, desc_mp, kf, df, matched_in, gp, sf, th):
    """(nmatches, assigned): assigned[idx] = index of the point written to vpMatched[idx]."""
    grid = py_grid(kf, gp)
    matched = [bool(m) for m in matched_in]
    assigned = [-1] * len(kf)
    n = 0
    for i in range(len(proj)):
        if not flags[i] & 1:
            continue
        lvl = int(level[i])
        radius = F32(F32(int(th)) * F32(sf[lvl]))
        best, bidx = 256, -1
        for idx in py_features_in_area(kf, grid, gp, F32(proj[i][0]), F32(proj[i][1]), radius, -1, -1):
            if matched[idx]:
                continue
            if kf["octave"][idx] < lvl - 1 or kf["octave"][idx] > lvl:
                continue
            d = py_distance(desc_mp[i], df[idx])
            if d < best:
                best, bidx = d, idx
        if best <= TH_LOW:
            matched[bidx] = True
            assigned[bidx] = i
            n += 1
    return n, np.array(assigned, np.int32)
