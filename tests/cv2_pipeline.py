"""Independent second implementation of ORBextractor::operator() assembled from Python cv2 4.13 calls.

Used only to pin the C++ oracle (and to generate tests/golden/*.npz): every pixel operation here is the
real OpenCV one (cv2.resize, cv2.FastFeatureDetector on the cell ROI, cv2.GaussianBlur, cv2.fastAtan2);
the control flow follows /root/reference/src/ORBextractor.cc:466-700 and is written separately from
oracle/sdorb_oracle.cc.  retainBest has no Python binding; its ordering is pinned separately against
cv2.ORB (tests/test_oracle_primitives.py::test_retain_best_matches_cv2_orb) and then reused from the oracle.
"""
import ctypes
import math

import cv2
import numpy as np

from oracle import binding as orc

_libm = ctypes.CDLL("libm.so.6")
_libm.sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_float)]
f32 = np.float32
EDGE = 19


def _pattern():
    txt = open(orc._HERE + "/orb_pattern.inc").read()
    vals = [int(v) for line in txt.splitlines() if not line.startswith("//") for v in line.split(",") if v.strip()]
    return np.array(vals, np.int32).reshape(512, 2)


def cv_round(v):
    """cvRound: round-half-to-even."""
    return int(np.rint(v))


def tables(nfeatures, scale_factor, nlevels):
    sfd = float(f32(scale_factor))  # member `double scaleFactor` initialised from a float
    sf = [f32(1.0)]
    for _ in range(1, nlevels):
        sf.append(f32(float(sf[-1]) * sfd))
    inv = [f32(1.0) / s for s in sf]
    factor = f32(1.0 / sfd)
    ndes = f32(nfeatures) * (f32(1) - factor) / (f32(1) - f32(math.pow(float(factor), float(nlevels))))
    npl, tot = [], 0
    for _ in range(nlevels - 1):
        npl.append(cv_round(ndes))
        tot += npl[-1]
        ndes = f32(ndes * factor)
    npl.append(max(nfeatures - tot, 0))
    umax = [0] * 16
    vmax = int(math.floor(f32(15) * f32(math.sqrt(2.0)) / f32(2) + f32(1)))
    vmin = int(math.ceil(f32(15) * f32(math.sqrt(2.0)) / f32(2)))
    for v in range(vmax + 1):
        umax[v] = cv_round(math.sqrt(225.0 - v * v))
    v0 = 0
    for v in range(15, vmin - 1, -1):
        while umax[v0] == umax[v0 + 1]:
            v0 += 1
        umax[v] = v0
        v0 += 1
    return sf, inv, npl, umax


def octree_cells(level, ini_th, min_th):
    """ORB-SLAM2 ComputeKeyPointsOctTree, the cell loop: real cv2 FAST on every 30-pixel cell ROI with iniThFAST, with
    minThFAST where that finds nothing.  Returns vToDistributeKeys as (x, y, response) relative to (minBorderX, minBorderY)."""
    lh, lw = level.shape
    det = [cv2.FastFeatureDetector_create(threshold=t, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
           for t in (ini_th, min_th)]
    min_bx = min_by = EDGE - 3
    max_bx, max_by = lw - EDGE + 3, lh - EDGE + 3
    width, height = f32(max_bx - min_bx), f32(max_by - min_by)
    ncols, nrows = int(width / f32(30)), int(height / f32(30))
    wcell, hcell = int(np.ceil(width / f32(ncols))), int(np.ceil(height / f32(nrows)))
    keys = []
    for i in range(nrows):
        ini_y = min_by + i * hcell
        if ini_y >= max_by - 3:
            continue
        max_y = min(ini_y + hcell + 6, max_by)
        for j in range(ncols):
            ini_x = min_bx + j * wcell
            if ini_x >= max_bx - 6:
                continue
            max_x = min(ini_x + wcell + 6, max_bx)
            roi = level[ini_y:max_y, ini_x:max_x]
            found = det[0].detect(roi)
            if not found:
                found = det[1].detect(roi)
            keys += [(k.pt[0] + j * wcell, k.pt[1] + i * hcell, k.response) for k in found]
    return keys, (min_bx, max_bx, min_by, max_by)


def distribute_oct_tree(keys, min_x, max_x, min_y, max_y, n_wanted):
    """ORB-SLAM2 DistributeOctTree written as whole-list passes over arrays instead of std::list surgery (the form the CUDA
    kernel uses): the list is always in descending creation order because every node enters at the front; a node is
    (x0, x1, y0, y1, [key indices in input order]).  Equal sizes in the (size, pointer) sort: later-created node first."""
    if not keys:
        return []
    n_ini = int(math.floor(float(f32(max_x - min_x) / f32(max_y - min_y)) + 0.5))
    hx = f32(max_x - min_x) / f32(n_ini)
    ini = [[int(hx * f32(i)), int(hx * f32(i + 1)), 0, max_y - min_y, []] for i in range(n_ini)]
    for k, (x, y, r) in enumerate(keys):
        ini[int(f32(x) / hx)][4].append(k)
    nodes = [nd for nd in ini if nd[4]]  # front .. back

    def divide(nd):
        x0, x1, y0, y1, ks = nd
        sx = x0 + int(math.ceil(float(f32(x1 - x0) / f32(2))))
        sy = y0 + int(math.ceil(float(f32(y1 - y0) / f32(2))))
        ch = [[x0, sx, y0, sy, []], [sx, x1, y0, sy, []], [x0, sx, sy, y1, []], [sx, x1, sy, y1, []]]
        for k in ks:
            x, y, _ = keys[k]
            ch[(0 if x < sx else 1) + (0 if y < sy else 2)][4].append(k)
        return [c for c in ch if c[4]]

    sorted_mode = False
    while True:
        prev = len(nodes)
        cand = [i for i, nd in enumerate(nodes) if len(nd[4]) > 1]  # list order = newest first
        if sorted_mode:
            cand.sort(key=lambda i: (-len(nodes[i][4]), i))  # biggest first; equal sizes: newest (= smallest position) first
        created, done, size = [], set(), len(nodes)
        for i in cand:
            ch = divide(nodes[i])
            created += ch
            done.add(i)
            size += len(ch) - 1
            if sorted_mode and size >= n_wanted:
                break
        n_expand = sum(1 for c in created if len(c[4]) > 1)
        nodes = created[::-1] + [nd for i, nd in enumerate(nodes) if i not in done]
        if len(nodes) >= n_wanted or len(nodes) == prev:
            break
        if not sorted_mode and len(nodes) + 3 * n_expand > n_wanted:
            sorted_mode = True
    out = []
    for nd in nodes:
        best = nd[4][0]
        for k in nd[4][1:]:
            if keys[k][2] > keys[best][2]:
                best = k
        out.append(keys[best])
    return out


def extract(img, nfeatures=1000, scale_factor=1.2, nlevels=8, th_fast=20, min_th_fast=None):
    """min_th_fast given: the ORB-SLAM2-style mode (th_fast = iniThFAST), SURVEY.md section 8 row f1."""
    assert img.dtype == np.uint8 and img.ndim == 2
    h0, w0 = img.shape
    sf, inv, npl, umax = tables(nfeatures, scale_factor, nlevels)
    pat = _pattern()
    fast = cv2.FastFeatureDetector_create(threshold=th_fast, nonmaxSuppression=True,
                                          type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    # ComputePyramid: chained resize into the ROI of a padded buffer
    pyr = []
    for l in range(nlevels):
        w, h = cv_round(f32(w0) * inv[l]), cv_round(f32(h0) * inv[l])
        inner = img if l == 0 else cv2.resize(pyr[l - 1][EDGE:-EDGE, EDGE:-EDGE], (w, h), interpolation=cv2.INTER_LINEAR)
        padded = cv2.copyMakeBorder(inner, EDGE, EDGE, EDGE, EDGE, cv2.BORDER_REFLECT_101)
        pyr.append(padded)
    ratio = f32(w0) / f32(h0)
    out_k, out_d = [], []
    for l in range(nlevels):
        level = pyr[l][EDGE:-EDGE, EDGE:-EDGE]
        lh, lw = level.shape
        ndes = npl[l]
        cols = int(np.sqrt(f32(ndes) / (f32(5) * ratio)))
        rows = int(ratio * f32(cols))
        kps = []  # (x, y, response)
        if min_th_fast is not None:
            keys, (bx0, bx1, by0, by1) = octree_cells(level, th_fast, min_th_fast)
            kps = [(x + bx0, y + by0, r) for x, y, r in distribute_oct_tree(keys, bx0, bx1, by0, by1, ndes)]
        elif cols > 0 and rows > 0:
            max_bx, max_by = lw - EDGE, lh - EDGE
            W, H = max_bx - EDGE, max_by - EDGE
            cw, ch = int(np.ceil(f32(W) / f32(cols))), int(np.ceil(f32(H) / f32(rows)))
            ncells = rows * cols
            nfc = int(np.ceil(f32(ndes) / f32(ncells)))
            cell = [[[] for _ in range(cols)] for _ in range(rows)]
            n_ret = [[0] * cols for _ in range(rows)]
            n_tot = [[0] * cols for _ in range(rows)]
            no_more = [[False] * cols for _ in range(rows)]
            ini_x, ini_y = [0] * cols, [0] * rows
            n_no_more = n_dist = 0
            hy = ch + 6
            for i in range(rows):
                iy = EDGE + i * ch - 3
                ini_y[i] = iy
                if i == rows - 1:
                    hy = max_by + 3 - iy
                    if hy <= 0:
                        continue
                hx = cw + 6
                for j in range(cols):
                    if i == 0:
                        ini_x[j] = EDGE + j * cw - 3
                    ix = ini_x[j]
                    if j == cols - 1:
                        hx = max_bx + 3 - ix
                        if hx <= 0:
                            continue
                    assert 0 <= iy <= iy + hy <= lh and 0 <= ix <= ix + hx <= lw
                    roi = level[iy:iy + hy, ix:ix + hx]
                    cell[i][j] = [(k.pt[0], k.pt[1], k.response) for k in fast.detect(roi)] if roi.size else []
                    n = len(cell[i][j])
                    n_tot[i][j] = n
                    if n > nfc:
                        n_ret[i][j] = nfc
                    else:
                        n_ret[i][j] = n
                        n_dist += nfc - n
                        no_more[i][j] = True
                        n_no_more += 1
            while n_dist > 0 and n_no_more < ncells:
                n_new = nfc + int(np.ceil(f32(n_dist) / f32(ncells - n_no_more)))
                n_dist = 0
                for i in range(rows):
                    for j in range(cols):
                        if not no_more[i][j]:
                            if n_tot[i][j] > n_new:
                                n_ret[i][j] = n_new
                            else:
                                n_ret[i][j] = n_tot[i][j]
                                n_dist += n_new - n_tot[i][j]
                                no_more[i][j] = True
                                n_no_more += 1
            for i in range(rows):
                for j in range(cols):
                    c = cell[i][j]
                    if len(c) > n_ret[i][j] >= 0:
                        if n_ret[i][j] == 0:
                            c = []
                        else:
                            order = orc.retain_best_order([r for _, _, r in c], n_ret[i][j])
                            c = [c[o] for o in order][:n_ret[i][j]]
                    kps += [(x + ini_x[j], y + ini_y[i], r) for x, y, r in c]
            if len(kps) > ndes:
                order = orc.retain_best_order([r for _, _, r in kps], ndes)
                kps = [kps[o] for o in order][:ndes]
        if not kps:
            continue
        blurred = cv2.GaussianBlur(level.copy(), (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        size = f32(int(f32(31) * sf[l]))
        for x, y, r in kps:
            xi, yi = int(x), int(y)
            m01 = m10 = 0
            for u in range(-15, 16):
                m10 += u * int(level[yi, xi + u])
            for v in range(1, 16):
                d = umax[v]
                vs = 0
                for u in range(-d, d + 1):
                    p, m = int(level[yi + v, xi + u]), int(level[yi - v, xi + u])
                    vs += p - m
                    m10 += u * (p + m)
                m01 += v * vs
            angle = f32(cv2.fastAtan2(float(m01), float(m10)))
            s, c = ctypes.c_float(), ctypes.c_float()
            _libm.sincosf(f32(angle * f32(math.pi / 180.0)), ctypes.byref(s), ctypes.byref(c))
            a, b = f32(c.value), f32(s.value)
            desc = np.zeros(32, np.uint8)
            for i in range(256):
                vals = []
                for px, py in (pat[2 * i], pat[2 * i + 1]):
                    # fma(px, b, py*a), fma(px, a, -(py*b)) evaluated exactly: float64 holds the exact
                    # product px*b (<= 24+4 bits) and the sum before the single rounding to float32
                    row = cv_round(f32(float(f32(px)) * float(b) + float(f32(f32(py) * a))))
                    col = cv_round(f32(float(f32(px)) * float(a) - float(f32(f32(py) * b))))
                    vals.append(int(blurred[yi + row, xi + col]))
                desc[i // 8] |= (vals[0] < vals[1]) << (i % 8)
            if l:
                x, y = f32(x) * sf[l], f32(y) * sf[l]
            out_k.append((x, y, size, angle, r, l, -1))
            out_d.append(desc)
    k = np.array(out_k, orc.KP_DTYPE) if out_k else np.zeros(0, orc.KP_DTYPE)
    d = np.array(out_d, np.uint8).reshape(-1, 32)
    inner_pyr = [p[EDGE:-EDGE, EDGE:-EDGE].copy() for p in pyr]
    return k, d, inner_pyr
