"""ctypes binding of the CPU oracle (oracle/libsdorb_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs as the checker / reported CPU baseline.  The product package (sdslam_b200) never
imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libsdorb_oracle.so")

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
MATCH_DTYPE = np.dtype([("best_idx", "<i4"), ("best_dist", "<i4"), ("second_dist", "<i4"), ("accepted", "<i4")])
GEOM_DTYPE = np.dtype([(n, "<i4") for n in ("width", "height", "n_desired", "level_cols", "level_rows", "cell_w",
                                             "cell_h", "n_features_cell", "scaled_patch_size")])
assert KP_DTYPE.itemsize == 28


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("th_fast", C.c_int)]


class Dump(C.Structure):
    _fields_ = [("pyramid", C.c_void_p), ("blurred", C.c_void_p), ("level_count", C.c_void_p),
                ("raw_cell_count", C.c_void_p), ("raw_cell_cap", C.c_int32), ("raw_kps", C.c_void_p),
                ("raw_kps_cap", C.c_int32), ("raw_kps_total", C.c_int32), ("n_to_retain_cap", C.c_int32),
                ("n_to_retain", C.c_void_p)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("sdorb_oracle.cc", "sdorb_oracle.h", "orb_pattern.inc", "Makefile")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


def build_native():
    """The same oracle compiled on THIS machine with the reference's optimisation flags (-O3 -march=native,
    /root/reference/CMakeLists.txt:39-40) for the CPU baseline timing.  -ffp-contract=off stays: every FMA the
    reference build contracts is an explicit fmaf() in the source, so results are unchanged (checked by the caller)."""
    out_dir = os.path.join(_HERE, "_native")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libsdorb_oracle_native.so")
    subprocess.check_call(["g++", "-O3", "-march=native", "-std=c++17", "-fPIC", "-ffp-contract=off", "-fno-fast-math",
                           "-pthread", "-shared", "-o", so, os.path.join(_HERE, "sdorb_oracle.cc"), "-lm"])
    return so


def use_native():
    """Switch this module to the -march=native build (bench.py CPU legs).  Returns True when it is in use."""
    global _lib, _SO
    try:
        so = build_native()
    except Exception:
        return False
    _SO = so
    _lib = None
    lib(path=so)
    return True


class FrameGrid(C.Structure):
    _fields_ = [("cell_start", C.c_void_p), ("indices", C.c_void_p), ("mnMinX", C.c_float), ("mnMinY", C.c_float),
                ("mfGridElementWidthInv", C.c_float), ("mfGridElementHeightInv", C.c_float)]


_lib = None


def lib(path=None):
    global _lib
    if _lib is None:
        _lib = C.CDLL(path or build())
        L = _lib
        vp, i, sz, f = C.c_void_p, C.c_int, C.c_size_t, C.c_float
        L.orc_resize_linear_8u.argtypes = [vp, i, i, sz, vp, i, i, sz]
        L.orc_copy_make_border_reflect101.argtypes = [vp, i, i, sz, vp, sz, i]
        L.orc_fast.argtypes = [vp, i, i, sz, i, i, vp, i]
        L.orc_fast.restype = i
        L.orc_fast_score_map.argtypes = [vp, i, i, sz, i, vp, sz]
        L.orc_gaussian_blur_7x7_s2.argtypes = [vp, i, i, sz, vp, sz]
        L.orc_fast_atan2.argtypes = [f, f]
        L.orc_fast_atan2.restype = f
        L.orc_retain_best.argtypes = [vp, i, i]
        L.orc_retain_best.restype = i
        L.orc_retain_best_idx.argtypes = [vp, i, i, vp]
        L.orc_retain_best_idx.restype = i
        L.orc_sincosf_restated.argtypes = [f, C.POINTER(f), C.POINTER(f)]
        L.orc_sincosf_mismatches.argtypes = [C.c_uint32, C.c_uint32, i]
        L.orc_sincosf_mismatches.restype = C.c_uint64
        L.orc_pattern_rotate.argtypes = [i, i, f, f, C.POINTER(i), C.POINTER(i)]
        L.orc_create.argtypes = [C.POINTER(Params)]
        L.orc_create.restype = vp
        L.orc_destroy.argtypes = [vp]
        L.orc_get_tables.argtypes = [vp] * 7
        L.orc_level_geometry.argtypes = [vp, i, i, vp]
        L.orc_extract.argtypes = [vp, vp, i, i, sz, vp, vp, i, C.POINTER(Dump)]
        L.orc_extract.restype = i
        L.orc_extract_many.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, i]
        L.orc_extract_many.restype = C.c_long
        L.orc_descriptor_distance.argtypes = [vp, vp]
        L.orc_descriptor_distance.restype = i
        L.orc_match_best2.argtypes = [vp, i, vp, i, f, i, vp]
        L.orc_match_greedy.argtypes = [vp, i, vp, i, f, i, vp]
        L.orc_match_many.argtypes = [vp, vp, vp, vp, i, i, i, f, i, i, vp]
        L.orc_hamming_matrix.argtypes = [vp, i, vp, i, vp]
        L.orc_assign_grid.argtypes = [vp, i, f, f, f, f, vp, vp]
        L.orc_stereo_from_rgbd.argtypes = [vp, vp, i, vp, i, f, vp, vp]
        L.orc_undistort_keypoints.argtypes = [vp, i, vp, vp, i, vp]
        L.orc_image_bounds.argtypes = [i, i, vp, vp, i, vp]
        L.orc_distinctive.argtypes = [vp, i, C.POINTER(i)]
        L.orc_distinctive.restype = i
        L.orc_distinctive_many.argtypes = [vp, vp, i, i, vp, vp]
        L.orc_set_orbslam2_mode.argtypes = [vp, i, i]
        L.orc_set_orbslam2_mode.restype = None
        L.orc_distribute_oct_tree.argtypes = [vp, i, i, i, i, i, i, vp, i]
        L.orc_distribute_oct_tree.restype = i
        L.orc_features_in_area.argtypes = [vp, C.POINTER(FrameGrid), f, f, f, i, i, vp]
        L.orc_features_in_area.restype = i
        L.orc_three_maxima.argtypes = [vp, i, C.POINTER(i), C.POINTER(i), C.POINTER(i)]
        L.orc_search_for_initialization.argtypes = [vp, vp, i, vp, vp, i, C.POINTER(FrameGrid), vp, i, f, i, vp]
        L.orc_search_for_initialization.restype = i
        L.orc_search_by_projection.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, C.POINTER(FrameGrid), vp, vp, f, f, i, i, vp, i]
        L.orc_search_by_projection.restype = i
        L.orc_search_map_points.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, C.POINTER(FrameGrid), vp, f, f, vp]
        L.orc_search_map_points.restype = i
        L.orc_search_by_points.argtypes = [vp, vp, vp, i, vp, vp, vp, i, f, i, vp]
        L.orc_search_by_points.restype = i
        L.orc_fuse_search.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, C.POINTER(FrameGrid), vp, vp, f, i, i, vp, vp]
        L.orc_fuse_search.restype = None
        L.orc_search_by_sim3.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, C.POINTER(FrameGrid), vp, vp, C.POINTER(FrameGrid),
                                         vp, vp, f, vp, vp, vp]
        L.orc_search_by_sim3.restype = i
        L.orc_search_by_projection_sim3.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, i, C.POINTER(FrameGrid), vp, i, vp]
        L.orc_search_by_projection_sim3.restype = i
        L.orc_check_dist_epipolar_line.argtypes = [f, f, f, f, vp, f]
        L.orc_check_dist_epipolar_line.restype = i
        L.orc_search_for_triangulation.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, f, f, vp, vp, i, vp]
        L.orc_search_for_triangulation.restype = i
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _u8(img):
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim == 2 and img.strides[1] == 1
    return img


# ---------------------------------------------------------------- primitives
def resize_linear(img, dw, dh):
    img = _u8(img)
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_8u(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), dw, dh, out.strides[0])
    return out


def border_reflect101(img, b):
    img = _u8(img)
    out = np.empty((img.shape[0] + 2 * b, img.shape[1] + 2 * b), np.uint8)
    lib().orc_copy_make_border_reflect101(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), out.strides[0], b)
    return out


def fast(img, th, nonmax=True):
    img = _u8(img)
    cap = max(1, img.size)
    out = np.zeros(cap, KP_DTYPE)
    n = lib().orc_fast(_p(img), img.shape[1], img.shape[0], img.strides[0], th, int(nonmax), _p(out), cap)
    return out[:n]


def fast_score_map(img, th):
    img = _u8(img)
    out = np.zeros(img.shape, np.uint8)
    lib().orc_fast_score_map(_p(img), img.shape[1], img.shape[0], img.strides[0], th, _p(out), out.strides[0])
    return out


def gaussian_blur(img):
    img = _u8(img)
    out = np.empty(img.shape, np.uint8)
    lib().orc_gaussian_blur_7x7_s2(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(out), out.strides[0])
    return out


def fast_atan2(y, x):
    return lib().orc_fast_atan2(float(y), float(x))


def retain_best_order(responses, n):
    r = np.ascontiguousarray(responses, np.float32)
    order = np.empty(len(r), np.int32)
    m = lib().orc_retain_best_idx(_p(r), len(r), n, _p(order))
    return order[:m]


def sincosf_restated(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincosf_restated(float(x), C.byref(s), C.byref(c))
    return s.value, c.value


def sincosf_mismatches(lo, hi, nthreads=8):
    lo_bits = int(np.float32(lo).view(np.uint32))
    hi_bits = int(np.float32(hi).view(np.uint32))
    return int(lib().orc_sincosf_mismatches(lo_bits, hi_bits, nthreads))


def pattern_rotate(px, py, a, b):
    r, c = C.c_int(), C.c_int()
    lib().orc_pattern_rotate(px, py, float(a), float(b), C.byref(r), C.byref(c))
    return r.value, c.value


# ---------------------------------------------------------------- extractor
class Extractor:
    """Oracle twin of SD_SLAM::ORBextractor(nfeatures, scaleFactor, nlevels, thFAST)."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, th_fast=20, min_th_fast=None):
        """min_th_fast given: the ORB-SLAM2-style mode (th_fast = iniThFAST; 30-pixel cells, DistributeOctTree)."""
        self.params = Params(nfeatures, scale_factor, nlevels, th_fast)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = lib().orc_create(C.byref(self.params))
        if not self._h:
            raise ValueError("bad oracle parameters")
        self.octree = min_th_fast is not None and min_th_fast >= 0
        if self.octree:
            lib().orc_set_orbslam2_mode(self._h, th_fast, min_th_fast)

    def _cap(self):
        # DistributeOctTree stops at >= N nodes per level: up to N + 2 (or the 4 children of each initial node) keypoints
        return max(self.nfeatures, 1) + (8 * self.nlevels if self.octree else 0)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    def tables(self):
        n = self.nlevels
        sf, inv, s2, is2 = (np.empty(n, np.float32) for _ in range(4))
        npl, umax = np.empty(n, np.int32), np.empty(16, np.int32)
        lib().orc_get_tables(self._h, _p(sf), _p(inv), _p(s2), _p(is2), _p(npl), _p(umax))
        return dict(scale=sf, inv_scale=inv, sigma2=s2, inv_sigma2=is2, n_per_level=npl, umax=umax)

    def geometry(self, w, h):
        g = np.zeros(self.nlevels, GEOM_DTYPE)
        lib().orc_level_geometry(self._h, w, h, _p(g))
        return g

    def extract(self, img, dump=False):
        """Returns (kps, desc) or, with dump=True, (kps, desc, stages dict)."""
        img = _u8(img)
        h, w = img.shape
        cap = self._cap()
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        d = None
        keep = {}
        if dump:
            g = self.geometry(w, h)
            npx = int(sum(max(int(a), 0) * max(int(b), 0) for a, b in zip(g["width"], g["height"])))
            ncell = int(sum(max(int(a), 0) * max(int(b), 0) for a, b in zip(g["level_cols"], g["level_rows"])))
            keep = dict(pyramid=np.zeros(npx, np.uint8), blurred=np.zeros(npx, np.uint8),
                        level_count=np.zeros(self.nlevels, np.int32), raw_cell_count=np.zeros(max(ncell, 1), np.int32),
                        raw_kps=np.zeros(max(npx // 4 + 16, 16), KP_DTYPE), n_to_retain=np.zeros(max(ncell, 1), np.int32))
            d = Dump(_p(keep["pyramid"]), _p(keep["blurred"]), _p(keep["level_count"]), _p(keep["raw_cell_count"]),
                     len(keep["raw_cell_count"]), _p(keep["raw_kps"]), len(keep["raw_kps"]), 0,
                     len(keep["n_to_retain"]), _p(keep["n_to_retain"]))
        n = lib().orc_extract(self._h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap,
                              C.byref(d) if d is not None else None)
        if n < 0:
            raise RuntimeError("oracle: cell ROI outside the level image (the reference throws cv::Exception)")
        assert n <= cap
        if not dump:
            return kps[:n], desc[:n]
        keep["raw_kps"] = keep["raw_kps"][:d.raw_kps_total]
        keep["geometry"] = g
        return kps[:n], desc[:n], keep

    def extract_many(self, imgs, nthreads=1, want_outputs=True):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        nf, h, w = imgs.shape
        cap = self._cap()
        kps = np.zeros((nf, cap), KP_DTYPE) if want_outputs else None
        desc = np.zeros((nf, cap, 32), np.uint8) if want_outputs else None
        counts = np.zeros(nf, np.int32)
        lib().orc_extract_many(self._h, _p(imgs), nf, w, h, nthreads, _p(kps), _p(desc), _p(counts), cap)
        return kps, desc, counts


def distribute_oct_tree(keys, min_x, max_x, min_y, max_y, n):
    """ORB-SLAM2's ORBextractor::DistributeOctTree (row f1); keys relative to (min_x, min_y)."""
    k = np.ascontiguousarray(keys, KP_DTYPE)
    out = np.zeros(len(k) + 1, KP_DTYPE)  # one keypoint per node, every node holds at least one
    m = lib().orc_distribute_oct_tree(_p(k), len(k), min_x, max_x, min_y, max_y, n, _p(out), len(out))
    assert m <= len(out)
    return out[:m]


# ---------------------------------------------------------------- matcher
def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().orc_descriptor_distance(_p(a), _p(b))


def match_best2(A, B, ratio=0.75, th_low=50, greedy=False):
    A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32)
    B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
    out = np.zeros(len(A), MATCH_DTYPE)
    fn = lib().orc_match_greedy if greedy else lib().orc_match_best2
    fn(_p(A), len(A), _p(B), len(B), ratio, th_low, _p(out))
    return out


def match_many(A, nA, B, nB, ratio=0.75, th_low=50, nthreads=1):
    A = np.ascontiguousarray(A, np.uint8)
    B = np.ascontiguousarray(B, np.uint8)
    npairs, sa, sb = A.shape[0], A.shape[1], B.shape[1]
    nA = np.ascontiguousarray(nA, np.int32)
    nB = np.ascontiguousarray(nB, np.int32)
    out = np.zeros((npairs, sa), MATCH_DTYPE)
    lib().orc_match_many(_p(A), _p(nA), _p(B), _p(nB), npairs, sa, sb, ratio, th_low, nthreads, _p(out))
    return out


def hamming_matrix(A, B):
    A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32)
    B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
    out = np.zeros((len(A), len(B)), np.uint16)
    lib().orc_hamming_matrix(_p(A), len(A), _p(B), len(B), _p(out))
    return out


def distinctive_many(desc, offsets, nthreads=1):
    """MapPoint::ComputeDistinctiveDescriptors for sets given as rows [offsets[s], offsets[s+1]) of desc."""
    desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    n = len(offsets) - 1
    best, med = np.zeros(n, np.int32), np.zeros(n, np.int32)
    lib().orc_distinctive_many(_p(desc), _p(offsets), n, nthreads, _p(best), _p(med))
    return best, med


def assign_grid(kps_un, min_x, min_y, inv_w, inv_h):
    """Frame::AssignFeaturesToGrid: returns (cell_start[64*48+1], indices[n_in_grid])."""
    k = np.ascontiguousarray(kps_un, KP_DTYPE)
    cs = np.zeros(64 * 48 + 1, np.int32)
    idx = np.zeros(max(len(k), 1), np.int32)
    lib().orc_assign_grid(_p(k), len(k), min_x, min_y, inv_w, inv_h, _p(cs), _p(idx))
    return cs, idx[:cs[-1]]


def stereo_from_rgbd(kps, kps_un, depth, mbf):
    k, ku = np.ascontiguousarray(kps, KP_DTYPE), np.ascontiguousarray(kps_un, KP_DTYPE)
    depth = np.ascontiguousarray(depth, np.float32)
    ur, z = np.zeros(len(k), np.float32), np.zeros(len(k), np.float32)
    lib().orc_stereo_from_rgbd(_p(k), _p(ku), len(k), _p(depth), depth.shape[1], mbf, _p(ur), _p(z))
    return ur, z


def undistort_keypoints(kps, K4, dist):
    """Frame::UndistortKeyPoints: K4 = (fx, fy, cx, cy), dist = (k1, k2, p1, p2[, k3]) as float32."""
    k = np.ascontiguousarray(kps, KP_DTYPE)
    K4 = np.ascontiguousarray(K4, np.float32)
    dist = np.ascontiguousarray(dist, np.float32)
    out = np.zeros(len(k), KP_DTYPE)
    lib().orc_undistort_keypoints(_p(k), len(k), _p(K4), _p(dist), len(dist), _p(out))
    return out


def image_bounds(cols, rows, K4, dist):
    K4 = np.ascontiguousarray(K4, np.float32)
    dist = np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    lib().orc_image_bounds(cols, rows, _p(K4), _p(dist), len(dist), _p(b))
    return b


def _grid(cell_start, indices, min_x, min_y, inv_w, inv_h):
    cs = np.ascontiguousarray(cell_start, np.int32)
    ix = np.ascontiguousarray(indices, np.int32)
    if len(ix) == 0:
        ix = np.zeros(1, np.int32)
    return FrameGrid(_p(cs), _p(ix), min_x, min_y, inv_w, inv_h), (cs, ix)


def features_in_area(kps_un, grid, x, y, r, min_level, max_level=-1):
    """Frame::GetFeaturesInArea; grid = (cell_start, indices, min_x, min_y, inv_w, inv_h)."""
    k = np.ascontiguousarray(kps_un, KP_DTYPE)
    g, keep = _grid(*grid)
    out = np.zeros(max(len(k), 1), np.int32)
    n = lib().orc_features_in_area(_p(k), C.byref(g), x, y, r, min_level, max_level, _p(out))
    return out[:n]


def three_maxima(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    a, b, c = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    lib().orc_three_maxima(_p(s), len(s), C.byref(a), C.byref(b), C.byref(c))
    return a.value, b.value, c.value


def search_for_initialization(kps1_un, desc1, kps2_un, desc2, grid2, prev_matched, window_size=100, nnratio=0.9,
                              check_orientation=True):
    """ORBmatcher::SearchForInitialization: returns (nmatches, matches12, prev_matched_updated)."""
    k1, k2 = np.ascontiguousarray(kps1_un, KP_DTYPE), np.ascontiguousarray(kps2_un, KP_DTYPE)
    d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    g, keep = _grid(*grid2)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = lib().orc_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), C.byref(g), _p(prev), window_size,
                                            nnratio, int(check_orientation), _p(m12))
    return n, m12[:len(k1)], prev


def search_by_projection(kps_last, kps_last_un, proj, flags_last, desc_mp, kps_cur_un, desc_cur, u_right_cur, occupied_cur,
                         grid_cur, scale_factors, bounds, th, mbf, mode, check_orientation=True, orb_dist=0):
    """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, bMono) from the projection on: (nmatches, assigned)."""
    kl, klu = np.ascontiguousarray(kps_last, KP_DTYPE), np.ascontiguousarray(kps_last_un, KP_DTYPE)
    kc = np.ascontiguousarray(kps_cur_un, KP_DTYPE)
    pr = np.ascontiguousarray(proj, np.float32).reshape(-1, 3)
    fl, oc = np.ascontiguousarray(flags_last, np.uint8), np.ascontiguousarray(occupied_cur, np.uint8)
    dm, dc = np.ascontiguousarray(desc_mp, np.uint8), np.ascontiguousarray(desc_cur, np.uint8)
    ur = np.ascontiguousarray(u_right_cur, np.float32)
    sf, bd = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(bounds, np.float32)
    g, keep = _grid(*grid_cur)
    asg = np.full(max(len(kc), 1), -1, np.int32)
    n = lib().orc_search_by_projection(_p(kl), _p(klu), _p(pr), _p(fl), _p(dm), len(kl), _p(kc), _p(dc), _p(ur), _p(oc), len(kc),
                                       C.byref(g), _p(sf), _p(bd), th, mbf, mode, int(check_orientation), _p(asg), int(orb_dist))
    return n, asg[:len(kc)]


def check_dist_epipolar_line(x1, y1, x2, y2, F12, sigma2):
    F = np.ascontiguousarray(F12, np.float64).reshape(9)
    return bool(lib().orc_check_dist_epipolar_line(x1, y1, x2, y2, _p(F), sigma2))


def search_for_triangulation(kps1_un, desc1, has_mp1, u_right1, kps2_un, desc2, has_mp2, u_right2, F12, ex, ey, scale_factors,
                             level_sigma2, check_orientation=True):
    """ORBmatcher::SearchForTriangulation from the epipole on: (nmatches, vMatches12)."""
    k1, k2 = np.ascontiguousarray(kps1_un, KP_DTYPE), np.ascontiguousarray(kps2_un, KP_DTYPE)
    d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    m1, m2 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(has_mp2, np.uint8)
    r1, r2 = np.ascontiguousarray(u_right1, np.float32), np.ascontiguousarray(u_right2, np.float32)
    F = np.ascontiguousarray(F12, np.float64).reshape(9)
    sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = lib().orc_search_for_triangulation(_p(k1), _p(d1), _p(m1), _p(r1), len(k1), _p(k2), _p(d2), _p(m2), _p(r2), len(k2), _p(F),
                                           ex, ey, _p(sf), _p(s2), int(check_orientation), _p(m12))
    return n, m12[:len(k1)]


def search_map_points(proj, view_cos, level, flags, desc_mp, kps_un, desc, u_right, occupied, grid, scale_factors, th, nnratio=0.8):
    """ORBmatcher::SearchByProjection(Frame, vpMapPoints, th): (nmatches, assigned)."""
    pr = np.ascontiguousarray(proj, np.float32).reshape(-1, 3)
    vc, lv = np.ascontiguousarray(view_cos, np.float32), np.ascontiguousarray(level, np.int32)
    fl, dm = np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc_mp, np.uint8)
    k, d = np.ascontiguousarray(kps_un, KP_DTYPE), np.ascontiguousarray(desc, np.uint8)
    ur, oc = np.ascontiguousarray(u_right, np.float32), np.ascontiguousarray(occupied, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    g, keep = _grid(*grid)
    asg = np.full(max(len(k), 1), -1, np.int32)
    n = lib().orc_search_map_points(_p(pr), _p(vc), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(ur), _p(oc), len(k), C.byref(g),
                                    _p(sf), th, nnratio, _p(asg))
    return n, asg[:len(k)]


def search_by_points(kps1_un, desc1, valid1, kps2_un, desc2, valid2, nnratio=0.75, check_orientation=True):
    """ORBmatcher::SearchByPoints: (nmatches, matches12)."""
    k1, k2 = np.ascontiguousarray(kps1_un, KP_DTYPE), np.ascontiguousarray(kps2_un, KP_DTYPE)
    d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
    v1, v2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = lib().orc_search_by_points(_p(k1), _p(d1), _p(v1), len(k1), _p(k2), _p(d2), _p(v2), len(k2), nnratio, int(check_orientation), _p(m12))
    return n, m12[:len(k1)]


def fuse_search(proj, level, flags, desc_mp, kps_un, desc, u_right, grid, scale_factors, inv_level_sigma2, th, check_reprojection=True,
                th_dist=50):
    """The keypoint search of ORBmatcher::Fuse (check_reprojection=False: its Sim3 overload; with th_dist=100 one direction of
    SearchBySim3): (best_idx, best_dist) per map point."""
    pr = np.ascontiguousarray(proj, np.float32).reshape(-1, 3)
    lv, fl, dm = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc_mp, np.uint8)
    k, d = np.ascontiguousarray(kps_un, KP_DTYPE), np.ascontiguousarray(desc, np.uint8)
    ur = np.ascontiguousarray(u_right, np.float32) if u_right is not None else None
    sf = np.ascontiguousarray(scale_factors, np.float32)
    is2 = np.ascontiguousarray(inv_level_sigma2, np.float32) if inv_level_sigma2 is not None else None
    assert not check_reprojection or (ur is not None and is2 is not None)
    g, keep = _grid(*grid)
    bi, bd = np.full(max(len(pr), 1), -1, np.int32), np.full(max(len(pr), 1), 256, np.int32)
    lib().orc_fuse_search(_p(pr), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(ur) if ur is not None else None, C.byref(g), _p(sf),
                          _p(is2) if is2 is not None else None, th, int(check_reprojection), int(th_dist), _p(bi), _p(bd))
    return bi[:len(pr)], bd[:len(pr)]


def search_by_sim3(side1, side2, scale_factors, th):
    """ORBmatcher::SearchBySim3 from the projections on.  side = (proj, level, flags, desc_mp, kps_un, desc, grid): the map points of
    that keyframe (entry i belongs to keypoint i) projected into the other one, and the keyframe's own keypoints / descriptors /
    grid.  Returns (nFound, matches12, vnMatch1, vnMatch2)."""
    def prep(side):
        pr, lv, fl, dm, k, d, grid = side
        g, keep = _grid(*grid)
        return [np.ascontiguousarray(pr, np.float32).reshape(-1, 3), np.ascontiguousarray(lv, np.int32), np.ascontiguousarray(fl, np.uint8),
                np.ascontiguousarray(dm, np.uint8), np.ascontiguousarray(k, KP_DTYPE), np.ascontiguousarray(d, np.uint8), g, keep]
    a, b = prep(side1), prep(side2)
    n1, n2 = len(a[4]), len(b[4])
    assert len(a[0]) == n1 and len(b[0]) == n2
    sf = np.ascontiguousarray(scale_factors, np.float32)
    m1, m2, m12 = np.full(max(n1, 1), -1, np.int32), np.full(max(n2, 1), -1, np.int32), np.full(max(n1, 1), -1, np.int32)
    n = lib().orc_search_by_sim3(_p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), n1, _p(b[0]), _p(b[1]), _p(b[2]), _p(b[3]), n2,
                                 _p(a[4]), _p(a[5]), C.byref(a[6]), _p(b[4]), _p(b[5]), C.byref(b[6]), _p(sf), _p(sf), th, _p(m1), _p(m2), _p(m12))
    return n, m12[:n1], m1[:n1], m2[:n2]


def search_by_projection_sim3(proj, level, flags, desc_mp, kps_un, desc, matched, grid, scale_factors, th):
    """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) from the projection on: (nmatches, assigned)."""
    pr = np.ascontiguousarray(proj, np.float32).reshape(-1, 3)
    lv, fl, dm = np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc_mp, np.uint8)
    k, d, mt = np.ascontiguousarray(kps_un, KP_DTYPE), np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(matched, np.uint8)
    sf = np.ascontiguousarray(scale_factors, np.float32)
    g, keep = _grid(*grid)
    asg = np.full(max(len(k), 1), -1, np.int32)
    n = lib().orc_search_by_projection_sim3(_p(pr), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(mt), len(k), C.byref(g), _p(sf), int(th), _p(asg))
    return n, asg[:len(k)]
