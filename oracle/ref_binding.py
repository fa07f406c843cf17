"""ctypes binding of oracle/_ref/libsdorb_ref.so: the reference's OWN sources (/root/reference/src/ORBextractor.cc, unmodified,
and ORBmatcher::DescriptorDistance) compiled against oracle/ref_compat by oracle/ref_build/Makefile.

TEST INFRASTRUCTURE ONLY: loaded by tests/ (to pin the oracle restatement and to generate tests/golden/), by
__graft_entry__.build() (building the checker is not using it) and by bench.py's CPU legs (cpu_baseline kind "reference").
/root/reference exists only in the build container: on the GPU box the prebuilt .so that travelled with the snapshot is used.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from oracle.binding import KP_DTYPE

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
SO = os.path.join(_DIR, "libsdorb_ref.so")
SO_NATIVE = os.path.join(_DIR, "libsdorb_ref_native.so")
REFERENCE = os.environ.get("SDORB_REFERENCE", "/root/reference")


def reference_present():
    return os.path.exists(os.path.join(REFERENCE, "src", "ORBextractor.cc"))


def build(force=False):
    """(Re)build oracle/_ref when the reference sources are present; otherwise the prebuilt library must already be there."""
    if reference_present():
        cmd = ["make", "-C", os.path.join(_HERE, "ref_build"), "-s", "REF=" + REFERENCE]
        if force:
            subprocess.check_call(cmd + ["clean"])
        subprocess.check_call(cmd)
    return SO if os.path.exists(SO) else None


def available():
    return os.path.exists(SO) or (reference_present() and build() is not None)


def _native_runs_here():
    """libsdorb_ref_native.so was compiled with -march=native on the build machine: only load it where every CPU flag of that
    machine is present."""
    f = os.path.join(_DIR, "native_cpu_flags.txt")
    if not (os.path.exists(SO_NATIVE) and os.path.exists(f)):
        return False
    try:
        built = set(open(f).read().split())
        with open("/proc/cpuinfo") as c:
            here = set(next(l for l in c if l.startswith("flags")).split(":", 1)[1].split())
    except Exception:
        return False
    return built <= here


_libs = {}


def lib(native=False):
    path = SO_NATIVE if native else SO
    if path not in _libs:
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        vp, i, f, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t
        L.ref_create.restype = vp
        L.ref_create.argtypes = [i, f, i, i]
        L.ref_destroy.argtypes = [vp]
        L.ref_get_tables.argtypes = [vp] * 10
        L.ref_extract.restype = i
        L.ref_extract.argtypes = [vp, vp, i, i, sz, vp, vp, i, vp, vp]
        L.ref_extract_many.restype = C.c_long
        L.ref_extract_many.argtypes = [vp, vp, i, i, i, i, vp, vp, vp, i]
        L.ref_descriptor_distance.restype = i
        L.ref_descriptor_distance.argtypes = [vp, vp]
        L.ref_hamming_matrix.argtypes = [vp, i, vp, i, vp]
        _libs[path] = L
    return _libs[path]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Extractor:
    """SD_SLAM::ORBextractor(nfeatures, scaleFactor, nlevels, thFAST) of the reference itself."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, th_fast=20, native=False):
        self._lib = lib(native)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self._h = self._lib.ref_create(nfeatures, scale_factor, nlevels, th_fast)

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.ref_destroy(self._h)
            self._h = None

    def tables(self):
        n = self.nlevels
        nl, sf = C.c_int(), C.c_float()
        a, b, c, d = (np.empty(n, np.float32) for _ in range(4))
        npl, umax, pat = np.empty(n, np.int32), np.empty(16, np.int32), np.empty(1024, np.int32)
        self._lib.ref_get_tables(self._h, C.addressof(nl), C.addressof(sf), _p(a), _p(b), _p(c), _p(d), _p(npl), _p(umax), _p(pat))
        return dict(nlevels=nl.value, scale_factor=sf.value, scale=a, inv_scale=b, sigma2=c, inv_sigma2=d, n_per_level=npl,
                    umax=umax, pattern=pat.reshape(512, 2))

    def level_sizes(self, w, h):
        inv = self.tables()["inv_scale"]
        # src/ORBextractor.cc:683: cvRound((float)cols * scale); np.rint is round-half-even like cvRound
        return [(int(np.rint(np.float32(w) * s)), int(np.rint(np.float32(h) * s))) for s in inv]

    def _cap(self):
        return max(int(self.tables()["n_per_level"].sum()), self.nfeatures, 1)

    def extract(self, img, pyramid=False):
        """(kps, desc) or (kps, desc, [level arrays], [padded level arrays]); raises where the reference throws cv::Exception."""
        img = np.asarray(img)
        assert img.dtype == np.uint8 and img.ndim == 2 and (img.size == 0 or img.strides[1] == 1)
        h, w = img.shape
        cap = self._cap()
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        tight = padded = None
        if pyramid:
            sizes = self.level_sizes(w, h)
            tight = np.zeros(sum(a * b for a, b in sizes), np.uint8)
            padded = np.zeros(sum((a + 38) * (b + 38) for a, b in sizes), np.uint8)
        n = self._lib.ref_extract(self._h, _p(img), w, h, img.strides[0], _p(kps), _p(desc), cap, _p(tight), _p(padded))
        if n == -4:
            raise RuntimeError("reference: cv::Exception (cell ROI outside the level image)")
        if n < 0:
            raise RuntimeError("reference shim error %d" % n)
        assert n <= cap
        if not pyramid:
            return kps[:n], desc[:n]
        lv, pd, o, q = [], [], 0, 0
        for a, b in sizes:
            lv.append(tight[o:o + a * b].reshape(b, a))
            pd.append(padded[q:q + (a + 38) * (b + 38)].reshape(b + 38, a + 38))
            o += a * b
            q += (a + 38) * (b + 38)
        return kps[:n], desc[:n], lv, pd

    def extract_many(self, imgs, nthreads=1, want_outputs=True):
        imgs = np.ascontiguousarray(imgs, np.uint8)
        nf, h, w = imgs.shape
        cap = self._cap()
        kps = np.zeros((nf, cap), KP_DTYPE) if want_outputs else None
        desc = np.zeros((nf, cap, 32), np.uint8) if want_outputs else None
        counts = np.zeros(nf, np.int32)
        self._lib.ref_extract_many(self._h, _p(imgs), nf, w, h, nthreads, _p(kps), _p(desc), _p(counts), cap)
        return kps, desc, counts


def descriptor_distance(a, b, native=False):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib(native).ref_descriptor_distance(_p(a), _p(b))


def hamming_matrix(A, B, native=False):
    A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32)
    B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
    out = np.zeros((len(A), len(B)), np.uint16)
    lib(native).ref_hamming_matrix(_p(A), len(A), _p(B), len(B), _p(out))
    return out


# ---------------------------------------------------------------- matcher / Frame object graph (oracle/ref_build/ref_matcher_shim.cc)
_MATCHER_SIGS_SET = set()


def _mlib(native=False):
    L = lib(native)
    if id(L) not in _MATCHER_SIGS_SET:
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_assign_grid.argtypes = [vp, i, f, f, f, f, vp, vp]
        L.ref_features_in_area.restype = i
        L.ref_features_in_area.argtypes = [vp, i, f, f, f, f, f, f, f, i, i, vp]
        L.ref_undistort_keypoints.argtypes = [vp, i, vp, vp, i, vp]
        L.ref_image_bounds.argtypes = [i, i, vp, vp, i, vp]
        L.ref_stereo_from_rgbd.argtypes = [vp, vp, i, vp, i, i, f, vp, vp]
        L.ref_three_maxima.argtypes = [vp, i, vp, vp, vp]
        L.ref_search_for_initialization.restype = i
        L.ref_search_for_initialization.argtypes = [vp, vp, i, vp, vp, i, f, f, f, f, vp, i, f, i, vp]
        L.ref_search_by_points.restype = i
        L.ref_search_by_points.argtypes = [vp, vp, vp, i, vp, vp, vp, i, f, i, vp]
        L.ref_distinctive.restype = i
        L.ref_distinctive.argtypes = [vp, i]
        L.ref_search_by_projection.restype = i
        L.ref_search_by_projection.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, i, vp, f, f, f, f, i, i, vp, i]
        L.ref_search_map_points.restype = i
        L.ref_search_map_points.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, i, vp, f, f, f, f, vp]
        _MATCHER_SIGS_SET.add(id(L))
    return L


def _k(a):
    return np.ascontiguousarray(a, KP_DTYPE)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def _f32(a):
    return np.ascontiguousarray(a, np.float32)


def assign_grid(kps_un, min_x, min_y, inv_w, inv_h):
    """Frame::AssignFeaturesToGrid of the reference: (cell_start[64*48+1], indices)."""
    k = _k(kps_un)
    cs = np.zeros(64 * 48 + 1, np.int32)
    idx = np.zeros(max(len(k), 1), np.int32)
    _mlib().ref_assign_grid(_p(k), len(k), min_x, min_y, inv_w, inv_h, _p(cs), _p(idx))
    return cs, idx[:cs[-1]]


def features_in_area(kps_un, gp, x, y, r, min_level, max_level=-1):
    """Frame::GetFeaturesInArea of the reference; gp = (min_x, min_y, inv_w, inv_h)."""
    k = _k(kps_un)
    out = np.zeros(max(len(k), 1), np.int32)
    n = _mlib().ref_features_in_area(_p(k), len(k), gp[0], gp[1], gp[2], gp[3], x, y, r, min_level, max_level, _p(out))
    return out[:n]


def undistort_keypoints(kps, K4, dist):
    k, K4, dist = _k(kps), _f32(K4), _f32(dist)
    out = np.zeros(len(k), KP_DTYPE)
    _mlib().ref_undistort_keypoints(_p(k), len(k), _p(K4), _p(dist), len(dist), _p(out))
    return out


def image_bounds(cols, rows, K4, dist):
    K4, dist = _f32(K4), _f32(dist)
    b = np.zeros(4, np.float32)
    _mlib().ref_image_bounds(cols, rows, _p(K4), _p(dist), len(dist), _p(b))
    return b


def stereo_from_rgbd(kps, kps_un, depth, mbf):
    k, ku, depth = _k(kps), _k(kps_un), _f32(depth)
    ur, z = np.zeros(len(k), np.float32), np.zeros(len(k), np.float32)
    _mlib().ref_stereo_from_rgbd(_p(k), _p(ku), len(k), _p(depth), depth.shape[1], depth.shape[0], mbf, _p(ur), _p(z))
    return ur, z


def three_maxima(sizes):
    s = np.ascontiguousarray(sizes, np.int32)
    a, b, c = C.c_int(-1), C.c_int(-1), C.c_int(-1)
    _mlib().ref_three_maxima(_p(s), len(s), C.addressof(a), C.addressof(b), C.addressof(c))
    return a.value, b.value, c.value


def search_for_initialization(kps1_un, desc1, kps2_un, desc2, gp, prev_matched, window_size=100, nnratio=0.9, check_orientation=True):
    k1, k2, d1, d2 = _k(kps1_un), _k(kps2_un), _u8(desc1), _u8(desc2)
    prev = np.array(prev_matched, np.float32).reshape(-1, 2).copy()
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = _mlib().ref_search_for_initialization(_p(k1), _p(d1), len(k1), _p(k2), _p(d2), len(k2), gp[0], gp[1], gp[2], gp[3], _p(prev),
                                              window_size, nnratio, int(check_orientation), _p(m12))
    return n, m12[:len(k1)], prev


def search_by_points(kps1_un, desc1, valid1, kps2_un, desc2, valid2, nnratio=0.75, check_orientation=True):
    k1, k2, d1, d2, v1, v2 = _k(kps1_un), _k(kps2_un), _u8(desc1), _u8(desc2), _u8(valid1), _u8(valid2)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = _mlib().ref_search_by_points(_p(k1), _p(d1), _p(v1), len(k1), _p(k2), _p(d2), _p(v2), len(k2), nnratio, int(check_orientation), _p(m12))
    return n, m12[:len(k1)]


def distinctive(desc):
    d = _u8(desc).reshape(-1, 32)
    return _mlib().ref_distinctive(_p(d), len(d))


def search_by_projection(kps_last, kps_last_un, proj, flags_last, desc_mp, kps_cur_un, desc_cur, u_right_cur, occupied_cur, gp,
                         scale_factors, bounds, th, mbf, mode, check_orientation=True, keyframe_overload=False):
    """SearchByProjection(CurrentFrame, LastFrame, th, bMono = false) -- or, keyframe_overload, its KeyFrame twin (:1077-1207);
    proj[:, 2] (invzc) must be 1 (see the shim)."""
    kl, klu, kc = _k(kps_last), _k(kps_last_un), _k(kps_cur_un)
    pr = _f32(proj).reshape(-1, 3)
    assert (pr[:, 2] == 1).all()
    fl, oc, dm, dc, ur = _u8(flags_last), _u8(occupied_cur), _u8(desc_mp), _u8(desc_cur), _f32(u_right_cur)
    sf, bd = _f32(scale_factors), _f32(bounds)
    asg = np.full(max(len(kc), 1), -1, np.int32)
    n = _mlib().ref_search_by_projection(_p(kl), _p(klu), _p(pr), _p(fl), _p(dm), len(kl), _p(kc), _p(dc), _p(ur), _p(oc), len(kc), _p(sf), len(sf),
                                         _p(bd), gp[2], gp[3], th, mbf, mode, int(check_orientation), _p(asg), int(keyframe_overload))
    return n, asg[:len(kc)]


def search_map_points(proj, view_cos, level, flags, desc_mp, kps_un, desc, u_right, occupied, gp, scale_factors, bounds, th, nnratio=0.8):
    pr = _f32(proj).reshape(-1, 3)
    vc, lv, fl, dm = _f32(view_cos), np.ascontiguousarray(level, np.int32), _u8(flags), _u8(desc_mp)
    k, d, ur, oc, sf, bd = _k(kps_un), _u8(desc), _f32(u_right), _u8(occupied), _f32(scale_factors), _f32(bounds)
    asg = np.full(max(len(k), 1), -1, np.int32)
    n = _mlib().ref_search_map_points(_p(pr), _p(vc), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(ur), _p(oc), len(k), _p(sf), len(sf),
                                      _p(bd), gp[2], gp[3], th, nnratio, _p(asg))
    return n, asg[:len(k)]


def _msigs2():
    L = _mlib()
    if not getattr(L, "_sdorb_sigs2", False):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_search_for_triangulation.restype = i
        L.ref_search_for_triangulation.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, f, f, vp, i, i, vp]
        L.ref_fuse_search.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, i, f, f, f, f, vp, i, f, f, vp]
        L.ref_search_by_sim3.restype = i
        L.ref_search_by_sim3.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, vp, i, vp, vp, vp, vp, f, f, f, f, vp, i, f, vp]
        L._sdorb_sigs2 = True
    return L


def search_for_triangulation(kps1_un, desc1, has_mp1, u_right1, kps2_un, desc2, has_mp2, u_right2, F12, ex, ey, scale_factors,
                             level_sigma2=None, check_orientation=True):
    """ORBmatcher::SearchForTriangulation of the reference: (nmatches, vMatches12).  level_sigma2 must be scale_factors ** 2 (the
    KeyFrame derives it)."""
    k1, k2, d1, d2 = _k(kps1_un), _k(kps2_un), _u8(desc1), _u8(desc2)
    m1, m2, r1, r2 = _u8(has_mp1), _u8(has_mp2), _f32(u_right1), _f32(u_right2)
    F = np.ascontiguousarray(F12, np.float64).reshape(9)
    sf = _f32(scale_factors)
    if level_sigma2 is not None:
        assert np.array_equal(_f32(level_sigma2), sf * sf)
    m12 = np.full(max(len(k1), 1), -1, np.int32)
    n = _msigs2().ref_search_for_triangulation(_p(k1), _p(d1), _p(m1), _p(r1), len(k1), _p(k2), _p(d2), _p(m2), _p(r2), len(k2), _p(F),
                                               float(ex), float(ey), _p(sf), len(sf), int(check_orientation), _p(m12))
    return n, m12[:len(k1)]


def fuse_search(proj, invz, level, flags, desc_mp, kps_un, desc, u_right, gp, scale_factors, th, bf):
    """The keypoint search of ORBmatcher::Fuse(pKF, vpMapPoints, th): best_idx per map point (-1: nothing fused).  proj[:, :2] =
    (u, v); the right coordinate the reference derives is u - bf * invz (invz: powers of two)."""
    pr, iz = _f32(proj).reshape(-1, 3), _f32(invz)
    lv, fl, dm = np.ascontiguousarray(level, np.int32), _u8(flags), _u8(desc_mp)
    k, d, ur, sf = _k(kps_un), _u8(desc), _f32(u_right), _f32(scale_factors)
    out = np.full(max(len(pr), 1), -1, np.int32)
    _msigs2().ref_fuse_search(_p(pr), _p(iz), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(ur), len(k), gp[0], gp[1], gp[2], gp[3], _p(sf),
                              len(sf), th, bf, _p(out))
    return out[:len(pr)]


def search_by_sim3(side1, side2, gp, scale_factors, th):
    """ORBmatcher::SearchBySim3 of the reference at the identity transform; side = (proj, level, flags, desc_mp, kps_un, desc):
    (nFound, matches12)."""
    p1, l1, f1, m1, k1, d1 = side1[:6]
    p2, l2, f2, m2, k2, d2 = side2[:6]
    p1, p2 = _f32(p1).reshape(-1, 3), _f32(p2).reshape(-1, 3)
    l1, l2 = np.ascontiguousarray(l1, np.int32), np.ascontiguousarray(l2, np.int32)
    f1, f2, m1, m2 = _u8(f1), _u8(f2), _u8(m1), _u8(m2)
    k1, k2, d1, d2, sf = _k(k1), _k(k2), _u8(d1), _u8(d2), _f32(scale_factors)
    out = np.full(max(len(k1), 1), -1, np.int32)
    n = _msigs2().ref_search_by_sim3(_p(p1), _p(l1), _p(f1), _p(m1), len(k1), _p(p2), _p(l2), _p(f2), _p(m2), len(k2), _p(k1), _p(d1), _p(k2), _p(d2),
                                     gp[0], gp[1], gp[2], gp[3], _p(sf), len(sf), th, _p(out))
    return n, out[:len(k1)]


def _msigs3():
    L = _msigs2()
    if not getattr(L, "_sdorb_sigs3", False):
        vp, i, f = C.c_void_p, C.c_int, C.c_float
        L.ref_fuse_sim3_search.argtypes = [vp, vp, vp, vp, i, vp, vp, i, f, f, f, f, vp, i, f, vp]
        L.ref_search_by_projection_sim3.restype = i
        L.ref_search_by_projection_sim3.argtypes = [vp, vp, vp, vp, i, vp, vp, vp, i, f, f, f, f, vp, i, i, vp]
        L.ref_search_by_projection_reloc.restype = i
        L.ref_search_by_projection_reloc.argtypes = [vp, vp, vp, vp, vp, i, vp, vp, vp, i, vp, i, vp, f, f, f, i, i, vp]
        L._sdorb_sigs3 = True
    return L


def fuse_sim3_search(proj, level, flags, desc_mp, kps_un, desc, gp, scale_factors, th):
    """The keypoint search of ORBmatcher::Fuse(pKF, Scw, vpPoints, th, vpReplacePoint) at Scw = identity: best_idx per point."""
    pr = _f32(proj).reshape(-1, 3)
    lv, fl, dm = np.ascontiguousarray(level, np.int32), _u8(flags), _u8(desc_mp)
    k, d, sf = _k(kps_un), _u8(desc), _f32(scale_factors)
    out = np.full(max(len(pr), 1), -1, np.int32)
    _msigs3().ref_fuse_sim3_search(_p(pr), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), len(k), gp[0], gp[1], gp[2], gp[3], _p(sf), len(sf), th,
                                   _p(out))
    return out[:len(pr)]


def search_by_projection_sim3(proj, level, flags, desc_mp, kps_un, desc, matched, gp, scale_factors, th):
    """ORBmatcher::SearchByProjection(pKF, Scw, vpPoints, vpMatched, th) at Scw = identity: (nmatches, assigned)."""
    pr = _f32(proj).reshape(-1, 3)
    lv, fl, dm = np.ascontiguousarray(level, np.int32), _u8(flags), _u8(desc_mp)
    k, d, mt, sf = _k(kps_un), _u8(desc), _u8(matched), _f32(scale_factors)
    out = np.full(max(len(k), 1), -1, np.int32)
    n = _msigs3().ref_search_by_projection_sim3(_p(pr), _p(lv), _p(fl), _p(dm), len(pr), _p(k), _p(d), _p(mt), len(k), gp[0], gp[1], gp[2], gp[3],
                                                _p(sf), len(sf), int(th), _p(out))
    return n, out[:len(k)]


def search_by_projection_reloc(kps_kf_un, proj, valid, pred_level, desc_mp, kps_cur_un, desc_cur, has_mp_cur, gp, scale_factors, bounds, th,
                               orb_dist, check_orientation=True):
    """ORBmatcher::SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist): (nmatches, assigned)."""
    kk, kc = _k(kps_kf_un), _k(kps_cur_un)
    pr = _f32(proj).reshape(-1, 3)
    assert (pr[:, 2] == 1).all()
    vl, pl, dm = _u8(valid), np.ascontiguousarray(pred_level, np.int32), _u8(desc_mp)
    dc, hm, sf, bd = _u8(desc_cur), _u8(has_mp_cur), _f32(scale_factors), _f32(bounds)
    out = np.full(max(len(kc), 1), -1, np.int32)
    n = _msigs3().ref_search_by_projection_reloc(_p(kk), _p(pr), _p(vl), _p(pl), _p(dm), len(kk), _p(kc), _p(dc), _p(hm), len(kc), _p(sf), len(sf),
                                                 _p(bd), gp[2], gp[3], th, int(orb_dist), int(check_orientation), _p(out))
    return n, out[:len(kc)]
