"""Host-side mirror of the reference interface over the C ABI (include/sdorb.h) via ctypes.

`ORBextractor` keeps the reference's constructor / call surface
(/root/reference/src/ORBextractor.h:38-70: ORBextractor(nfeatures, scaleFactor, nlevels, thFAST),
operator()(image, mask, keypoints, descriptors, imagePyramid), the six getters) and adds the batched entry
points that make the GPU worthwhile.  `ORBmatcher.DescriptorDistance` mirrors src/ORBmatcher.h:43.

Everything here goes through libsdorb.so; there is no Python or CPU implementation of the path.  If the
library is missing it is built with nvcc, and if that fails the import raises.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                     ("octave", "<i4"), ("class_id", "<i4")])
MATCH_DTYPE = np.dtype([("best_idx", "<i4"), ("best_dist", "<i4"), ("second_dist", "<i4"), ("accepted", "<i4")])
MEM_HOST, MEM_DEVICE = 0, 1
STAGES = ("pyramid", "fast", "select", "blur", "describe", "match")
DBG_PYRAMID_LEVEL, DBG_BLURRED_LEVEL, DBG_CELL_COUNTS, DBG_LEVEL_SELECTED = 0, 1, 2, 3
EDGE_THRESHOLD = 19


class SdorbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("sdorb error %d: %s" % (code, msg))
        self.code = code


class _Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int), ("scale_factor", C.c_float), ("nlevels", C.c_int), ("th_fast", C.c_int),
                ("min_th_fast", C.c_int), ("device", C.c_int), ("max_width", C.c_int), ("max_height", C.c_int),
                ("max_batch", C.c_int)]


class _PyrView(C.Structure):
    _fields_ = [("data", C.c_void_p), ("width", C.c_int), ("height", C.c_int), ("stride", C.c_size_t),
                ("border", C.c_int)]


class _FrameGrid(C.Structure):
    _fields_ = [("cell_start", C.c_void_p), ("indices", C.c_void_p), ("min_x", C.c_float), ("min_y", C.c_float),
                ("inv_w", C.c_float), ("inv_h", C.c_float)]


class _ProjectionSearch(C.Structure):
    _fields_ = [("kps_last", C.c_void_p), ("kps_last_un", C.c_void_p), ("proj", C.c_void_p), ("flags_last", C.c_void_p),
                ("desc_mp", C.c_void_p), ("n_last", C.c_void_p), ("kps_cur_un", C.c_void_p), ("desc_cur", C.c_void_p),
                ("u_right_cur", C.c_void_p), ("occupied_cur", C.c_void_p), ("n_cur", C.c_void_p), ("grid_cur", _FrameGrid),
                ("scale_factors", C.c_void_p), ("nlevels", C.c_int), ("bounds", C.c_float * 4), ("th", C.c_float),
                ("mbf", C.c_float), ("mode", C.c_int), ("check_orientation", C.c_int), ("orb_dist", C.c_int)]


class _MapPointSearch(C.Structure):
    _fields_ = [("proj", C.c_void_p), ("view_cos", C.c_void_p), ("level", C.c_void_p), ("flags", C.c_void_p), ("desc_mp", C.c_void_p),
                ("n_mp", C.c_void_p), ("kps_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("occupied", C.c_void_p),
                ("n_frame", C.c_void_p), ("grid", _FrameGrid), ("scale_factors", C.c_void_p), ("nlevels", C.c_int), ("th", C.c_float),
                ("nnratio", C.c_float), ("sim3_form", C.c_int)]


class _FuseSearch(C.Structure):
    _fields_ = [("proj", C.c_void_p), ("level", C.c_void_p), ("flags", C.c_void_p), ("desc_mp", C.c_void_p), ("n_mp", C.c_void_p),
                ("kps_un", C.c_void_p), ("desc", C.c_void_p), ("u_right", C.c_void_p), ("grid", _FrameGrid),
                ("scale_factors", C.c_void_p), ("inv_level_sigma2", C.c_void_p), ("nlevels", C.c_int), ("th", C.c_float),
                ("check_reprojection", C.c_int), ("th_dist", C.c_int)]


class _TriangulationSearch(C.Structure):
    _fields_ = [("kps1_un", C.c_void_p), ("desc1", C.c_void_p), ("has_mp1", C.c_void_p), ("u_right1", C.c_void_p), ("n1", C.c_void_p),
                ("kps2_un", C.c_void_p), ("desc2", C.c_void_p), ("has_mp2", C.c_void_p), ("u_right2", C.c_void_p), ("n2", C.c_void_p),
                ("F12", C.c_void_p), ("epipole", C.c_void_p), ("scale_factors", C.c_void_p), ("level_sigma2", C.c_void_p),
                ("nlevels", C.c_int), ("check_orientation", C.c_int)]


_lib = None


def lib():
    """Loads (building if necessary) libsdorb.so.  Raises if the CUDA library cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    so = _build.build()
    L = C.CDLL(so)
    vp, i, sz, f, i64 = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_int64
    L.sdorb_create.argtypes = [C.POINTER(_Params), C.POINTER(vp)]
    L.sdorb_destroy.argtypes = [vp]
    L.sdorb_destroy.restype = None
    L.sdorb_strerror.argtypes = [i]
    L.sdorb_strerror.restype = C.c_char_p
    L.sdorb_last_cuda_error.argtypes = [vp]
    L.sdorb_last_cuda_error.restype = C.c_char_p
    L.sdorb_get_tables.argtypes = [vp] * 6
    L.sdorb_max_keypoints.argtypes = [vp]
    L.sdorb_level_size.argtypes = [vp, i, i, i, C.POINTER(i), C.POINTER(i)]
    L.sdorb_extract.argtypes = [vp, vp, i, i, sz, vp, vp, i, C.POINTER(i), vp]
    L.sdorb_extract_batch.argtypes = [vp, vp, i, i, i, sz, sz, vp, vp, vp, i, i, vp]
    L.sdorb_extract_batch_pyr.argtypes = [vp, vp, i, i, i, sz, sz, vp, vp, vp, i, vp, i, i, vp]
    L.sdorb_pyramid_layout.argtypes = [vp, i, i, vp, vp]
    L.sdorb_shard_range.argtypes = [i, i, i, C.POINTER(i), C.POINTER(i)]
    L.sdorb_extract_batch_multi.argtypes = [vp, i, vp, i, i, i, sz, sz, vp, vp, vp, i, vp, i]
    L.sdorb_batch_status.argtypes = [vp]
    L.sdorb_match_batch.argtypes = [vp, vp, vp, i, vp, vp, i, i, f, i, vp, i, vp]
    L.sdorb_match_greedy_batch.argtypes = [vp, vp, vp, i, vp, vp, i, i, f, i, vp, i, vp]
    L.sdorb_hamming_matrix.argtypes = [vp, vp, i, vp, i, vp, i, vp]
    L.sdorb_distinctive_batch.argtypes = [vp, vp, vp, i, vp, vp, i, vp]
    L.sdorb_assign_grid_batch.argtypes = [vp, vp, vp, i, i, f, f, f, f, vp, vp, i, vp]
    L.sdorb_undistort_keypoints_batch.argtypes = [vp, vp, vp, i, i, vp, vp, i, vp, i, vp]
    L.sdorb_host_image_bounds.argtypes = [i, i, vp, vp, i, vp]
    L.sdorb_search_for_initialization_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(_FrameGrid), i, i, vp, i, f, i, vp, vp,
                                                        i, vp]
    L.sdorb_search_by_projection_batch.argtypes = [vp, C.POINTER(_ProjectionSearch), i, i, vp, vp, i, vp]
    L.sdorb_search_map_points_batch.argtypes = [vp, C.POINTER(_MapPointSearch), i, i, i, vp, vp, i, vp]
    L.sdorb_search_by_points_batch.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, vp, i, i, f, i, vp, vp, i, vp]
    L.sdorb_fuse_search_batch.argtypes = [vp, C.POINTER(_FuseSearch), i, i, i, vp, vp, i, vp]
    L.sdorb_search_by_sim3_batch.argtypes = [vp, C.POINTER(_FuseSearch), C.POINTER(_FuseSearch), i, i, vp, vp, vp, vp, i, vp]
    L.sdorb_search_for_triangulation_batch.argtypes = [vp, C.POINTER(_TriangulationSearch), i, i, vp, vp, i, vp]
    L.sdorb_stereo_from_rgbd_batch.argtypes = [vp, vp, vp, vp, i, i, vp, i, i, sz, sz, f, vp, vp, i, vp]
    L.sdorb_fill_border_reflect101.argtypes = [vp, i, i, sz, i]
    L.sdorb_fill_border_reflect101.restype = None
    L.sdorb_host_tables.argtypes = [i, f, i] + [vp] * 6
    L.sdorb_host_level_geometry.argtypes = [i, f, i, i, i, i, vp]
    L.sdorb_set_profiling.argtypes = [vp, i]
    L.sdorb_get_stage_times.argtypes = [vp, vp, vp, i]
    L.sdorb_kernel_launches.argtypes = [vp]
    L.sdorb_kernel_launches.restype = i64
    L.sdorb_debug_read.argtypes = [vp, i, i, i, vp, sz]
    L.sdorb_debug_read.restype = i64
    L.sdorb_debug_nth_element.argtypes = [vp, vp, i, i]
    L.sdorb_debug_pipe_probe.argtypes = [vp, i, C.POINTER(C.c_double), C.POINTER(C.c_double)]
    L.sdorb_debug_guard_check.argtypes = [vp]
    L.sdorb_debug_guard_check.restype = i64
    _lib = L
    return L


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):  # torch tensor
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(int(a))


GEOM_DTYPE = np.dtype([(n, "<i4") for n in ("width", "height", "n_desired", "level_cols", "level_rows", "cell_w",
                                             "cell_h", "n_features_cell", "scaled_patch_size")])


def host_tables(nfeatures, scaleFactor, nlevels):
    """The constructor tables (src/ORBextractor.cc:406-457) computed on the host; needs no GPU."""
    sf, isf, s2, is2 = (np.empty(nlevels, np.float32) for _ in range(4))
    npl, umax = np.empty(nlevels, np.int32), np.empty(16, np.int32)
    rc = lib().sdorb_host_tables(nfeatures, scaleFactor, nlevels, _ptr(sf), _ptr(isf), _ptr(s2), _ptr(is2), _ptr(npl),
                                 _ptr(umax))
    if rc:
        raise SdorbError(rc, lib().sdorb_strerror(rc).decode())
    return dict(scale=sf, inv_scale=isf, sigma2=s2, inv_sigma2=is2, n_per_level=npl, umax=umax)


def host_level_geometry(nfeatures, scaleFactor, nlevels, thFAST, width, height):
    """Level sizes and cell grids (src/ORBextractor.cc:469-488, 683) for one input size; needs no GPU."""
    g = np.zeros(nlevels, GEOM_DTYPE)
    rc = lib().sdorb_host_level_geometry(nfeatures, scaleFactor, nlevels, thFAST, width, height, _ptr(g))
    if rc:
        raise SdorbError(rc, lib().sdorb_strerror(rc).decode())
    return g


def host_image_bounds(cols, rows, K4, dist):
    """Frame::ComputeImageBounds (src/Frame.cc:368-397): (mnMinX, mnMaxX, mnMinY, mnMaxY); needs no GPU."""
    K4, dist = np.ascontiguousarray(K4, np.float32), np.ascontiguousarray(dist, np.float32)
    b = np.zeros(4, np.float32)
    rc = lib().sdorb_host_image_bounds(cols, rows, _ptr(K4), _ptr(dist), len(dist), _ptr(b))
    if rc:
        raise SdorbError(rc, lib().sdorb_strerror(rc).decode())
    return b


def shard_range(shard, nshards, nframes):
    a, b = C.c_int(), C.c_int()
    rc = lib().sdorb_shard_range(shard, nshards, nframes, C.byref(a), C.byref(b))
    if rc:
        raise SdorbError(rc, lib().sdorb_strerror(rc).decode())
    return a.value, b.value


def extract_batch_multi(extractors, images, keypoints=None, descriptors=None, counts=None, pyramid=None, first_level=1):
    """sdorb_extract_batch_multi: one host batch over several handles (one per GPU), each on its own host thread inside the
    library.  images: uint8 [F,H,W] host array / pinned tensor.  Returns (kps[F,cap], desc[F,cap,32], counts[F])."""
    nf, h, w = images.shape
    cap = max(extractors[0].max_keypoints, 1)
    is_t = hasattr(images, "data_ptr")
    if keypoints is None:
        keypoints = np.zeros((nf, cap), KP_DTYPE)
        descriptors = np.zeros((nf, cap, 32), np.uint8)
        counts = np.zeros(nf, np.int32)
    row = images.stride(1) if is_t else images.strides[1]
    frame = images.stride(0) if is_t else images.strides[0]
    hs = (C.c_void_p * len(extractors))(*[e._h for e in extractors])
    rc = lib().sdorb_extract_batch_multi(C.cast(hs, C.c_void_p), len(extractors), _ptr(images), nf, w, h, row, frame, _ptr(keypoints),
                                         _ptr(descriptors), _ptr(counts), cap, _ptr(pyramid), first_level)
    if rc:
        extractors[0]._check(rc)
    return keypoints, descriptors, counts


class ORBextractor:
    """SD_SLAM::ORBextractor on a B200.  One instance = one GPU handle (not thread-safe, like one CUDA stream)."""

    HARRIS_SCORE, FAST_SCORE = 0, 1  # src/ORBextractor.h:36

    def __init__(self, nfeatures, scaleFactor, nlevels, thFAST, minThFAST=None, device=-1, max_width=1920,
                 max_height=1088, max_batch=64):
        # minThFAST given (the north-star 5-argument form, thFAST = iniThFAST): the ORB-SLAM2-style mode -- 30-pixel cells
        # with the ini / min threshold fallback and DistributeOctTree (not in SD-SLAM itself,
        # SURVEY.md section 0).  None: the reference's ComputeKeyPoints.
        self._h = C.c_void_p()
        self.nfeatures, self.scaleFactor, self.nlevels, self.thFAST = nfeatures, scaleFactor, nlevels, thFAST
        self.max_batch = max_batch
        self.minThFAST = minThFAST
        p = _Params(nfeatures, scaleFactor, nlevels, thFAST, -1 if minThFAST is None else int(minThFAST), device, max_width,
                    max_height, max_batch)
        self._check(lib().sdorb_create(C.byref(p), C.byref(self._h)))
        self.max_keypoints = lib().sdorb_max_keypoints(self._h)
        n = nlevels
        self._sf, self._isf, self._s2, self._is2 = (np.empty(n, np.float32) for _ in range(4))
        self._npl = np.empty(n, np.int32)
        self._check(lib().sdorb_get_tables(self._h, _ptr(self._sf), _ptr(self._isf), _ptr(self._s2), _ptr(self._is2),
                                            _ptr(self._npl)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            lib().sdorb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except TypeError:  # interpreter shutdown: the module globals are gone, and so is the process
            pass

    def _check(self, rc):
        if rc != 0:
            msg = lib().sdorb_strerror(rc).decode()
            if rc == -3 and self._h.value:
                msg += " (" + lib().sdorb_last_cuda_error(self._h).decode() + ")"
            raise SdorbError(rc, msg)

    # ---- getters, src/ORBextractor.h:48-70
    def GetLevels(self):
        return self.nlevels

    def GetScaleFactor(self):
        return float(np.float32(self.scaleFactor))

    def GetScaleFactors(self):
        return self._sf.copy()

    def GetInverseScaleFactors(self):
        return self._isf.copy()

    def GetScaleSigmaSquares(self):
        return self._s2.copy()

    def GetInverseScaleSigmaSquares(self):
        return self._is2.copy()

    def features_per_level(self):
        return self._npl.copy()

    def level_size(self, width, height, level):
        w, h = C.c_int(), C.c_int()
        self._check(lib().sdorb_level_size(self._h, width, height, level, C.byref(w), C.byref(h)))
        return w.value, h.value

    # ---- operator()(image, mask, keypoints, descriptors, imagePyramid), src/ORBextractor.cc:620-678
    def __call__(self, image, mask=None, want_pyramid=True):
        """Returns (keypoints[KP_DTYPE], descriptors[N,32] uint8, pyramid list).  `mask` is ignored, as in the
        reference.  An empty image returns (None, None, None): the reference leaves its outputs untouched."""
        image = np.asarray(image)
        if image.size == 0:
            return None, None, None
        assert image.dtype == np.uint8 and image.ndim == 2 and image.strides[1] == 1, "CV_8UC1 image expected"
        h, w = image.shape
        cap = max(self.max_keypoints, 1)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        views, pyr, padded = None, None, []
        if want_pyramid:
            views = (_PyrView * self.nlevels)()
            pyr = []
            for l in range(self.nlevels):
                lw, lh = self.level_size(w, h, l)
                buf = np.zeros((max(lh, 0) + 2 * EDGE_THRESHOLD, max(lw, 0) + 2 * EDGE_THRESHOLD), np.uint8)
                inner = buf[EDGE_THRESHOLD:EDGE_THRESHOLD + lh, EDGE_THRESHOLD:EDGE_THRESHOLD + lw]
                padded.append(buf)
                pyr.append(inner)  # a view into the padded buffer, like imagePyramid[l] in the reference
                views[l] = _PyrView(inner.ctypes.data if inner.size else None, lw, lh, buf.strides[0], EDGE_THRESHOLD)
        self._check(lib().sdorb_extract(self._h, _ptr(image), w, h, image.strides[0], _ptr(kps), _ptr(desc), cap,
                                         C.byref(n), C.cast(views, C.c_void_p) if views is not None else None))
        return kps[:n.value], desc[:n.value], pyr

    def single_frame_call(self, width, height, want_pyramid=True):
        """A reusable single-frame call with every output buffer allocated once (what a C++ caller that keeps its cv::Mat
        buffers does): returns f(image) -> (count, kps, desc, pyramid views) with the arrays overwritten by every call."""
        cap = max(self.max_keypoints, 1)
        kps = np.zeros(cap, KP_DTYPE)
        desc = np.zeros((cap, 32), np.uint8)
        n = C.c_int(0)
        views, pyr, keep = None, None, []
        if want_pyramid:
            views = (_PyrView * self.nlevels)()
            pyr = []
            for l in range(self.nlevels):
                lw, lh = self.level_size(width, height, l)
                buf = np.zeros((max(lh, 0) + 2 * EDGE_THRESHOLD, max(lw, 0) + 2 * EDGE_THRESHOLD), np.uint8)
                inner = buf[EDGE_THRESHOLD:EDGE_THRESHOLD + lh, EDGE_THRESHOLD:EDGE_THRESHOLD + lw]
                keep.append(buf)
                pyr.append(inner)
                views[l] = _PyrView(inner.ctypes.data if inner.size else None, lw, lh, buf.strides[0], EDGE_THRESHOLD)
        vp = C.cast(views, C.c_void_p) if views is not None else None
        fn, hnd, pk, pd, pn = lib().sdorb_extract, self._h, _ptr(kps), _ptr(desc), C.byref(n)

        def call(image):
            rc = fn(hnd, image.ctypes.data_as(C.c_void_p), width, height, image.strides[0], pk, pd, cap, pn, vp)
            if rc:
                self._check(rc)
            return n.value, kps, desc, pyr

        call._keep = keep
        return call

    def pyramid_layout(self, width, height):
        """(level_offset[nlevels], frame_bytes) of the imagePyramid slab of the batch calls (sdorb_pyramid_layout)."""
        off = (C.c_size_t * self.nlevels)()
        fb = C.c_size_t(0)
        self._check(lib().sdorb_pyramid_layout(self._h, width, height, C.cast(off, C.c_void_p), C.cast(C.byref(fb), C.c_void_p)))
        return [int(o) for o in off], int(fb.value)

    def pyramid_levels(self, slab, frame, width, height):
        """Views of the levels of `frame` in a host slab filled by extract_batch_host(..., pyramid=slab)."""
        off, fb = self.pyramid_layout(width, height)
        out = []
        for l in range(self.nlevels):
            lw, lh = self.level_size(width, height, l)
            a = frame * fb + off[l]
            out.append(slab[a:a + lw * lh].reshape(lh, lw))
        return out

    def extract_batch_host(self, images, keypoints=None, descriptors=None, counts=None, pyramid=None, first_level=1):
        """images: uint8 [F,H,W] host array (or pinned torch CPU tensor).  Returns (kps[F,cap], desc[F,cap,32], counts[F]).
        pyramid: None, or a flat uint8 host buffer of F * frame_bytes (pyramid_layout) that receives imagePyramid of every frame
        (levels first_level ..; first_level = 0 includes the copy of the input the reference returns as level 0)."""
        nf, h, w = images.shape
        cap = max(self.max_keypoints, 1)
        is_t = hasattr(images, "data_ptr")
        if keypoints is None:
            keypoints = np.zeros((nf, cap), KP_DTYPE)
            descriptors = np.zeros((nf, cap, 32), np.uint8)
            counts = np.zeros(nf, np.int32)
        row = images.stride(1) if is_t else images.strides[1]
        frame = images.stride(0) if is_t else images.strides[0]
        self._check(lib().sdorb_extract_batch_pyr(self._h, _ptr(images), nf, w, h, row, frame, _ptr(keypoints),
                                                   _ptr(descriptors), _ptr(counts), cap, _ptr(pyramid), first_level, MEM_HOST, None))
        return keypoints, descriptors, counts

    def extract_batch_device(self, images, keypoints, descriptors, counts, stream=None, pyramid=None, first_level=1):
        """All arguments are torch CUDA tensors on this handle's device: images uint8 [F,H,W]; keypoints
        [F,cap,7] float32 (cv::KeyPoint layout, octave / class_id as int bits); descriptors uint8 [F,cap,32];
        counts int32 [F].  Work is enqueued on `stream` (a raw cudaStream_t int; None = torch's current)."""
        import torch
        nf, h, w = images.shape
        cap = keypoints.shape[1]
        if stream is None:
            stream = torch.cuda.current_stream(images.device).cuda_stream
        self._check(lib().sdorb_extract_batch_pyr(self._h, _ptr(images), nf, w, h, images.stride(1), images.stride(0),
                                                   _ptr(keypoints), _ptr(descriptors), _ptr(counts), cap, _ptr(pyramid), first_level,
                                                   MEM_DEVICE, C.c_void_p(stream)))

    def batch_status(self):
        self._check(lib().sdorb_batch_status(self._h))

    # ---- matcher entry points live on the handle (they share its stream and scratch)
    def match_batch(self, descA, nA, descB, nB, ratio=0.75, th_low=50, out=None, device=False, stream=None, greedy=False):
        npairs, sa = descA.shape[0], descA.shape[1]
        sb = descB.shape[1]
        if out is None:
            assert not device
            out = np.zeros((npairs, sa), MATCH_DTYPE)
        fn = lib().sdorb_match_greedy_batch if greedy else lib().sdorb_match_batch
        if device and stream is None:
            import torch
            stream = torch.cuda.current_stream(descA.device).cuda_stream
        self._check(fn(self._h, _ptr(descA), _ptr(nA), sa, _ptr(descB), _ptr(nB), sb, npairs, ratio, th_low, _ptr(out),
                       MEM_DEVICE if device else MEM_HOST, C.c_void_p(stream) if stream else None))
        return out

    def hamming_matrix(self, A, B):
        A = np.ascontiguousarray(A, np.uint8).reshape(-1, 32)
        B = np.ascontiguousarray(B, np.uint8).reshape(-1, 32)
        out = np.zeros((len(A), len(B)), np.uint16)
        self._check(lib().sdorb_hamming_matrix(self._h, _ptr(A), len(A), _ptr(B), len(B), _ptr(out), MEM_HOST, None))
        return out

    def distinctive_batch(self, desc, offsets, best_idx=None, best_median=None, device=False, stream=None):
        """MapPoint::ComputeDistinctiveDescriptors (src/MapPoint.cc:225-284) for many map points at once: set s = rows
        [offsets[s], offsets[s+1]) of desc.  Returns (best_idx, best_median) per set."""
        nsets = len(offsets) - 1
        if not device:
            desc = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
            offsets = np.ascontiguousarray(offsets, np.int32)
            best_idx, best_median = np.zeros(nsets, np.int32), np.zeros(nsets, np.int32)
        elif stream is None:
            import torch
            stream = torch.cuda.current_stream(desc.device).cuda_stream
        self._check(lib().sdorb_distinctive_batch(self._h, _ptr(desc), _ptr(offsets), nsets, _ptr(best_idx), _ptr(best_median),
                                                   MEM_DEVICE if device else MEM_HOST, C.c_void_p(stream) if stream else None))
        return best_idx, best_median

    # ---- Frame post-processing (src/Frame.cc:179-192, 323-332, 399-417)
    GRID_COLS, GRID_ROWS = 64, 48

    def assign_grid_batch(self, keypoints_un, counts, min_x, min_y, inv_w, inv_h, cell_start=None, indices=None, device=False,
                          stream=None):
        """Frame::AssignFeaturesToGrid for a batch: returns (cell_start[F, 64*48+1], indices[F, cap]).  Host arrays, or with
        device=True torch tensors on the handle's GPU (keypoints [F, cap, 7] float32, outputs int32, caller-allocated)."""
        if not device:
            keypoints_un = np.ascontiguousarray(keypoints_un)
            counts = np.ascontiguousarray(counts, np.int32)
            cell_start = np.zeros((keypoints_un.shape[0], self.GRID_COLS * self.GRID_ROWS + 1), np.int32)
            indices = np.zeros(keypoints_un.shape[:2], np.int32)
        nf, cap = keypoints_un.shape[0], keypoints_un.shape[1]
        self._check(lib().sdorb_assign_grid_batch(self._h, _ptr(keypoints_un), _ptr(counts), nf, cap, min_x, min_y, inv_w, inv_h,
                                                   _ptr(cell_start), _ptr(indices), MEM_DEVICE if device else MEM_HOST,
                                                   C.c_void_p(stream) if stream else None))
        return cell_start, indices

    def undistort_keypoints_batch(self, keypoints, counts, K4, dist):
        """Frame::UndistortKeyPoints for a batch (host arrays); K4 = (fx, fy, cx, cy), dist = (k1, k2, p1, p2[, k3])."""
        kps = np.ascontiguousarray(keypoints)
        counts = np.ascontiguousarray(counts, np.int32)
        K4, dist = np.ascontiguousarray(K4, np.float32), np.ascontiguousarray(dist, np.float32)
        out = np.zeros_like(kps)
        self._check(lib().sdorb_undistort_keypoints_batch(self._h, _ptr(kps), _ptr(counts), kps.shape[0], kps.shape[1], _ptr(K4),
                                                           _ptr(dist), len(dist), _ptr(out), MEM_HOST, None))
        return out

    def stereo_from_rgbd_batch(self, keypoints, keypoints_un, counts, depth, mbf):
        """Frame::ComputeStereoFromRGBD for a batch (host arrays, depth float32 [F,H,W]): returns (u_right, z) [F, cap]."""
        kps, kun = np.ascontiguousarray(keypoints), np.ascontiguousarray(keypoints_un)
        depth = np.ascontiguousarray(depth, np.float32)
        nf, cap = kps.shape[0], kps.shape[1]
        counts = np.ascontiguousarray(counts, np.int32)
        ur, z = np.zeros((nf, cap), np.float32), np.zeros((nf, cap), np.float32)
        self._check(lib().sdorb_stereo_from_rgbd_batch(self._h, _ptr(kps), _ptr(kun), _ptr(counts), nf, cap, _ptr(depth), depth.shape[2],
                                                        depth.shape[1], depth.shape[2], depth.shape[1] * depth.shape[2], mbf,
                                                        _ptr(ur), _ptr(z), MEM_HOST, None))
        return ur, z

    # ---- guided matchers (ORBmatcher::SearchForInitialization / SearchByProjection(Frame, Frame)), host arrays
    def search_for_initialization_batch(self, kps1_un, desc1, n1, kps2_un, desc2, n2, grid2, prev_matched, window_size=100,
                                        nnratio=0.9, check_orientation=True, matches12=None, nmatches=None, device=False,
                                        stream=None):
        """grid2 = (cell_start[P, 3073], indices[P, cap], min_x, min_y, inv_w, inv_h) of the second frames (assign_grid_batch).
        Returns (nmatches[P], matches12[P, cap], prev_matched[P, cap, 2] updated).  Host arrays (prev_matched is copied), or
        with device=True torch tensors on the handle's GPU: prev_matched is updated in place, matches12 / nmatches are the
        caller's int32 tensors."""
        if not device:
            kps1_un, kps2_un = np.ascontiguousarray(kps1_un), np.ascontiguousarray(kps2_un)
            desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
            n1, n2 = np.ascontiguousarray(n1, np.int32), np.ascontiguousarray(n2, np.int32)
            cs, ix = np.ascontiguousarray(grid2[0], np.int32), np.ascontiguousarray(grid2[1], np.int32)
            prev_matched = np.array(prev_matched, np.float32).reshape(kps1_un.shape[0], kps1_un.shape[1], 2).copy()
            matches12 = np.zeros(kps1_un.shape[:2], np.int32)
            nmatches = np.zeros(kps1_un.shape[0], np.int32)
        else:
            cs, ix = grid2[0], grid2[1]
        P, cap = kps1_un.shape[0], kps1_un.shape[1]
        g = _FrameGrid(_ptr(cs), _ptr(ix), *[float(v) for v in grid2[2:6]])
        self._check(lib().sdorb_search_for_initialization_batch(
            self._h, _ptr(kps1_un), _ptr(desc1), _ptr(n1), _ptr(kps2_un), _ptr(desc2), _ptr(n2), C.byref(g), P, cap,
            _ptr(prev_matched), int(window_size), float(nnratio), int(check_orientation), _ptr(matches12), _ptr(nmatches),
            MEM_DEVICE if device else MEM_HOST, C.c_void_p(stream) if stream else None))
        return nmatches, matches12, prev_matched

    def search_by_projection_batch(self, kps_last, kps_last_un, proj, flags_last, desc_mp, n_last, kps_cur_un, desc_cur,
                                   u_right_cur, occupied_cur, n_cur, grid_cur, scale_factors, bounds, th, mbf, mode,
                                   check_orientation=True, orb_dist=0):
        """Returns (nmatches[P], assigned[P, cap]); see sdorb_search_by_projection_batch in include/sdorb.h."""
        kl, klu, kc = np.ascontiguousarray(kps_last), np.ascontiguousarray(kps_last_un), np.ascontiguousarray(kps_cur_un)
        P, cap = kl.shape[0], kl.shape[1]
        keep = [np.ascontiguousarray(proj, np.float32), np.ascontiguousarray(flags_last, np.uint8),
                np.ascontiguousarray(desc_mp, np.uint8), np.ascontiguousarray(n_last, np.int32),
                np.ascontiguousarray(desc_cur, np.uint8), np.ascontiguousarray(u_right_cur, np.float32),
                np.ascontiguousarray(occupied_cur, np.uint8), np.ascontiguousarray(n_cur, np.int32),
                np.ascontiguousarray(grid_cur[0], np.int32), np.ascontiguousarray(grid_cur[1], np.int32),
                np.ascontiguousarray(scale_factors, np.float32)]
        q = _ProjectionSearch()
        q.kps_last, q.kps_last_un, q.kps_cur_un = _ptr(kl), _ptr(klu), _ptr(kc)
        q.proj, q.flags_last, q.desc_mp, q.n_last = _ptr(keep[0]), _ptr(keep[1]), _ptr(keep[2]), _ptr(keep[3])
        q.desc_cur, q.u_right_cur, q.occupied_cur, q.n_cur = _ptr(keep[4]), _ptr(keep[5]), _ptr(keep[6]), _ptr(keep[7])
        q.grid_cur = _FrameGrid(_ptr(keep[8]), _ptr(keep[9]), *[float(v) for v in grid_cur[2:6]])
        q.scale_factors, q.nlevels = _ptr(keep[10]), len(keep[10])
        q.bounds = (C.c_float * 4)(*[float(v) for v in bounds])
        q.th, q.mbf, q.mode, q.check_orientation = float(th), float(mbf), int(mode), int(check_orientation)
        q.orb_dist = int(orb_dist)
        asg = np.zeros((P, cap), np.int32)
        nm = np.zeros(P, np.int32)
        self._check(lib().sdorb_search_by_projection_batch(self._h, C.byref(q), P, cap, _ptr(asg), _ptr(nm), MEM_HOST, None))
        return nm, asg

    def search_map_points_batch(self, proj, view_cos, level, flags, desc_mp, n_mp, kps_un, desc, u_right, occupied, n_frame, grid,
                                scale_factors, th, nnratio=0.8):
        """ORBmatcher::SearchByProjection(Frame, vpMapPoints, th) for a batch of frames (host arrays; see include/sdorb.h):
        returns (nmatches[F], assigned[F, cap])."""
        keep = [np.ascontiguousarray(proj, np.float32), np.ascontiguousarray(view_cos, np.float32), np.ascontiguousarray(level, np.int32),
                np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc_mp, np.uint8), np.ascontiguousarray(n_mp, np.int32),
                np.ascontiguousarray(kps_un), np.ascontiguousarray(desc, np.uint8), np.ascontiguousarray(u_right, np.float32),
                np.ascontiguousarray(occupied, np.uint8), np.ascontiguousarray(n_frame, np.int32),
                np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32),
                np.ascontiguousarray(scale_factors, np.float32)]
        F, capmp, cap = keep[1].shape[0], keep[1].shape[1], keep[6].shape[1]
        q = _MapPointSearch(*[_ptr(k) for k in keep[:11]], _FrameGrid(_ptr(keep[11]), _ptr(keep[12]), *[float(v) for v in grid[2:6]]),
                            _ptr(keep[13]), len(keep[13]), float(th), float(nnratio), 0)
        asg = np.zeros((F, cap), np.int32)
        nm = np.zeros(F, np.int32)
        self._check(lib().sdorb_search_map_points_batch(self._h, C.byref(q), F, capmp, cap, _ptr(asg), _ptr(nm), MEM_HOST, None))
        return nm, asg

    def search_by_projection_sim3_batch(self, proj, level, flags, desc_mp, n_mp, kps_un, desc, matched, n_kf, grid, scale_factors, th):
        """ORBmatcher::SearchByProjection(KeyFrame*, Scw, vpPoints, vpMatched, th) (loop closing) for a batch of keyframes (host
        arrays; sdorb_map_point_search::sim3_form): returns (nmatches[F], assigned[F, cap])."""
        keep = [np.ascontiguousarray(proj, np.float32), None, np.ascontiguousarray(level, np.int32),
                np.ascontiguousarray(flags, np.uint8), np.ascontiguousarray(desc_mp, np.uint8), np.ascontiguousarray(n_mp, np.int32),
                np.ascontiguousarray(kps_un), np.ascontiguousarray(desc, np.uint8), None,
                np.ascontiguousarray(matched, np.uint8), np.ascontiguousarray(n_kf, np.int32),
                np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32),
                np.ascontiguousarray(scale_factors, np.float32)]
        F, capmp, cap = keep[2].shape[0], keep[2].shape[1], keep[6].shape[1]
        q = _MapPointSearch(*[_ptr(k) if k is not None else None for k in keep[:11]],
                            _FrameGrid(_ptr(keep[11]), _ptr(keep[12]), *[float(v) for v in grid[2:6]]),
                            _ptr(keep[13]), len(keep[13]), float(int(th)), 0.0, 1)
        asg = np.zeros((F, cap), np.int32)
        nm = np.zeros(F, np.int32)
        self._check(lib().sdorb_search_map_points_batch(self._h, C.byref(q), F, capmp, cap, _ptr(asg), _ptr(nm), MEM_HOST, None))
        return nm, asg

    def search_by_points_batch(self, kps1_un, desc1, valid1, n1, kps2_un, desc2, valid2, n2, nnratio=0.75, check_orientation=True):
        """ORBmatcher::SearchByPoints for a batch of keyframe pairs (host arrays): returns (nmatches[P], matches12[P, cap])."""
        k1, k2 = np.ascontiguousarray(kps1_un), np.ascontiguousarray(kps2_un)
        d1, d2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
        v1, v2 = np.ascontiguousarray(valid1, np.uint8), np.ascontiguousarray(valid2, np.uint8)
        n1, n2 = np.ascontiguousarray(n1, np.int32), np.ascontiguousarray(n2, np.int32)
        P, cap = k1.shape[0], k1.shape[1]
        m12 = np.zeros((P, cap), np.int32)
        nm = np.zeros(P, np.int32)
        self._check(lib().sdorb_search_by_points_batch(self._h, _ptr(k1), _ptr(d1), _ptr(v1), _ptr(n1), _ptr(k2), _ptr(d2), _ptr(v2),
                                                        _ptr(n2), P, cap, float(nnratio), int(check_orientation), _ptr(m12), _ptr(nm),
                                                        MEM_HOST, None))
        return nm, m12

    @staticmethod
    def _fuse_query(proj, level, flags, desc_mp, n_mp, kps_un, desc, u_right, grid, scale_factors, inv_level_sigma2, th, check, th_dist,
                    device=False):
        hostf = lambda a: np.ascontiguousarray(a, np.float32) if a is not None else None
        if device:  # torch tensors on the handle's GPU, taken as they are; the level tables stay host arrays
            keep = [proj, level, flags, desc_mp, n_mp, kps_un, desc, u_right, grid[0], grid[1]]
        else:
            keep = [np.ascontiguousarray(proj, np.float32), np.ascontiguousarray(level, np.int32), np.ascontiguousarray(flags, np.uint8),
                    np.ascontiguousarray(desc_mp, np.uint8), np.ascontiguousarray(n_mp, np.int32), np.ascontiguousarray(kps_un),
                    np.ascontiguousarray(desc, np.uint8), hostf(u_right),
                    np.ascontiguousarray(grid[0], np.int32), np.ascontiguousarray(grid[1], np.int32)]
        keep += [hostf(scale_factors), hostf(inv_level_sigma2)]
        ptr = lambda a: _ptr(a) if a is not None else None
        q = _FuseSearch(*[ptr(k) for k in keep[:8]], _FrameGrid(_ptr(keep[8]), _ptr(keep[9]), *[float(v) for v in grid[2:6]]),
                        _ptr(keep[10]), ptr(keep[11]), len(keep[10]), float(th), int(check), int(th_dist))
        return q, keep

    def fuse_search_batch(self, proj, level, flags, desc_mp, n_mp, kps_un, desc, u_right, grid, scale_factors, inv_level_sigma2, th,
                          check_reprojection=True, th_dist=50, best_idx=None, best_dist=None, device=False, stream=None):
        """The keypoint search of ORBmatcher::Fuse for a batch of keyframes: (best_idx[F, cap_mp], best_dist[F, cap_mp]).
        check_reprojection=False: the search of the Sim3 overload (u_right / inv_level_sigma2 may be None).  Host arrays, or with
        device=True torch tensors on the handle's GPU (best_idx / best_dist are then the caller's int32 tensors)."""
        q, keep = self._fuse_query(proj, level, flags, desc_mp, n_mp, kps_un, desc, u_right, grid, scale_factors, inv_level_sigma2, th,
                                   check_reprojection, th_dist, device)
        F, capmp, cap = keep[1].shape[0], keep[1].shape[1], keep[5].shape[1]
        if not device:
            best_idx, best_dist = np.zeros((F, capmp), np.int32), np.zeros((F, capmp), np.int32)
        self._check(lib().sdorb_fuse_search_batch(self._h, C.byref(q), F, capmp, cap, _ptr(best_idx), _ptr(best_dist),
                                                   MEM_DEVICE if device else MEM_HOST, C.c_void_p(stream) if stream else None))
        return best_idx, best_dist

    def search_by_sim3_batch(self, side1, side2, scale_factors, th, out=None, device=False, stream=None):
        """ORBmatcher::SearchBySim3 for a batch of keyframe pairs.  side = (proj[P, cap, 3], level[P, cap], flags[P, cap],
        desc_mp[P, cap, 32], n[P], kps_un[P, cap], desc[P, cap, 32], grid) of that keyframe: its map points projected into the other
        keyframe, and its own keypoints / descriptors / grid.  Returns (nfound[P], matches12[P, cap], match1, match2).  Host arrays,
        or with device=True torch tensors and out = (nfound, matches12, match1, match2) int32 tensors of the caller."""
        p1, l1, f1, m1, n1, k1, d1, g1 = side1
        p2, l2, f2, m2, n2, k2, d2, g2 = side2
        q12, keep12 = self._fuse_query(p1, l1, f1, m1, n1, k2, d2, None, g2, scale_factors, None, th, 0, 100, device)
        q21, keep21 = self._fuse_query(p2, l2, f2, m2, n2, k1, d1, None, g1, scale_factors, None, th, 0, 100, device)
        P, cap = keep12[1].shape
        assert tuple(keep21[1].shape) == (P, cap) and tuple(keep12[5].shape[:2]) == (P, cap) and tuple(keep21[5].shape[:2]) == (P, cap)
        if device:
            nf, o12, o1, o2 = out
        else:
            o1, o2, o12, nf = np.zeros((P, cap), np.int32), np.zeros((P, cap), np.int32), np.zeros((P, cap), np.int32), np.zeros(P, np.int32)
        self._check(lib().sdorb_search_by_sim3_batch(self._h, C.byref(q12), C.byref(q21), P, cap, _ptr(o1), _ptr(o2), _ptr(o12), _ptr(nf),
                                                     MEM_DEVICE if device else MEM_HOST, C.c_void_p(stream) if stream else None))
        return nf, o12, o1, o2

    def search_for_triangulation_batch(self, kps1_un, desc1, has_mp1, u_right1, n1, kps2_un, desc2, has_mp2, u_right2, n2, F12, epipole,
                                       scale_factors, level_sigma2, check_orientation=True, matches12=None, nmatches=None,
                                       device=False, stream=None):
        """ORBmatcher::SearchForTriangulation for many keyframe pairs (see include/sdorb.h): returns (nmatches[P],
        matches12[P, cap]).  Host arrays, or with device=True torch tensors on the handle's GPU (scale_factors / level_sigma2
        stay host arrays; matches12 / nmatches are the caller's int32 tensors)."""
        sf, s2 = np.ascontiguousarray(scale_factors, np.float32), np.ascontiguousarray(level_sigma2, np.float32)
        if not device:
            kps1_un, kps2_un = np.ascontiguousarray(kps1_un), np.ascontiguousarray(kps2_un)
            desc1, desc2 = np.ascontiguousarray(desc1, np.uint8), np.ascontiguousarray(desc2, np.uint8)
            has_mp1, has_mp2 = np.ascontiguousarray(has_mp1, np.uint8), np.ascontiguousarray(has_mp2, np.uint8)
            u_right1, u_right2 = np.ascontiguousarray(u_right1, np.float32), np.ascontiguousarray(u_right2, np.float32)
            n1, n2 = np.ascontiguousarray(n1, np.int32), np.ascontiguousarray(n2, np.int32)
            F12, epipole = np.ascontiguousarray(F12, np.float64), np.ascontiguousarray(epipole, np.float32)
            matches12 = np.zeros(kps1_un.shape[:2], np.int32)
            nmatches = np.zeros(kps1_un.shape[0], np.int32)
        P, cap = kps1_un.shape[0], kps1_un.shape[1]
        q = _TriangulationSearch(_ptr(kps1_un), _ptr(desc1), _ptr(has_mp1), _ptr(u_right1), _ptr(n1), _ptr(kps2_un), _ptr(desc2),
                                 _ptr(has_mp2), _ptr(u_right2), _ptr(n2), _ptr(F12), _ptr(epipole), _ptr(sf), _ptr(s2), len(sf),
                                 int(check_orientation))
        self._check(lib().sdorb_search_for_triangulation_batch(self._h, C.byref(q), P, cap, _ptr(matches12), _ptr(nmatches),
                                                                MEM_DEVICE if device else MEM_HOST,
                                                                C.c_void_p(stream) if stream else None))
        return nmatches, matches12

    # ---- instrumentation
    def set_profiling(self, on):
        self._check(lib().sdorb_set_profiling(self._h, int(on)))

    def stage_times(self, reset=True):
        ms = np.zeros(len(STAGES), np.float64)
        launches = np.zeros(len(STAGES), np.int64)
        self._check(lib().sdorb_get_stage_times(self._h, _ptr(ms), _ptr(launches), int(reset)))
        return dict(zip(STAGES, ms.tolist())), dict(zip(STAGES, launches.tolist()))

    def guard_check(self):
        """Guarded run (SDORB_GUARD=1): overwritten guard bytes over all live device buffers (0 = intact); None when not guarded."""
        n = int(lib().sdorb_debug_guard_check(self._h))
        if n == -1000:
            return None
        if n < 0:
            self._check(n)
        if n:
            raise SdorbError(-6, lib().sdorb_last_cuda_error(self._h).decode())
        return n

    def pipe_probe(self, pipe):
        """Measured peak of an execution pipe (0 POPC, 1 VIMNMX3.U16x2 on the ALU pipe, 2 PRMT): (warp-instructions / s,
        warp-instructions / clk / SM)."""
        a, b = C.c_double(0), C.c_double(0)
        self._check(lib().sdorb_debug_pipe_probe(self._h, pipe, C.byref(a), C.byref(b)))
        return a.value, b.value

    def kernel_launches(self):
        return int(lib().sdorb_kernel_launches(self._h))

    def debug_nth_element(self, entries, nth):
        """std::nth_element(e, e+nth, end, response >) as the selection kernel does it; returns the permuted copy."""
        e = np.ascontiguousarray(entries, np.uint32).copy()
        self._check(lib().sdorb_debug_nth_element(self._h, _ptr(e), len(e), nth))
        return e

    def debug_read(self, what, frame, level, nbytes, dtype=np.uint8):
        buf = np.zeros(max(nbytes, 4), np.uint8)
        n = lib().sdorb_debug_read(self._h, what, frame, level, _ptr(buf), buf.size)
        if n < 0:
            self._check(int(n))
        return buf[:n].view(dtype)


class ORBmatcher:
    """The static distance of src/ORBmatcher.h:43 plus its batched forms, bound to one extractor handle."""
    TH_HIGH, TH_LOW, HISTO_LENGTH = 100, 50, 30  # src/ORBmatcher.cc:36-38

    def __init__(self, extractor, nnratio=0.6):
        self.ex, self.nnratio = extractor, nnratio

    def DescriptorDistance(self, a, b):
        return int(self.ex.hamming_matrix(np.asarray(a).reshape(1, 32), np.asarray(b).reshape(1, 32))[0, 0])

    def match(self, descA, nA, descB, nB, greedy=False, **kw):
        return self.ex.match_batch(descA, nA, descB, nB, ratio=kw.pop("ratio", self.nnratio), th_low=kw.pop("th_low", self.TH_LOW),
                                   greedy=greedy, **kw)
