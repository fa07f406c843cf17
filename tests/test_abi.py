"""The C-ABI library on a machine without a GPU: it loads, exports every symbol include/sdorb.h declares, its
host-only helpers agree with the oracle, and device entry points fail loudly (no CPU fallback).  CPU only."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import binding as orc
from sdslam_b200 import api

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "sdorb.h")).read()
    return sorted(set(re.findall(r"SDORB_API[^;(]*?\b(sdorb_\w+)\s*\(", txt)))


def test_header_declares_the_boundary():
    syms = declared_symbols()
    for must in ("sdorb_create", "sdorb_destroy", "sdorb_extract", "sdorb_extract_batch", "sdorb_match_batch",
                 "sdorb_hamming_matrix", "sdorb_get_tables", "sdorb_strerror"):
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol():
    L = api.lib()
    for s in declared_symbols():
        assert hasattr(L, s), "libsdorb.so does not export " + s


def test_library_does_not_link_the_oracle():
    import subprocess
    out = subprocess.check_output(["ldd", api._build.SO], text=True)
    assert "oracle" not in out
    syms = subprocess.check_output(["nm", "-D", "--defined-only", api._build.SO], text=True)
    assert "orc_" not in syms


def test_strerror():
    L = api.lib()
    assert L.sdorb_strerror(0) == b"ok"
    assert len({L.sdorb_strerror(-c) for c in range(1, 8)}) == 7
    assert L.sdorb_strerror(-99) == b"unknown error"


def test_keypoint_layout_is_cv_keypoint():
    assert api.KP_DTYPE.itemsize == 28 and api.MATCH_DTYPE.itemsize == 16
    assert [api.KP_DTYPE.fields[n][1] for n in api.KP_DTYPE.names] == [0, 4, 8, 12, 16, 20, 24]


@pytest.mark.parametrize("params", [(1000, 1.2, 8, 20), (1000, 2.0, 5, 20), (2000, 1.2, 8, 20), (4000, 1.2, 12, 20),
                                    (500, 1.5, 6, 7), (1, 1.2, 8, 20), (0, 1.2, 3, 20), (300, 1.2, 1, 20)])
def test_host_tables_equal_oracle(params):
    """Constructor tables, /root/reference/src/ORBextractor.cc:406-457."""
    t = api.host_tables(*params[:3])
    o = orc.Extractor(*params).tables()
    for k in ("scale", "inv_scale", "sigma2", "inv_sigma2", "n_per_level", "umax"):
        assert t[k].tobytes() == o[k].tobytes(), k


@pytest.mark.parametrize("params", [(1000, 1.2, 8, 20), (1000, 2.0, 5, 20), (4000, 1.2, 12, 20), (300, 1.2, 4, 10)])
@pytest.mark.parametrize("wh", [(640, 480), (752, 480), (1920, 1080), (800, 450), (200, 300), (321, 243), (1241, 376)])
def test_host_level_geometry_equals_oracle(params, wh):
    """Level sizes and cell grids, src/ORBextractor.cc:469-488, 683."""
    try:
        g = api.host_level_geometry(*params, *wh)
    except api.SdorbError as e:
        assert e.code == -4
        with pytest.raises(RuntimeError):
            orc.Extractor(*params).extract(np.zeros((wh[1], wh[0]), np.uint8))
        return
    o = orc.Extractor(*params).geometry(*wh)
    for name in g.dtype.names:
        if name in ("cell_w", "cell_h", "n_features_cell"):
            m = (o["level_cols"] > 0) & (o["level_rows"] > 0)
            assert np.array_equal(g[name][m], o[name][m]), name
        else:
            assert np.array_equal(g[name], o[name]), name


def test_host_geometry_errors():
    with pytest.raises(api.SdorbError) as e:
        api.host_level_geometry(1000, 1.2, 8, 20, 40, 30)
    assert e.value.code == -4  # the reference throws cv::Exception for this size
    with pytest.raises(api.SdorbError):
        api.host_tables(1000, 0.0, 8)
    with pytest.raises(api.SdorbError):
        api.host_tables(1000, 1.2, 0)


def test_fill_border_reflect101_matches_oracle():
    rng = np.random.default_rng(3)
    for w, h in ((50, 37), (20, 20), (64, 21)):
        inner = rng.integers(0, 256, (h, w), dtype=np.uint8)
        buf = np.zeros((h + 38, w + 38), np.uint8)
        buf[19:19 + h, 19:19 + w] = inner
        view = buf[19:, 19:]
        api.lib().sdorb_fill_border_reflect101(view.ctypes.data_as(C.c_void_p), w, h, buf.strides[0], 19)
        assert np.array_equal(buf, orc.border_reflect101(inner, 19))


def test_bad_arguments_rejected_without_touching_cuda():
    L = api.lib()
    h = C.c_void_p()
    assert L.sdorb_create(None, C.byref(h)) == -1
    p = api._Params(1000, 1.2, 0, 20, -1, 0, 640, 480, 4)  # nlevels 0
    assert L.sdorb_create(C.byref(p), C.byref(h)) == -1 and not h.value
    p = api._Params(1000, 1.2, 8, 20, 7, 0, 640, 480, 0)   # max_batch 0 (the ORB-SLAM2-style mode itself is accepted: row f1)
    assert L.sdorb_create(C.byref(p), C.byref(h)) == -1 and not h.value
    assert L.sdorb_extract(None, None, 0, 0, 0, None, None, 0, None, None) == -1
    assert L.sdorb_kernel_launches(None) == 0


@pytest.mark.skipif(_has_gpu(), reason="checks the no-GPU failure mode")
def test_no_cpu_fallback_without_gpu():
    """On a box without a CUDA device the product path must fail loudly, not compute on the CPU."""
    with pytest.raises(api.SdorbError) as e:
        api.ORBextractor(1000, 1.2, 8, 20)
    assert e.value.code in (-3, -5)


def test_host_image_bounds_equal_oracle():
    """Frame::ComputeImageBounds (src/Frame.cc:368-397) runs on the host, no GPU involved."""
    from test_oracle_primitives import CAMERAS
    for cam, (K4, dist) in CAMERAS.items():
        for cols, rows in ((640, 480), (752, 480)):
            assert api.host_image_bounds(cols, rows, K4, dist).tobytes() == orc.image_bounds(cols, rows, K4, dist).tobytes(), cam


def test_ctypes_mirrors_match_the_header_layout(tmp_path):
    """Every struct of include/sdorb.h against its ctypes / numpy mirror in sdslam_b200/api.py: sizeof and the offset of every
    field, taken from a C program compiled against the header (gcc, the C ABI of this platform)."""
    import ctypes as C
    import subprocess
    import numpy as np
    from sdslam_b200 import api
    pairs = {"sdorb_params": api._Params, "sdorb_pyr_view": api._PyrView, "sdorb_frame_grid": api._FrameGrid,
             "sdorb_projection_search": api._ProjectionSearch, "sdorb_map_point_search": api._MapPointSearch,
             "sdorb_fuse_search": api._FuseSearch, "sdorb_triangulation_search": api._TriangulationSearch}
    dtypes = {"sdorb_keypoint": api.KP_DTYPE, "sdorb_match": api.MATCH_DTYPE, "sdorb_level_geom": api.GEOM_DTYPE}
    lines = []
    for cname, cls in pairs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    for cname, dt in dtypes.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname in dt.names:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "sdorb.h"\nint main(void) {\n%s\nreturn 0; }\n' % "\n".join(lines))
    exe = str(tmp_path / "layout")
    subprocess.check_call(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), "-o", exe, str(src)])
    got = dict(l.split() for l in subprocess.check_output([exe], text=True).splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got["%s.%s" % (cname, fname)]) == getattr(cls, fname).offset, "%s.%s" % (cname, fname)
    for cname, dt in dtypes.items():
        assert int(got[cname]) == dt.itemsize, cname
        for fname in dt.names:
            assert int(got["%s.%s" % (cname, fname)]) == dt.fields[fname][1], "%s.%s" % (cname, fname)
