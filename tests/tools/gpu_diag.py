#!/usr/bin/env python3
"""Stage-by-stage comparison of libsdorb against the CPU oracle on one frame (diagnostics for the GPU box)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as orc  # noqa: E402
from sdslam_b200 import api, synth  # noqa: E402


def diag(img, params, label):
    h, w = img.shape
    o = orc.Extractor(*params)
    ok, od, st = o.extract(img, dump=True)
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=4)
    t = time.time()
    gk, gd, pyr = ex(img)
    dt = time.time() - t
    geom = st["geometry"]
    print("== %s %dx%d params=%s  oracle n=%d gpu n=%d (%.1f ms)" % (label, w, h, params, len(ok), len(gk), dt * 1e3))
    off = 0
    cell_off = 0
    raw_off = 0
    for l, g in enumerate(geom):
        lw, lh = int(g["width"]), int(g["height"])
        n = lw * lh
        ref = st["pyramid"][off:off + n].reshape(lh, lw)
        got = pyr[l]
        pm = int((ref != got).sum())
        bl = ex.debug_read(api.DBG_BLURRED_LEVEL, 0, l, n).reshape(lh, lw)
        bm = int((st["blurred"][off:off + n].reshape(lh, lw) != bl).sum())
        nc = max(int(g["level_cols"]), 0) * max(int(g["level_rows"]), 0)
        cc = ex.debug_read(api.DBG_CELL_COUNTS, 0, l, nc * 4, np.int32) if nc else np.zeros(0, np.int32)
        rc = st["raw_cell_count"][cell_off:cell_off + nc]
        sel = ex.debug_read(api.DBG_LEVEL_SELECTED, 0, l, 4 * 8192, np.uint32)
        okl = ok[ok["octave"] == l]
        sf = o.tables()["scale"][l]
        print("  L%d %dx%d pyr_mismatch=%d blur_mismatch=%d cells=%d cellcount_mismatch=%d raw=%d/%d sel=%d/%d" % (
            l, lw, lh, pm, bm, nc, int((cc != rc).sum()) if nc else 0, int(cc.sum()), int(rc.sum()), len(sel), len(okl)))
        off += n
        cell_off += nc
    same_n = len(ok) == len(gk)
    if same_n:
        for f in ok.dtype.names:
            print("   kp.%s mismatches: %d" % (f, int((ok[f] != gk[f]).sum())))
        if len(ok):
            print("   max |angle diff| = %g" % float(np.abs(ok["angle"] - gk["angle"]).max()))
        print("   descriptor row mismatches: %d / %d" % (int((od != gd).any(axis=1).sum()), len(od)))
    return same_n and ok.tobytes() == gk.tobytes() and od.tobytes() == gd.tobytes()


if __name__ == "__main__":
    res = []
    res.append(diag(synth.smooth_noise(0), (1000, 1.2, 8, 20), "C1 smooth"))
    res.append(diag(synth.rects(0), (1000, 1.2, 8, 20), "C1 rects"))
    res.append(diag(synth.smooth_noise(1), (1000, 2.0, 5, 20), "C0 default"))
    res.append(diag(synth.smooth_noise(2, 752, 480), (1000, 1.2, 8, 20), "C2"))
    res.append(diag(synth.smooth_noise(5, 200, 300), (300, 1.2, 4, 10), "portrait"))
    res.append(diag(synth.smooth_noise(7, 800, 450), (4000, 1.2, 12, 20), "C5 small"))
    print("ALL BIT-EXACT" if all(res) else "MISMATCHES: %s" % res)
