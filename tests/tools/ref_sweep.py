#!/usr/bin/env python3
"""Oracle-versus-reference sweep on the CPU: the restatement (oracle/libsdorb_oracle.so) against the reference's own sources
compiled unmodified (oracle/_ref/libsdorb_ref.so, oracle/ref_build/Makefile) on thousands of synthetic frames -- the same frame
generators, seeds and shapes tests/tools/sweep.py feeds to the GPU, so   GPU == oracle (tests/tools/sweep.py)   and   oracle == reference
(this file)   meet on identical inputs.  Prints one line per configuration; exit code 1 on any mismatch.
Usage: python tests/tools/ref_sweep.py [scale]   (scale 1.0 = 4096 C3 frames; log committed under profiles/)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as orc  # noqa: E402
from oracle import ref_binding as ref  # noqa: E402
from sdslam_b200 import synth  # noqa: E402

CONFIGS = [  # label, generator, width, height, params, frames  (the reference-mode rows of tests/tools/sweep.py)
    ("C3 smooth_noise", "smooth_noise", 640, 480, (1000, 1.2, 8, 20), 4096),
    ("C3 rects", "rects", 640, 480, (1000, 1.2, 8, 20), 1024),
    ("C2 euroc", "smooth_noise", 752, 480, (1000, 1.2, 8, 20), 512),
    ("C0 defaults", "smooth_noise", 640, 480, (1000, 2.0, 5, 20), 512),
    ("ini 2000", "smooth_noise", 640, 480, (2000, 1.2, 8, 20), 256),
    ("C5 1080p", "smooth_noise", 1920, 1080, (4000, 1.2, 12, 20), 64),
]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    threads = os.cpu_count() or 8
    native = ref._native_runs_here()
    bad_total = 0
    for label, kind, w, h, params, n in CONFIGS:
        n = max(8, int(n * scale))
        bad = kp_total = 0
        t_o = t_r = 0.0
        for f0 in range(0, n, 256):  # chunks bound the memory of the generator
            m = min(256, n - f0)
            imgs = synth.frames(m, w, h, kind, start=10_000 + f0)
            t = time.time()
            ok, od, oc = orc.Extractor(*params).extract_many(imgs, nthreads=threads)
            t_o += time.time() - t
            t = time.time()
            rk, rd, rc = ref.Extractor(*params, native=native).extract_many(imgs, nthreads=threads)
            t_r += time.time() - t
            for f in range(m):
                c = int(rc[f])
                if oc[f] != c or ok[f, :c].tobytes() != rk[f, :c].tobytes() or od[f, :c].tobytes() != rd[f, :c].tobytes():
                    bad += 1
            kp_total += int(rc.sum())
        bad_total += bad
        print("%-16s %4d frames %dx%d %s: frames where oracle != reference: %d  (keypoints %d; oracle %.0fs, reference%s %.0fs on %d threads)" % (
            label, n, w, h, params, bad, kp_total, t_o, " (-march=native)" if native else "", t_r, threads), flush=True)
    print("REF SWEEP", "OK" if bad_total == 0 else "FAILED (%d frames)" % bad_total)
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
