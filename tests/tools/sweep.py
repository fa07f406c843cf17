#!/usr/bin/env python3
"""Large parity sweep: the CUDA path against the CPU oracle on thousands of synthetic frames (SURVEY.md section 4:
"GPU end-to-end ... over >= 4096 synthetic frames").  Prints one line per configuration; exit code 1 on any mismatch."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import binding as orc  # noqa: E402
from sdslam_b200 import api, synth  # noqa: E402

CONFIGS = [  # label, generator, width, height, params, frames
    ("C3 smooth_noise", "smooth_noise", 640, 480, (1000, 1.2, 8, 20), 4096),
    ("C3 rects", "rects", 640, 480, (1000, 1.2, 8, 20), 1024),
    ("C2 euroc", "smooth_noise", 752, 480, (1000, 1.2, 8, 20), 512),
    ("C0 defaults", "smooth_noise", 640, 480, (1000, 2.0, 5, 20), 512),
    ("ini 2000", "smooth_noise", 640, 480, (2000, 1.2, 8, 20), 256),
    ("C5 1080p", "smooth_noise", 1920, 1080, (4000, 1.2, 12, 20), 64),
    # ORB-SLAM2-style mode (row f1): the fifth parameter is minThFAST
    ("f1 smooth_noise", "smooth_noise", 640, 480, (1000, 1.2, 8, 20, 7), 1024),
    ("f1 rects", "rects", 640, 480, (1000, 1.2, 8, 20, 7), 512),
    ("f1 euroc 2000", "smooth_noise", 752, 480, (2000, 1.2, 8, 20, 7), 256),
]


def main():
    scale = float(sys.argv[1]) if len(sys.argv) > 1 else 1.0
    threads = os.cpu_count() or 8
    bad_total = 0
    for label, kind, w, h, params, n in CONFIGS:
        n = max(8, int(n * scale))
        t = time.time()
        imgs = synth.frames(n, w, h, kind, start=10_000)
        t_gen = time.time() - t
        min_th = params[4] if len(params) > 4 else None
        ex = api.ORBextractor(*params[:4], minThFAST=min_th, max_width=w, max_height=h, max_batch=min(n, 256))
        t = time.time()
        gk, gd, gc = ex.extract_batch_host(imgs)
        t_gpu = time.time() - t
        ex.close()
        t = time.time()
        ok, od, oc = orc.Extractor(*params[:4], min_th_fast=min_th).extract_many(imgs, nthreads=threads)
        t_cpu = time.time() - t
        cap = min(gk.shape[1], ok.shape[1])
        bad = 0
        for f in range(n):
            c = int(oc[f])
            if gc[f] != c or gk[f, :c].tobytes() != ok[f, :c].tobytes() or gd[f, :c].tobytes() != od[f, :c].tobytes():
                bad += 1
        bad_total += bad
        print("%-16s %4d frames %dx%d %s: mismatching frames %d  (keypoints %d, gen %.0fs gpu %.2fs oracle %.0fs on %d threads)" % (
            label, n, w, h, params, bad, int(oc.sum()), t_gen, t_gpu, t_cpu, threads), flush=True)
    print("SWEEP", "OK" if bad_total == 0 else "FAILED (%d frames)" % bad_total)
    return 1 if bad_total else 0


if __name__ == "__main__":
    sys.exit(main())
