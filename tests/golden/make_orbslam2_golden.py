#!/usr/bin/env python3
"""Generate tests/golden/orbslam2/*.npz for the ORB-SLAM2-style mode with the independent
cv2 pipeline (tests/cv2_pipeline.py: real cv2 FAST per 30-pixel cell with the ini / min threshold fallback, the
pass-based DistributeOctTree, real cv2 resize / GaussianBlur / fastAtan2).  Run in the build container:
    python tests/golden/make_orbslam2_golden.py
params = (nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2_pipeline as cvp  # noqa: E402
from sdslam_b200 import synth  # noqa: E402


def flat_with_texture(w, h):
    """Mostly low-contrast (only minThFAST finds corners there) with a few high-contrast patches (iniThFAST cells)."""
    rng = np.random.default_rng(5)
    img = (128 + rng.normal(0, 4.5, (h, w))).clip(0, 255)
    for _ in range(12):
        x, y = int(rng.integers(0, w - 50)), int(rng.integers(0, h - 50))
        img[y:y + 40, x:x + 40] = rng.integers(0, 256, (40, 40))
    return img.astype(np.uint8)


CASES = {
    "os2_smooth_640x480": (synth.smooth_noise(0), (1000, 1.2, 8, 20, 7)),
    "os2_rects_640x480": (synth.rects(0), (1000, 1.2, 8, 20, 7)),
    "os2_fallback_480x360": (flat_with_texture(480, 360), (800, 1.2, 6, 20, 7)),
    "os2_wide_620x188": (synth.smooth_noise(9, 620, 188), (600, 1.2, 4, 20, 7)),   # KITTI aspect: 3-4 initial nodes
    "os2_few_320x240": (synth.rects(4, 320, 240), (2000, 1.2, 5, 40, 30)),         # fewer keypoints than N: every node ends single
    "os2_tiny_n_320x240": (synth.smooth_noise(6, 320, 240), (20, 1.5, 3, 20, 7)),
}

if __name__ == "__main__":
    os.makedirs(os.path.join(HERE, "orbslam2"), exist_ok=True)
    for name, (img, params) in CASES.items():
        k, d, pyr = cvp.extract(img, *params[:4], min_th_fast=params[4])
        np.savez_compressed(os.path.join(HERE, "orbslam2", name + ".npz"), image=img, params=np.array(params, np.float64), kps=k,
                            desc=d)
        print(name, img.shape, params, len(k), np.bincount(k["octave"]) if len(k) else "")
