#!/usr/bin/env python3
"""Generate tests/golden/*.npz FROM THE REFERENCE ITSELF: /root/reference/src/ORBextractor.cc compiled unmodified into
oracle/_ref/libsdorb_ref.so (oracle/ref_build/Makefile, cv:: surface = oracle/ref_compat).  Every case is cross-checked while it
is generated against the independent pipeline assembled from real OpenCV 4.13 calls (tests/cv2_pipeline.py): the file is only
written when both agree byte for byte.  Run in the build container:  python tests/golden/make_golden.py
Each file holds the input image, the parameters, keypoints, descriptors, a SHA-256 per pyramid level (and per padded level
buffer) and the generator tag.  The oracle is checked against these on CPU; the CUDA library is checked against them on the GPU
box (where neither /root/reference nor this script is needed).
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import cv2_pipeline as cvp  # noqa: E402
from oracle import ref_binding as ref  # noqa: E402
from sdslam_b200 import synth  # noqa: E402


def checker(w, h, s=9):
    y, x = np.mgrid[0:h, 0:w]
    return (((x // s + y // s) & 1) * 200 + 20).astype(np.uint8)


CASES = {
    "wide_kitti_1241x376": (synth.smooth_noise(8, 1241, 376), (2000, 1.2, 8, 20)),
    "noise_th7_400x300": (np.random.default_rng(48).integers(0, 256, (300, 400), dtype=np.uint8), (1500, 1.2, 8, 7)),
    "ties_binary_260x200": ((np.random.default_rng(7).integers(0, 2, (200, 260)) * 255).astype(np.uint8), (300, 1.2, 4, 20)),
    "c1_smooth_640x480": (synth.smooth_noise(0), (1000, 1.2, 8, 20)),
    "c1_rects_640x480": (synth.rects(0), (1000, 1.2, 8, 20)),
    "c0_default_640x480": (synth.smooth_noise(1), (1000, 2.0, 5, 20)),
    "c2_smooth_752x480": (synth.smooth_noise(2, 752, 480), (1000, 1.2, 8, 20)),
    "ini_2000_320x240": (synth.smooth_noise(3, 320, 240), (2000, 1.2, 8, 20)),
    "portrait_200x300": (synth.smooth_noise(5, 200, 300), (300, 1.2, 4, 10)),
    "checker_320x240": (checker(320, 240), (500, 1.2, 8, 20)),
    "constant_160x120": (np.full((120, 160), 77, np.uint8), (500, 1.2, 8, 20)),
    "c5_small_800x450": (synth.smooth_noise(7, 800, 450), (4000, 1.2, 12, 20)),
}

if __name__ == "__main__":
    for name, (img, params) in CASES.items():
        k, d, pyr, padded = ref.Extractor(*params).extract(img, pyramid=True)  # the reference's own compiled text
        ck, cd, cpyr = cvp.extract(img, *params)                               # real OpenCV 4.13 calls, independent control flow
        assert k.tobytes() == ck.tobytes() and np.array_equal(d, cd.reshape(-1, 32)), name
        assert all(np.array_equal(a, b) for a, b in zip(pyr, cpyr)) and len(pyr) == len(cpyr), name
        np.savez_compressed(os.path.join(HERE, name + ".npz"), image=img, params=np.array(params, np.float64),
                            kps=k, desc=d, pyramid_shape=np.array([p.shape for p in pyr], np.int32),
                            pyramid_sha256=np.array([hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in pyr]),
                            padded_sha256=np.array([hashlib.sha256(np.ascontiguousarray(p).tobytes()).hexdigest() for p in padded]),
                            generator=np.array("reference: /root/reference/src/ORBextractor.cc via oracle/_ref (cross-checked with cv2 %s)" % cvp.cv2.__version__))
        print(name, img.shape, params, len(k))
    # Hamming golden: distances by numpy bit counting (independent of the SWAR restatement)
    rng = np.random.default_rng(77)
    A = rng.integers(0, 256, (96, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (80, 32), dtype=np.uint8)
    B[5] = A[7]
    B[9] = A[7]          # exact duplicate rows: first index must win
    B[11] = A[3] ^ np.eye(32, dtype=np.uint8)[0] * 1
    dist = np.unpackbits(A[:, None, :] ^ B[None, :, :], axis=2).sum(axis=2).astype(np.uint16)
    np.savez_compressed(os.path.join(HERE, "hamming_96x80.npz"), A=A, B=B, dist=dist)
    print("hamming", dist.shape)
