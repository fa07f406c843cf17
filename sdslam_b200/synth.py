"""Deterministic synthetic gray frames: the same bytes go to the CPU oracle and the GPU.

"smooth+noise": an 8x-upsampled random field plus Gaussian pixel noise -- dense corners at every pyramid
level, so every per-cell and per-level trim of the extractor fires.  "rects": grey background with random
filled rectangles, lightly blurred -- the sparse-corner regime.
"""
import numpy as np

try:  # cv2 only supplies the bicubic upsample / small blur of the generator; a numpy path stands in without it
    import cv2 as _cv2
except Exception:  # pragma: no cover
    _cv2 = None


def _upsample(base, w, h):
    if _cv2 is not None:
        return _cv2.resize(base, (w, h), interpolation=_cv2.INTER_CUBIC)
    ys = (np.arange(h) + 0.5) * base.shape[0] / h - 0.5
    xs = (np.arange(w) + 0.5) * base.shape[1] / w - 0.5
    y0 = np.clip(np.floor(ys).astype(int), 0, base.shape[0] - 2)
    x0 = np.clip(np.floor(xs).astype(int), 0, base.shape[1] - 2)
    fy = np.clip(ys - y0, 0, 1)[:, None].astype(np.float32)
    fx = np.clip(xs - x0, 0, 1)[None, :].astype(np.float32)
    a, b = base[y0][:, x0], base[y0][:, x0 + 1]
    c, d = base[y0 + 1][:, x0], base[y0 + 1][:, x0 + 1]
    return (a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy


def smooth_noise(i, w=640, h=480):
    rng = np.random.default_rng(1000 + i)
    base = rng.integers(0, 256, (h // 8 + 2, w // 8 + 2)).astype(np.float32)
    img = _upsample(base, w, h) + rng.normal(0, 12, (h, w))
    return np.clip(img, 0, 255).astype(np.uint8)


def rects(i, w=640, h=480):
    rng = np.random.default_rng(5000 + i)
    img = np.full((h, w), 128, np.uint8)
    for _ in range(max(1, w * h // 1500)):
        rw, rh = rng.integers(6, 61, 2)
        x, y = rng.integers(0, max(1, w - 5)), rng.integers(0, max(1, h - 5))
        img[y:y + rh, x:x + rw] = rng.integers(0, 256)
    if _cv2 is not None:
        img = _cv2.GaussianBlur(img, (3, 3), 0.8)
    return img


def frames(n, w=640, h=480, kind="smooth_noise", start=0):
    gen = smooth_noise if kind == "smooth_noise" else rects
    return np.stack([gen(start + i, w, h) for i in range(n)])
