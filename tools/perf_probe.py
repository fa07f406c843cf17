#!/usr/bin/env python3
"""Per-stage device timing of libsdorb on a resident batch (development aid; bench.py is the judged number)."""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--nfeatures", type=int, default=1000)
    ap.add_argument("--nlevels", type=int, default=8)
    ap.add_argument("--frames", type=int, default=512)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--kind", default="smooth_noise")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    base = synth.frames(a.distinct, a.width, a.height, a.kind)
    reps = (a.frames + a.distinct - 1) // a.distinct
    imgs = torch.from_numpy(np.concatenate([base] * reps)[:a.frames]).to(dev)
    ex = api.ORBextractor(a.nfeatures, 1.2, a.nlevels, 20, max_width=a.width, max_height=a.height, max_batch=a.frames)
    cap = ex.max_keypoints
    kps = torch.zeros((a.frames, cap, 7), dtype=torch.float32, device=dev)
    desc = torch.zeros((a.frames, cap, 32), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(a.frames, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(st)
    sh = st.cuda_stream
    for _ in range(2):
        ex.extract_batch_device(imgs, kps, desc, cnt, stream=sh)
    torch.cuda.synchronize()
    ex.batch_status()
    ex.set_profiling(True)
    ex.stage_times(reset=True)
    t0 = time.time()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        ex.extract_batch_device(imgs, kps, desc, cnt, stream=sh)
    e1.record()
    torch.cuda.synchronize()
    ms, launches = ex.stage_times()
    tot = e0.elapsed_time(e1)
    print("frames=%d %dx%d iters=%d total %.3f ms/iter -> %.0f frames/s (wall %.3f s) counts mean %.1f" % (
        a.frames, a.width, a.height, a.iters, tot / a.iters, a.frames * a.iters / tot * 1e3, time.time() - t0,
        float(cnt.float().mean())))
    for k in api.STAGES:
        print("  %-9s %9.3f ms/iter  launches/iter %d" % (k, ms[k] / a.iters, launches[k] // a.iters))
    # matcher
    nA = torch.full((a.frames // 2,), cap, dtype=torch.int32, device=dev)
    out = torch.zeros((a.frames // 2, cap, 4), dtype=torch.int32, device=dev)
    dA, dB = desc[0::2].contiguous(), desc[1::2].contiguous()
    ex.match_batch(dA, cnt[0::2].contiguous(), dB, cnt[1::2].contiguous(), out=out, device=True, stream=sh)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(a.iters):
        ex.match_batch(dA, cnt[0::2].contiguous(), dB, cnt[1::2].contiguous(), out=out, device=True, stream=sh)
    e1.record()
    torch.cuda.synchronize()
    tm = e0.elapsed_time(e1) / a.iters
    pairs = float((cnt[0::2].double() * cnt[1::2].double()).sum())
    print("match: %.3f ms for %d frame pairs -> %.3f Gpairs/s" % (tm, a.frames // 2, pairs / tm / 1e6))


if __name__ == "__main__":
    main()
