#!/usr/bin/env python3
"""Steady-state latency of the synchronous single-frame entry point (sdorb_extract: host image in, results on the host)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api, synth

for (w, h, params) in ((640, 480, (1000, 1.2, 8, 20)), (752, 480, (1000, 1.2, 8, 20)), (1920, 1080, (4000, 1.2, 12, 20))):
    ex = api.ORBextractor(*params, max_width=w, max_height=h, max_batch=1)
    imgs = [synth.smooth_noise(i, w, h) for i in range(4)]
    for want in (False, True):
        for i in range(5):
            ex(imgs[i % 4], want_pyramid=want)
        ts = []
        for i in range(40):
            t = time.perf_counter()
            k, d, p = ex(imgs[i % 4], want_pyramid=want)
            ts.append(time.perf_counter() - t)
        ts = np.array(ts) * 1e3
        print("%dx%d %s pyramid=%s: median %.3f ms  p10 %.3f  p90 %.3f  (%d kps)" % (w, h, params, want, np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90), len(k)))
    for want in (False, True):  # the C ABI with caller-owned buffers reused across calls (no allocation in the timed region)
        call = ex.single_frame_call(w, h, want_pyramid=want)
        for i in range(5):
            call(imgs[i % 4])
        ts = []
        for i in range(100):
            t = time.perf_counter()
            call(imgs[i % 4])
            ts.append(time.perf_counter() - t)
        ts = np.array(ts) * 1e3
        print("%dx%d preallocated outputs pyramid=%s: median %.3f ms  p10 %.3f  p90 %.3f" % (w, h, want, np.median(ts), np.percentile(ts, 10), np.percentile(ts, 90)))
    ex.set_profiling(True); ex.stage_times(reset=True)
    for i in range(10):
        ex(imgs[i % 4], want_pyramid=False)
    ms, _ = ex.stage_times()
    print("   stage ms per frame:", {k: round(v / 10, 3) for k, v in ms.items()})
    ex.close()
