#!/usr/bin/env python3
"""End-to-end (host buffers) rate of sdorb_extract_batch with one and two compute lanes and several pass schedules
(SDORB_PIPE_DUAL / SDORB_PIPE_CONST / SDORB_PIPE_GROWTH are read when the handle is created); every variant's output is
byte-compared with the first one's."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
host_np = bench.make_frames(n, 640, 480)
host = torch.from_numpy(host_np).pin_memory()
ref = None
variants = [(0, 0, 125, 768), (1, 0, 125, 768), (1, 128, 125, 768), (1, 192, 125, 768), (1, 256, 125, 768), (1, 384, 125, 768),
            (0, 256, 125, 768), (1, 256, 125, 256), (1, 128, 125, 128)]
if len(sys.argv) > 2:
    variants = [tuple(int(v) for v in a.split(",")) for a in sys.argv[2:]]
for dual, const, growth, pf in variants:
    os.environ["SDORB_PIPE_DUAL"], os.environ["SDORB_PIPE_CONST"], os.environ["SDORB_PIPE_GROWTH"] = str(dual), str(const), str(growth)
    ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=pf)
    cap = ex.max_keypoints
    hk = torch.zeros((n, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.zeros(n, dtype=torch.int32).pin_memory()
    call = lambda: ex.extract_batch_host(host, hk.numpy().view(api.KP_DTYPE).reshape(n, cap), hd.numpy(), hc.numpy())
    call()
    call()
    ts = []
    for _ in range(5):
        t = time.perf_counter()
        call()
        ts.append(time.perf_counter() - t)
    dt = float(np.median(ts))
    sig = (hk.numpy().tobytes(), hd.numpy().tobytes(), hc.numpy().tobytes())
    if ref is None:
        ref = sig
    print("dual %d const %4d growth %d max_batch %4d: e2e %.0f frames/s (%.2f ms per %d frames, min %.2f) same=%s launches=%d" % (
        dual, const, growth, pf, n / dt, dt * 1e3, n, min(ts) * 1e3, sig == ref, ex.kernel_launches()), flush=True)
    ex.close()
    del hk, hd, hc
