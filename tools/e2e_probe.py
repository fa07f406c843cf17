#!/usr/bin/env python3
"""H2D / D2H bandwidth of the box and the end-to-end (host buffers) rate of sdorb_extract_batch for several pass sizes."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api
import bench

dev = torch.device("cuda:0")
n = 2048
host_np = bench.make_frames(n, 640, 480)
host = torch.from_numpy(host_np).pin_memory()
d = torch.empty_like(host, device=dev)
for _ in range(2):
    d.copy_(host, non_blocking=True)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3):
    d.copy_(host, non_blocking=True)
torch.cuda.synchronize()
dt = (time.perf_counter() - t) / 3
print("H2D pinned: %.1f GB/s (%.1f ms for %d frames)" % (host.numel() / dt / 1e9, dt * 1e3, n))
back = torch.empty_like(host).pin_memory()
t = time.perf_counter()
back.copy_(d, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t
print("D2H pinned: %.1f GB/s" % (host.numel() / dt / 1e9))
for pf in (64, 128, 256, 512):
    ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=pf)
    cap = ex.max_keypoints
    hk = torch.zeros((n, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.zeros(n, dtype=torch.int32).pin_memory()
    call = lambda: ex.extract_batch_host(host, hk.numpy().view(api.KP_DTYPE).reshape(n, cap), hd.numpy(), hc.numpy())
    call()
    t = time.perf_counter()
    for _ in range(3):
        call()
    dt = (time.perf_counter() - t) / 3
    print("pass %4d frames: e2e %.0f frames/s (%.1f ms per %d frames)" % (pf, n / dt, dt * 1e3, n))
    ex.close()
