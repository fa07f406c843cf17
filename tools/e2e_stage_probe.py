import os, sys, time
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from sdslam_b200 import api
import bench
dev = torch.device("cuda:0")
n = 4096
host = torch.from_numpy(bench.make_frames(n, 640, 480)).pin_memory()
for pf in (512, 768):
    ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=pf)
    cap = ex.max_keypoints
    hk = torch.zeros((n, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.zeros((n, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.zeros(n, dtype=torch.int32).pin_memory()
    call = lambda: ex.extract_batch_host(host, hk.numpy().view(api.KP_DTYPE).reshape(n, cap), hd.numpy(), hc.numpy())
    call(); call()
    t = time.perf_counter()
    for _ in range(3): call()
    dt = (time.perf_counter() - t) / 3
    ex.set_profiling(True); ex.stage_times(reset=True)
    t = time.perf_counter(); call(); dtp = time.perf_counter() - t
    ms, L = ex.stage_times()
    print("pass %d: e2e %.0f frames/s, %.2f ms per call; profiled call %.2f ms, kernel stage sum %.2f ms %s launches %d" % (
        pf, n / dt, dt * 1e3, dtp * 1e3, sum(ms.values()), {k: round(v, 2) for k, v in ms.items()}, sum(L.values())))
    ex.close()
