#!/usr/bin/env python3
"""Device-resident throughput of the local-mapping / loop-closing matchers (development aid): the keypoint search of Fuse,
SearchBySim3, the Sim3 SearchByProjection and the local-map search, on 1000-keypoint keyframes with 1000 map points each."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import search_cases as sc  # noqa: E402
from sdslam_b200 import api  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
DISTINCT, N = 8, 1000
dev = torch.device("cuda:0")
ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=8)
gp = sc.grid_params()
sf = (np.float32(1.2) ** np.arange(8)).astype(np.float32)
inv = (np.float32(1) / (sf * sf)).astype(np.float32)


def slab(arrs, dtype, tail=()):
    out = np.zeros((len(arrs), N) + tuple(tail), dtype)
    for p, a in enumerate(arrs):
        out[p, :len(a)] = a
    return out


def t(a):
    a = np.ascontiguousarray(a.view(np.float32).reshape(a.shape + (7,)) if a.dtype == api.KP_DTYPE else a)
    reps = (F + len(a) - 1) // len(a)
    return torch.from_numpy(np.concatenate([a] * reps)[:F]).to(dev)


def timed(fn, iters=10):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


torch.cuda.set_stream(torch.cuda.Stream(dev))  # a real stream: stream 0 would make the library fall back to its own
stream = torch.cuda.current_stream().cuda_stream
assert stream != 0
frames = []
for s in range(DISTINCT):
    _, _, kf, df = sc.frame_pair(s + 140, 10, N, level0=0.3)
    proj, lvl, fl, ur = sc.fuse_inputs(s, kf, N, stereo=True)
    frames.append((proj, lvl, fl, sc.fuse_descriptors(s, df, N, N), kf, df, ur, sc.map_point_inputs(s, kf, df, N)[1]))
kf = slab([f[4] for f in frames], api.KP_DTYPE)
n = np.full(DISTINCT, N, np.int32)
cs, idx = ex.assign_grid_batch(kf, n, *gp)
grid = (t(cs), t(idx)) + tuple(gp)
a = dict(proj=t(slab([f[0] for f in frames], np.float32, (3,))), level=t(slab([f[1] for f in frames], np.int32)),
         flags=t(slab([f[2] for f in frames], np.uint8)), dmp=t(slab([f[3] for f in frames], np.uint8, (32,))), n=t(n), kf=t(kf),
         df=t(slab([f[5] for f in frames], np.uint8, (32,))), ur=t(slab([f[6] for f in frames], np.float32)),
         vc=t(slab([f[7] for f in frames], np.float32)), occ=torch.zeros((F, N), dtype=torch.uint8, device=dev))
bi = torch.zeros((F, N), dtype=torch.int32, device=dev)
bd = torch.zeros((F, N), dtype=torch.int32, device=dev)
ms = timed(lambda: ex.fuse_search_batch(a["proj"], a["level"], a["flags"], a["dmp"], a["n"], a["kf"], a["df"], a["ur"], grid, sf, inv, 3.0,
                                        best_idx=bi, best_dist=bd, device=True, stream=stream))
print("Fuse search       : %d keyframes x %d map points: %.3f ms -> %.1f M map points/s, %.0f k keyframes/s (fused %d per keyframe)" % (
    F, N, ms, F * N / ms / 1e3, F / ms, int((bi >= 0).sum().item()) // F), flush=True)
cases = [sc.sim3_case(s, N, N) for s in range(DISTINCT)]


def sim3_side(j):
    ks = slab([c[j][4] for c in cases], api.KP_DTYPE)
    cs, idx = ex.assign_grid_batch(ks, n, *gp)
    return (t(slab([c[j][0] for c in cases], np.float32, (3,))), t(slab([c[j][1] for c in cases], np.int32)),
            t(slab([c[j][2] for c in cases], np.uint8)), t(slab([c[j][3] for c in cases], np.uint8, (32,))), t(n), t(ks),
            t(slab([c[j][5] for c in cases], np.uint8, (32,))), (t(cs), t(idx)) + tuple(gp))


side1, side2 = sim3_side(0), sim3_side(1)
out = (torch.zeros(F, dtype=torch.int32, device=dev),) + tuple(torch.zeros((F, N), dtype=torch.int32, device=dev) for _ in range(3))
ms = timed(lambda: ex.search_by_sim3_batch(side1, side2, sf, 7.5, out=out, device=True, stream=stream))
print("SearchBySim3      : %d keyframe pairs: %.3f ms -> %.0f k pairs/s (found %.0f per pair)" % (F, ms, F / ms, out[0].float().mean().item()),
      flush=True)
ex.close()
