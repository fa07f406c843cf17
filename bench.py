#!/usr/bin/env python3
"""bench.py -- the headline metric of BASELINE.json on this repo's CUDA path and on the CPU reference arm.

  python bench.py [--gpus N] [--steps K] [--warmup W]            (N > 1: launched by torch.distributed.run)
  python bench.py --impl reference ...                            (the reference's own code on the box's host cores)

Metric: ORB extract+describe frames/s on 640x480 frames, 1000 keypoints, 8 levels, scale 1.2, FAST 20
(SD-SLAM ORBextractor::operator(), /root/reference/src/ORBextractor.cc:620-678), plus Hamming pairs/s of the batched
ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:1459-1473) as a side figure.

Workload C3 (BASELINE.json configs[2]): ONE batch of 4096 synthetic frames, frame-sharded over the N GPUs (contiguous ranges,
sdslam_b200/sharding.py) -- strong scaling: the batch is fixed, every rank owns 4096 / N frames.  Frames are independent, there
is no collective in the loop; one NCCL gather of the result slabs to rank 0 at the end (timed separately).  A step = one pass of
the hot path over the batch.
  value : frames/s with the frames resident in HBM, CUDA events on the launching stream, max over ranks, NO profiling events
          inside the timed region (the per-stage table comes from a second loop)
  e2e   : the same batch through the C ABI with HOST buffers (sdorb_extract_batch, SDORB_MEM_HOST): pinned host frames -> H2D
          -> kernels -> D2H of keypoints / descriptors / counts inside the timed region (the north-star 4-argument surface);
          e2e_with_pyramid adds the reference's fifth output, imagePyramid (sdorb_extract_batch_pyr)
  roofline : the dominant kernel's algorithmic bytes / its CUDA-event time against MEASURED_PEAKS.json, plus -- because that
          kernel is bound by the integer ALU pipe, not by HBM -- its ALU-pipe instruction rate against the pipe's measured peak
  configs : the other single-GPU configurations of BASELINE.json (C2 752x480, C5 1920x1080 / 4000 kp / 12 levels), same legs
  cpu_baseline : the reference's own ORBextractor.cc compiled unmodified (oracle/_ref) on the host cores, bounded sample
Only the cpu_baseline / --impl reference legs touch oracle/; the reference arm never loads libsdorb.so.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "ORB extract+describe frames/s (640x480, 1k kp)"
WORKLOADS = {
    # name: (width, height, nfeatures, scale, nlevels, thFAST, frames per step over all GPUs)
    "C3_tum_640x480_1000kp_8lv": (640, 480, 1000, 1.2, 8, 20, 4096),
    "C2_euroc_752x480_1000kp_8lv": (752, 480, 1000, 1.2, 8, 20, 2048),
    "C5_1080p_4000kp_12lv": (1920, 1080, 4000, 1.2, 12, 20, 256),
}
DISTINCT = 64  # distinct generator frames; the rest of the batch are column-rotated copies (distinct bytes, same statistics)
PROBE_FRAMES = 64  # frames every rank extracts for the N-GPU == 1-GPU byte-identity check


def make_frames(nframes, w, h, first=0):
    """Frames [first, first + nframes) of the deterministic global batch: frame i = smooth_noise(i % DISTINCT) rotated by
    8 * (i // DISTINCT) columns."""
    from sdslam_b200 import synth
    need = sorted({(first + j) % DISTINCT for j in range(nframes)})
    base = {i: synth.smooth_noise(i, w, h) for i in need}
    out = np.empty((nframes, h, w), np.uint8)
    for j in range(nframes):
        i = first + j
        out[j] = np.roll(base[i % DISTINCT], 8 * (i // DISTINCT), axis=1)
    return out


def config_of(workload):
    """The SAME dict in both arms (the driver compares them)."""
    w, h, nf, sf, nl, th, frames = WORKLOADS[workload]
    return {"workload": workload, "width": w, "height": h, "nfeatures": nf, "scale_factor": sf, "nlevels": nl, "th_fast": th,
            "frames_per_step": frames}


def stage_bytes_per_frame(geom, nkp):
    """Algorithmic bytes per frame of every stage (SURVEY section 8d / DESIGN.md): each stage reads its input once
    and writes its output once."""
    px = [int(g["width"]) * int(g["height"]) for g in geom]
    P = sum(px)
    return {
        "pyramid": (P - px[-1]) + (P - px[0]),
        "fast": P,
        "blur": 2 * P,
        # latency-bound stages, listed for completeness (no roofline claim): packed entries in / out
        "select": 8 * nkp,
        "describe": nkp * (749 + 512 + 60),
        # what a single fused pass could not avoid: the frame in, the pyramid levels 1.. out, keypoints + descriptors out
        "fused_lower_bound": px[0] + (P - px[0]) + 60 * nkp,
    }


class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        rows = [r for r in self.rows if t0 - 0.05 <= r[0] <= t1 + 0.15] or self.rows
        for _, line in rows:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def measured_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


# ------------------------------------------------------------------------------------------- CPU legs
def cpu_extractor(params):
    """(extractor with .extract_many, kind, description): the reference's own ORBextractor.cc compiled unmodified (oracle/_ref;
    the -march=native build where this host has every CPU flag of the build machine), else the oracle port."""
    from oracle import ref_binding as ref
    if ref.available():
        native = ref._native_runs_here()
        return (ref.Extractor(*params, native=native), "reference",
                "/root/reference/src/ORBextractor.cc compiled unmodified (oracle/_ref, -O3 %s, cv:: = oracle/ref_compat)"
                % ("-march=native" if native else "-march=x86-64-v3"))
    from oracle import binding as orc
    orc.use_native()
    return orc.Extractor(*params), "port", "oracle port (oracle/sdorb_oracle.cc); oracle/_ref is not present on this box"


def cpu_extract_rate(ex, frames, nthreads):
    t = time.perf_counter()
    _, _, counts = ex.extract_many(frames, nthreads=nthreads, want_outputs=False)
    dt = time.perf_counter() - t
    return len(frames) / dt, int(counts.sum())


def cpu_match_rate(nthreads, npairs, rng):
    from oracle import binding as orc
    A = rng.integers(0, 256, (npairs, 1000, 32), dtype=np.uint8)
    B = rng.integers(0, 256, (npairs, 1000, 32), dtype=np.uint8)
    n = np.full(npairs, 1000, np.int32)
    t = time.perf_counter()
    orc.match_many(A, n, B, n, nthreads=nthreads)
    return npairs * 1e6 / (time.perf_counter() - t)


def cv2_primitives_rate(params, frames, nthreads):
    """Frames/s of OpenCV's own SIMD kernels for the three heavy primitives of the reference path (chained cv2.resize +
    copyMakeBorder, cv2.FAST on every cell ROI, cv2.GaussianBlur per level) -- no culling, orientation or descriptors, so
    an UPPER bound on what the reference linked against a real OpenCV could reach on this host.  cv2 releases the GIL, so
    frames are spread over a thread pool.  Context only.  (Geometry from the oracle: this leg must not load libsdorb.so.)"""
    try:
        import cv2
    except Exception:
        return None
    from concurrent.futures import ThreadPoolExecutor
    from oracle import binding as orc
    nf, sf, nl, th = params
    h, w = frames.shape[1:]
    geom = orc.Extractor(nf, sf, nl, th).geometry(w, h)
    cv2.setNumThreads(1)
    fast = cv2.FastFeatureDetector_create(threshold=th, nonmaxSuppression=True, type=cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)

    def one(img):
        lvl, n = img, 0
        for l, g in enumerate(geom):
            if l:
                lvl = cv2.resize(lvl, (int(g["width"]), int(g["height"])), interpolation=cv2.INTER_LINEAR)
            cv2.copyMakeBorder(lvl, 19, 19, 19, 19, cv2.BORDER_REFLECT_101)
            cols, rows, cw, ch = int(g["level_cols"]), int(g["level_rows"]), int(g["cell_w"]), int(g["cell_h"])
            lh, lw = lvl.shape
            for i in range(max(rows, 0)):
                y0 = 16 + i * ch
                y1 = lh - 16 if i == rows - 1 else y0 + ch + 6
                for j in range(max(cols, 0)):
                    x0 = 16 + j * cw
                    x1 = lw - 16 if j == cols - 1 else x0 + cw + 6
                    if y1 - y0 >= 7 and x1 - x0 >= 7:
                        n += len(fast.detect(lvl[y0:y1, x0:x1]))
            cv2.GaussianBlur(lvl, (7, 7), 2, None, 2, cv2.BORDER_REFLECT_101)
        return n

    one(frames[0])
    t = time.perf_counter()
    with ThreadPoolExecutor(nthreads) as pool:
        list(pool.map(one, frames))
    return len(frames) / (time.perf_counter() - t)


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path (oracle/_ref: ORBextractor.cc compiled unmodified)
    on all host threads, frame-parallel, on the same config; each step a bounded sample of the workload.  Rank 0 only."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    w, h, nf, sf, nl, th, _ = WORKLOADS[args.workload]
    cores = os.cpu_count() or 1
    per_step = max(16 * cores, 128)  # >= 16 frames per host thread per step: thread start-up is amortised
    if w * h > 1000 * 1000:
        per_step = max(4 * cores, 32)
    frames = make_frames(per_step, w, h)
    ex, kind, what = cpu_extractor((nf, sf, nl, th))
    for _ in range(max(args.warmup, 1)):
        cpu_extract_rate(ex, frames[:2 * cores], cores)
    t = time.perf_counter()
    total = 0
    for _ in range(args.steps):
        _, n = cpu_extract_rate(ex, frames, cores)
        total += n
    dt = time.perf_counter() - t
    fps = args.steps * per_step / dt
    sample = "%d frames per step (%d per host thread) of the workload's batch, %d threads frame-parallel; %s" % (
        per_step, per_step // cores, cores, what)
    cv2_n = cv2_primitives_rate((nf, sf, nl, th), frames[:max(2 * cores, 32)], cores)
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": config_of(args.workload),
            "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                             "cv2_primitives_only": {"value": cv2_n, "unit": "frames/s", "cores": cores,
                                                     "what": "OpenCV 4.13 SIMD resize + per-cell FAST + GaussianBlur only: upper bound "
                                                             "for the reference linked against a real OpenCV on this host"}},
            "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "keypoints_per_frame": total / max(1, args.steps * per_step)}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------- GPU arm
class Dist:
    """torch.distributed plumbing: rank / world, barrier, max / sum over ranks (device-side, NCCL)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = None
        if self.world > 1:
            # keep this rank's threads -- and so its pinned host batch (first touch) -- on the CPUs next to its GPU
            try:
                import pynvml
                pynvml.nvmlInit()
                pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(self.local))
                cpus = sorted(os.sched_getaffinity(0))
                self.numa = "%d cpus [%d..%d]" % (len(cpus), cpus[0], cpus[-1])
            except Exception as e:  # not fatal: affinity is an optimisation
                self.numa = "unset (%s)" % type(e).__name__
            # NCCL writes its version banner to stdout when the first communicator comes up; rank 0 must print exactly one JSON
            # line, so file descriptor 1 points at stderr until the communicator exists
            sys.stdout.flush()
            saved = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize(self.dev)
            finally:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def reduce(self, x, op="max"):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX if op == "max" else self.dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self.world > 1:
            self.dist.barrier()
            self.dist.destroy_process_group()


def copy_ceiling(D, host, d2h_bytes, reps=4):
    """What the platform gives bare pinned copies with all ranks running at once (no kernels): every rank uploads its pinned
    batch and downloads a result-sized buffer CONCURRENTLY on two streams, `reps` times between barriers; aggregate GB/s over the
    ranks.  The end-to-end leg cannot beat the H2D figure: it moves the same bytes through the same root complex."""
    torch = D.torch
    dbuf = torch.empty_like(host, device=D.dev)
    dres = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=D.dev)
    hres = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    s_in, s_out = torch.cuda.Stream(D.dev), torch.cuda.Stream(D.dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]

    def go(n):
        with torch.cuda.stream(s_in):
            ev[0].record(s_in)
            for _ in range(n):
                dbuf.copy_(host, non_blocking=True)
            ev[1].record(s_in)
        with torch.cuda.stream(s_out):
            ev[2].record(s_out)
            for _ in range(n):
                hres.copy_(dres, non_blocking=True)
            ev[3].record(s_out)
    go(1)
    D.barrier()
    t0 = time.perf_counter()
    go(reps)
    D.barrier()
    wall = D.reduce(time.perf_counter() - t0)
    h2d = D.reduce(host.numel() * reps / (ev[0].elapsed_time(ev[1]) * 1e-3) / 1e9, "sum")
    d2h = D.reduce(d2h_bytes * reps / (ev[2].elapsed_time(ev[3]) * 1e-3) / 1e9, "sum")
    both = D.world * (host.numel() + d2h_bytes) * reps / wall / 1e9
    del dbuf, dres, hres
    return {"h2d_gbs": h2d, "d2h_gbs": d2h, "both_directions_wall_gbs": both,
            "how": "every rank: %d x (pinned H2D of its %.0f MB batch || D2H of %.0f MB) on two streams, all ranks at once; sum of per-rank "
                   "CUDA-event rates" % (reps, host.numel() / 1e6, d2h_bytes / 1e6)}


def measure(D, params, w, h, frames_per_gpu, first_frame, pass_frames, e2e_pass_frames, steps, warmup, e2e_steps, want_pyr_e2e=True,
            want_ceiling=False, keep=False):
    """All legs of one workload on this rank's share of the batch.  Returns a dict of per-rank figures already reduced over the
    ranks where a reduction is meaningful (max time)."""
    torch = D.torch
    from sdslam_b200 import api
    dev, world = D.dev, D.world
    host_np = make_frames(frames_per_gpu, w, h, first=first_frame)
    host = torch.from_numpy(host_np).pin_memory()
    dimgs = host.to(dev, non_blocking=True)
    pass_frames = max(1, min(pass_frames, frames_per_gpu))
    ex = api.ORBextractor(*params, device=D.local, max_width=w, max_height=h, max_batch=pass_frames)
    cap = ex.max_keypoints
    kps = torch.zeros((frames_per_gpu, cap, 7), dtype=torch.float32, device=dev)
    desc = torch.zeros((frames_per_gpu, cap, 32), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(frames_per_gpu, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(dev)  # non-default: the events below and the kernels share it
    geom = api.host_level_geometry(*params, w, h)
    sbytes = stage_bytes_per_frame(geom, cap)

    def step_resident():
        ex.extract_batch_device(dimgs, kps, desc, cnt, stream=stream.cuda_stream)

    # nvidia-smi needs up to a second before its first sample: it starts ahead of the warm-up, and the warm-up lasts until it reports
    sampler = ClockSampler(D.local) if D.rank == 0 else None
    with torch.cuda.stream(stream):
        for _ in range(max(warmup, 3)):
            step_resident()
        t_wait = time.time()
        while sampler and sampler.proc and not sampler.rows and time.time() - t_wait < 3.0:
            step_resident()
            torch.cuda.synchronize()
    D.barrier()
    ex.batch_status()

    # ---- timed region 1 (the headline): frames resident in HBM, no profiling events between the kernels
    launches0 = ex.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    D.barrier()
    t_mark0 = time.time()
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            step_resident()
        e1.record(stream)
    D.barrier()
    t_mark1 = time.time()
    ms_total = D.reduce(e0.elapsed_time(e1))
    launches = ex.kernel_launches() - launches0
    ex.batch_status()
    if sampler and sampler.proc and sum(1 for r in sampler.rows if t_mark0 - 0.05 <= r[0] <= t_mark1 + 0.15) < 2:
        # a timed region shorter than two sampling periods: keep the same load up (untimed) until two samples lie inside the window
        t_wait = time.time()
        with torch.cuda.stream(stream):
            while sum(1 for r in sampler.rows if t_mark0 - 0.05 <= r[0]) < 2 and time.time() - t_wait < 2.0:
                step_resident()
                torch.cuda.synchronize()
        t_mark1 = time.time()
    clocks = sampler.stop(t_mark0, t_mark1) if sampler else None
    mean_kp = float(cnt.float().mean().item())

    # ---- per-stage table: a second loop with the library's stage events switched on (they serialise the stages: no programmatic
    # dependent launch across an event), same steps
    ex.set_profiling(True)
    ex.stage_times(reset=True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        p0.record(stream)
        for _ in range(steps):
            step_resident()
        p1.record(stream)
    D.barrier()
    ms_profiled = D.reduce(p0.elapsed_time(p1))
    stage_ms, stage_launches = ex.stage_times(reset=True)
    ex.set_profiling(False)

    # ---- timed region 2: end to end through the C ABI with host buffers
    hk = torch.zeros((frames_per_gpu, cap, 7), dtype=torch.float32).pin_memory()
    hd = torch.zeros((frames_per_gpu, cap, 32), dtype=torch.uint8).pin_memory()
    hc = torch.zeros(frames_per_gpu, dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(steps, e2e_steps))
    e2e_pass = max(1, min(e2e_pass_frames, frames_per_gpu))
    # its own handle: the host pipeline overlaps H2D / kernels / D2H pass by pass, so it wants smaller passes than the resident leg
    ex_e2e = ex if e2e_pass == pass_frames else api.ORBextractor(*params, device=D.local, max_width=w, max_height=h, max_batch=e2e_pass)
    hk_np = hk.numpy().view(api.KP_DTYPE).reshape(frames_per_gpu, cap)

    def timed_host_leg(fn):
        fn()
        D.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            fn()  # returns when the results are in host memory
        torch.cuda.synchronize(dev)
        return D.reduce(time.perf_counter() - t0)

    e2e_s = timed_host_leg(lambda: ex_e2e.extract_batch_host(host, hk_np, hd.numpy(), hc.numpy()))
    same = bool((hc.numpy() == cnt.cpu().numpy()).all()) and hd.numpy()[:8].tobytes() == desc[:8].cpu().numpy().tobytes()
    d2h_bytes = int(hk.numel() * 4 + hd.numel() + hc.numel() * 4)
    out = {"frames_per_gpu": frames_per_gpu, "pass_frames": pass_frames, "e2e_pass_frames": e2e_pass, "cap": cap, "sbytes": sbytes,
           "ms_total": ms_total, "ms_profiled": ms_profiled, "launches": int(launches), "clocks": clocks, "mean_kp": mean_kp,
           "stage_ms": stage_ms, "stage_launches": stage_launches, "e2e_s": e2e_s, "e2e_steps": e2e_steps, "same": same,
           "h2d_bytes": int(D.reduce(host.numel(), "sum")), "d2h_bytes": int(D.reduce(d2h_bytes, "sum"))}  # whole job
    if want_pyr_e2e:
        # the reference's real 5-argument call also returns imagePyramid: levels >= 1 of every frame come back as well
        # (level 0 is the caller's own input, first_level = 1)
        _, fb = ex_e2e.pyramid_layout(w, h)
        hp = torch.zeros(frames_per_gpu * fb, dtype=torch.uint8).pin_memory()
        out["e2e_pyr_s"] = timed_host_leg(lambda: ex_e2e.extract_batch_host(host, hk_np, hd.numpy(), hc.numpy(), pyramid=hp.numpy(), first_level=1))
        lvl_bytes = sum(int(g["width"]) * int(g["height"]) for g in geom[1:])
        out["d2h_pyr_bytes"] = int(D.reduce(d2h_bytes + frames_per_gpu * lvl_bytes, "sum"))
        del hp
    if want_ceiling:
        out["ceiling"] = copy_ceiling(D, host, d2h_bytes)
    if ex_e2e is not ex:
        ex_e2e.close()
    if keep:
        out.update(ex=ex, host_np=host_np, host=host, dimgs=dimgs, kps=kps, desc=desc, cnt=cnt, stream=stream, geom=geom)
    else:
        ex.close()
    return out


def stage_table(m, steps, peak):
    nframes_total = m["frames_per_gpu"] * steps
    stages = {}
    for s in ("pyramid", "fast", "select", "blur", "describe"):
        ms = m["stage_ms"][s]
        gbs = m["sbytes"][s] * nframes_total / (ms * 1e-3) / 1e9 if ms > 0 else None
        stages[s] = {"ms_per_step": ms / steps, "launches_per_step": m["stage_launches"][s] / steps, "bytes_per_frame": m["sbytes"][s],
                     "gbs": gbs, "frac": gbs / peak if gbs else None}
    return stages


def side_config(D, name, args, peak):
    """One of the other single-GPU configurations of BASELINE.json (C2 / C5): resident frames/s, per-stage fractions, e2e."""
    w, h, nf, sf, nl, th, frames = WORKLOADS[name]
    steps = 3
    m = measure(D, (nf, sf, nl, th), w, h, frames, 0, min(args.pass_frames, frames), max(frames // 4, 1), steps, 3, 2, want_pyr_e2e=False)
    return {"config": config_of(name), "value": frames * steps / (m["ms_total"] * 1e-3), "unit": "frames/s", "steps": steps,
            "ms_per_step": m["ms_total"] / steps, "frames_per_pass": m["pass_frames"],
            "e2e": {"value": frames * m["e2e_steps"] / m["e2e_s"], "unit": "frames/s", "h2d_bytes_per_step": m["h2d_bytes"],
                    "d2h_bytes_per_step": m["d2h_bytes"], "matches_resident_run": m["same"], "frames_per_pass": m["e2e_pass_frames"]},
            "stages": stage_table(m, steps, peak), "keypoints_per_frame": m["mean_kp"], "gpu_launches": m["launches"]}


def run_ours(args):
    D = Dist()
    torch, dist = D.torch, D.dist
    from sdslam_b200 import api, sharding
    world, rank, dev = D.world, D.rank, D.dev
    w, h, nf, sf, nl, th, frames_global = WORKLOADS[args.workload]
    if args.frames:
        frames_global = args.frames
    params = (nf, sf, nl, th)
    peak, peak_src = measured_peak()

    # ---- the headline: ONE batch of frames_global frames, frame-sharded over the ranks (strong scaling)
    lo, hi = sharding.frame_range(rank, world, frames_global)
    m = measure(D, params, w, h, hi - lo, lo, args.pass_frames, args.e2e_pass_frames, args.steps, args.warmup, args.e2e_steps,
                want_ceiling=True, keep=True)
    ex, host_np, dimgs, kps, desc, cnt, stream, cap = m["ex"], m["host_np"], m["dimgs"], m["kps"], m["desc"], m["cnt"], m["stream"], m["cap"]
    frames_per_gpu = hi - lo

    # ---- the one collective of the job: the result slabs of the whole batch gathered on rank 0 over NCCL (send / recv into place,
    # no padding, nobody but rank 0 receives); warmed once, then timed
    gather = None
    if world > 1:
        sharding.gather_slabs([kps, desc, cnt], frames_global)
        D.barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        got = sharding.gather_slabs([kps, desc, cnt], frames_global)
        g1.record()
        torch.cuda.synchronize(dev)
        gms = D.reduce(g0.elapsed_time(g1))
        gbytes = (frames_global - sharding.frame_range(0, world, frames_global)[1]) * (cap * 60 + 4)
        gather = {"ms": gms, "bytes_received_by_rank0": int(gbytes), "gbs": gbytes / (gms * 1e-3) / 1e9,
                  "how": "torch.distributed batch_isend_irecv (NCCL over NVLink), every rank's slab straight into rank 0's result"}
        if rank == 0:
            assert got[0].shape[0] == frames_global and torch.equal(got[2][:frames_per_gpu], cnt)
        del got

    # ---- N-GPU == 1-GPU byte identity on hardware: every rank extracts the SAME probe frames, rank 0
    # compares every rank's slabs with its own
    identical = None
    if world > 1:
        pn = min(PROBE_FRAMES, frames_global)
        probe = torch.from_numpy(make_frames(pn, w, h, first=0)).to(dev)
        pk = torch.zeros((pn, cap, 7), dtype=torch.float32, device=dev)
        pd = torch.zeros((pn, cap, 32), dtype=torch.uint8, device=dev)
        pc = torch.zeros(pn, dtype=torch.int32, device=dev)
        with torch.cuda.stream(stream):
            ex.extract_batch_device(probe, pk, pd, pc, stream=stream.cuda_stream)
        torch.cuda.synchronize(dev)
        blob = torch.cat([pk.view(torch.uint8).reshape(-1), pd.reshape(-1), pc.view(torch.uint8).reshape(-1)])
        allb = [torch.empty_like(blob) for _ in range(world)] if rank == 0 else None
        dist.gather(blob, allb, dst=0)
        if rank == 0:
            identical = all(torch.equal(b, blob) for b in allb)
        del probe, pk, pd, pc, blob, allb

    # ---- Hamming side figure: frame pairs (2k, 2k+1) of this rank's share, best / second-best, ratio 0.75, TH_LOW 50
    npairs = frames_per_gpu // 2
    dA, dB = desc[0:2 * npairs:2].contiguous(), desc[1:2 * npairs:2].contiguous()
    nA, nB = cnt[0:2 * npairs:2].contiguous(), cnt[1:2 * npairs:2].contiguous()
    mout = torch.zeros((npairs, cap, 4), dtype=torch.int32, device=dev)
    m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ex.match_batch(dA, nA, dB, nB, out=mout, device=True, stream=stream.cuda_stream)
        m0.record(stream)
        for _ in range(5):
            ex.match_batch(dA, nA, dB, nB, out=mout, device=True, stream=stream.cuda_stream)
        m1.record(stream)
    D.barrier()
    match_ms = D.reduce(m0.elapsed_time(m1) / 5)
    pairs_local = float((nA.double() * nB.double()).sum().item())
    pairs_per_s = D.reduce(pairs_local, "sum") / (match_ms * 1e-3)

    # ---- measured pipe peaks on this device (kernels_probe.cu): POPC for the matcher, the ALU pipe (VIMNMX3.U16x2, PRMT) for FAST
    popc_s, popc_clk = ex.pipe_probe(0)
    alu_s, alu_clk = ex.pipe_probe(1)
    prmt_s, prmt_clk = ex.pipe_probe(2)

    # ---- MapPoint::ComputeDistinctiveDescriptors side figure: groups of 16 consecutive descriptors of this batch as map points
    nsets = min(frames_per_gpu, 512) * (cap // 16)
    d_off = torch.arange(0, 16 * nsets + 1, 16, dtype=torch.int32, device=dev)
    d_idx = torch.zeros(nsets, dtype=torch.int32, device=dev)
    d_med = torch.zeros(nsets, dtype=torch.int32, device=dev)
    flat = desc.reshape(-1, 32)
    q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        ex.distinctive_batch(flat, d_off, d_idx, d_med, device=True, stream=stream.cuda_stream)
        q0.record(stream)
        ex.distinctive_batch(flat, d_off, d_idx, d_med, device=True, stream=stream.cuda_stream)
        q1.record(stream)
    D.barrier()
    distinctive_sets_per_s = world * nsets / (D.reduce(q0.elapsed_time(q1)) * 1e-3)

    # ---- guided-matcher side figure: ORBmatcher::SearchForInitialization (windowSize 100, nnratio 0.9, orientation check) on the
    # frame pairs (2k, 2k+1), device-resident: Frame::AssignFeaturesToGrid of frame 2k+1, then the search
    kA, kB = kps[0:2 * npairs:2].contiguous(), kps[1:2 * npairs:2].contiguous()
    inv_w, inv_h = float(np.float32(64) / np.float32(w)), float(np.float32(48) / np.float32(h))
    g_cs = torch.zeros((npairs, 64 * 48 + 1), dtype=torch.int32, device=dev)
    g_ix = torch.zeros((npairs, cap), dtype=torch.int32, device=dev)
    s_prev0 = kA[:, :, 0:2].contiguous()
    s_prev = s_prev0.clone()
    s_m12 = torch.zeros((npairs, cap), dtype=torch.int32, device=dev)
    s_nm = torch.zeros(npairs, dtype=torch.int32, device=dev)
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)

    def search_once():
        ex.assign_grid_batch(kB, nB, 0.0, 0.0, inv_w, inv_h, cell_start=g_cs, indices=g_ix, device=True, stream=stream.cuda_stream)
        ex.search_for_initialization_batch(kA, dA, nA, kB, dB, nB, (g_cs, g_ix, 0.0, 0.0, inv_w, inv_h), s_prev, 100, 0.9, True,
                                           matches12=s_m12, nmatches=s_nm, device=True, stream=stream.cuda_stream)

    with torch.cuda.stream(stream):
        search_once()
        s_prev.copy_(s_prev0)
        s0.record(stream)
        search_once()
        s1.record(stream)
    D.barrier()
    search_pairs_per_s = world * npairs / (D.reduce(s0.elapsed_time(s1)) * 1e-3)
    search_matches = float(s_nm.double().mean().item())

    # ---- ORB-SLAM2-style mode side figure (row f1: iniThFAST 20 / minThFAST 7 on 30-pixel cells + DistributeOctTree), the first
    # frames of the same share, resident
    o2_n = min(frames_per_gpu, 2 * args.pass_frames)
    ex2 = api.ORBextractor(nf, sf, nl, th, minThFAST=7, device=D.local, max_width=w, max_height=h, max_batch=min(args.pass_frames, o2_n))
    cap2 = ex2.max_keypoints
    kps2 = torch.zeros((o2_n, cap2, 7), dtype=torch.float32, device=dev)
    desc2 = torch.zeros((o2_n, cap2, 32), dtype=torch.uint8, device=dev)
    cnt2 = torch.zeros(o2_n, dtype=torch.int32, device=dev)
    o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        for _ in range(2):
            ex2.extract_batch_device(dimgs[:o2_n], kps2, desc2, cnt2, stream=stream.cuda_stream)
        o0.record(stream)
        for _ in range(3):
            ex2.extract_batch_device(dimgs[:o2_n], kps2, desc2, cnt2, stream=stream.cuda_stream)
        o1.record(stream)
    D.barrier()
    orbslam2 = {"frames_per_s": world * o2_n * 3 / (D.reduce(o0.elapsed_time(o1)) * 1e-3), "frames_per_gpu_per_step": o2_n, "ini_th_fast": th,
                "min_th_fast": 7, "keypoints_per_frame": float(cnt2.double().mean().item())}
    ex2.close()
    del kps2, desc2, cnt2

    # ---- free the big buffers of the headline before the side configurations
    ex.close()
    host_probe = host_np[:64].copy()
    del dimgs, kps, desc, cnt, mout, dA, dB, kA, kB, flat, m["host"], m["dimgs"], m["kps"], m["desc"], m["cnt"]
    torch.cuda.empty_cache()

    # ---- weak-scaling side figure at N > 1 (round 1's headline): every rank owns a full 4096-frame batch
    weak = None
    if world > 1 and not args.no_side:
        mw = measure(D, params, w, h, frames_global, rank * DISTINCT, args.pass_frames, args.e2e_pass_frames, min(args.steps, 5), 3,
                     min(args.e2e_steps, 3), want_pyr_e2e=False)
        ws = min(args.steps, 5)
        weak = {"value": world * frames_global * ws / (mw["ms_total"] * 1e-3), "unit": "frames/s", "frames_per_gpu_per_step": frames_global,
                "steps": ws, "e2e": {"value": world * frames_global * mw["e2e_steps"] / mw["e2e_s"], "unit": "frames/s"}}
        torch.cuda.empty_cache()

    # ---- the other single-GPU configurations of BASELINE.json (rank 0's GPU only; N = 1 line only)
    side = None
    if world == 1 and not args.no_side:
        side = {}
        for name in ("C2_euroc_752x480_1000kp_8lv", "C5_1080p_4000kp_12lv"):
            if name != args.workload:
                side[name] = side_config(D, name, args, peak)
                torch.cuda.empty_cache()

    if rank == 0:
        steps = args.steps
        sbytes, stage_ms = m["sbytes"], m["stage_ms"]
        stages = stage_table(m, steps, peak)
        hbm_stages = ("pyramid", "fast", "blur")
        dom = max(hbm_stages, key=lambda s: stage_ms[s])
        kernel_name = {"pyramid": "resize_level_pre_kernel", "fast": "fast_tiles_kernel", "blur": "blur_all_kernel"}[dom]
        # one "launch" of the stage = one pass of the library over pass_frames frames (the FAST stage is one kernel per pass)
        n_launch = max(1, -(-frames_per_gpu // m["pass_frames"]) * steps)
        avg_launch_s = stage_ms[dom] * 1e-3 / n_launch
        bytes_per_launch = sbytes[dom] * frames_per_gpu * steps / n_launch
        achieved = bytes_per_launch / avg_launch_s / 1e9
        traffic, ncu_pipes, alu = None, None, None
        frames_per_launch = min(m["pass_frames"], frames_per_gpu)
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(kernel_name, {})
            fpp = prof.get("frames_per_pass", 512)
            if prof.get("dram_bytes_per_launch") is not None:  # the capture is of 512-frame passes: bytes per frame x frames per launch here
                traffic = prof["dram_bytes_per_launch"] / fpp * frames_per_launch
            ncu_pipes = {k: prof[k] for k in ("alu_pipe_pct", "issue_active_pct", "dram_throughput_pct") if k in prof} or None
            if prof.get("alu_warp_inst_per_launch") is not None:
                # what bounds this kernel: ALU-pipe warp-instructions of one launch (ncu smsp__inst_executed_pipe_alu, committed
                # capture) / the launch's CUDA-event time in THIS run, against the pipe's rate measured in THIS run
                inst = prof["alu_warp_inst_per_launch"] / fpp * frames_per_launch
                alu = {"bound": "alu", "achieved": inst / avg_launch_s, "peak": alu_s, "unit": "warp-instructions/s", "frac": inst / avg_launch_s / alu_s,
                       "peak_per_clk_per_sm": alu_clk, "peak_source": "measured in this run (VIMNMX3.U16x2 probe, kernels_probe.cu)",
                       "alu_warp_inst_per_launch": inst, "instructions_source": "ncu smsp__inst_executed_pipe_alu.sum, profiles/traffic.json"}
            if prof.get("warp_inst_per_launch") is not None and alu_clk > 0:
                # the second limit of an instruction-bound kernel: issue slots, one warp-instruction per clock per SM sub-partition
                # (4 / clk / SM); alu_s / alu_clk = SM clock x SMs as measured by the probe of this run.  An instruction with three
                # register sources (VIMNMX3, HFMA2) takes two slots on this SM (profiles/r2_pipe_probe_fma.log), which the plain
                # count below does not weigh.
                winst = prof["warp_inst_per_launch"] / fpp * frames_per_launch
                issue_peak = 4.0 * alu_s / alu_clk
                alu["issue"] = {"achieved": winst / avg_launch_s, "peak": issue_peak, "unit": "warp-instructions/s",
                                "frac": winst / avg_launch_s / issue_peak, "warp_inst_per_launch": winst,
                                "instructions_source": "ncu smsp__inst_executed.sum, profiles/traffic.json"}
        except Exception:
            pass
        fps = frames_global * steps / (m["ms_total"] * 1e-3)
        e2e_fps = frames_global * m["e2e_steps"] / m["e2e_s"]
        ceil = m.get("ceiling")
        e2e = {"value": e2e_fps, "unit": "frames/s", "steps": m["e2e_steps"], "h2d_bytes_per_step": m["h2d_bytes"],
               "d2h_bytes_per_step": m["d2h_bytes"], "matches_resident_run": m["same"],
               "surface": "the north-star 4-argument call: keypoints + descriptors + counts come back (no imagePyramid)",
               "bytes_are": "whole job (sum over the ranks)"}
        if ceil:
            e2e["platform_ceiling"] = ceil
            e2e["h2d_gbs"] = e2e_fps * w * h / 1e9
            e2e["frac_of_h2d_ceiling"] = e2e_fps * w * h / 1e9 / ceil["h2d_gbs"]
            # both directions together against the same figure of the bare copies (the leg moves frames up and results down at once)
            e2e["both_directions_gbs"] = (m["h2d_bytes"] + m["d2h_bytes"]) * m["e2e_steps"] / m["e2e_s"] / 1e9
            e2e["frac_of_both_directions_ceiling"] = e2e["both_directions_gbs"] / ceil["both_directions_wall_gbs"]
            e2e["frac_of_resident"] = e2e_fps / fps  # at N = 1 the kernels, not the copies, bound the leg
        line = {
            "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": steps, "warmup": max(args.warmup, 3),
            "ms_per_step": m["ms_total"] / steps, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic", "config": config_of(args.workload),
            "details": {"frames_per_gpu_per_step": frames_per_gpu, "frames_per_pass": m["pass_frames"], "frames_per_pass_e2e": m["e2e_pass_frames"],
                        "generator": "smooth_noise, %d distinct frames + column rotations (all bytes distinct)" % DISTINCT,
                        "l2": "batch (%.0f MB of frames per GPU per step) larger than the 126 MB L2; no flush" % (frames_per_gpu * w * h / 1e6),
                        "sharding": "ONE batch of %d frames split into contiguous ranges over the ranks (strong scaling); no collective in "
                                    "the loop; final NCCL gather to rank 0 timed separately" % frames_global,
                        "profiling": "headline loop runs without stage events; the per-stage table comes from a second loop with them",
                        "ms_per_step_with_stage_events": m["ms_profiled"] / steps},
            "e2e": e2e,
            "gpu_launches": int(m["launches"]),
            "clocks": m["clocks"],
            "roofline": {"bound": "hbm", "kernel": kernel_name, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "bytes_per_launch": bytes_per_launch, "avg_launch_ms": avg_launch_s * 1e3,
                         "frac_of_nominal_8000": achieved / 8000.0, "ncu_profile": ncu_pipes, "alu": alu,
                         # the whole pipeline against the bytes a single fused pass could not avoid, per GPU
                         "fused_lower_bound": {"bytes_per_frame": sbytes["fused_lower_bound"],
                                               "gbs": fps / world * sbytes["fused_lower_bound"] / 1e9,
                                               "frac": fps / world * sbytes["fused_lower_bound"] / 1e9 / peak}},
            "stages": stages,
            "keypoints_per_frame": m["mean_kp"],
            "hamming": {"value": pairs_per_s, "unit": "pairs/s", "pairs_per_frame_pair": pairs_local / max(npairs, 1),
                        "frame_pairs_per_gpu": npairs, "ms": match_ms,
                        # bound: the POPC pipe; match_kernel folds the 8 XOR words of a pair into 5 popcounts with carry-save adders
                        # (DESIGN.md section 4), the reference's loop needs 8.  Peak = the POPC probe of this run.
                        "popc_pipe": {"popc_per_pair": 5, "achieved_warp_instr_per_s": pairs_per_s / world * 5 / 32, "peak_warp_instr_per_s": popc_s,
                                      "peak_per_clk_per_sm": popc_clk, "frac": pairs_per_s / world * 5 / 32 / popc_s,
                                      "frac_if_8_popc_per_pair": pairs_per_s / world * 8 / 32 / popc_s,
                                      "peak_source": "measured in this run (POPC probe, kernels_probe.cu)"},
                        "distinctive_sets_per_s": distinctive_sets_per_s, "distinctive_set_size": 16,
                        "search_for_initialization_frame_pairs_per_s": search_pairs_per_s,
                        "search_for_initialization_matches_per_pair": search_matches},
            "pipe_probes": {"popc_per_clk_per_sm": popc_clk, "vimnmx3_u16x2_per_clk_per_sm": alu_clk, "prmt_per_clk_per_sm": prmt_clk,
                            "unit": "warp-instructions / clk / SM, measured in this run"},
            "orbslam2_mode": orbslam2,
            "gather": gather,
            "multi_gpu_identical": identical,
            "weak_scaling": weak,
            "host_affinity": D.numa,
        }
        if "e2e_pyr_s" in m:
            line["e2e_with_pyramid"] = {"value": frames_global * m["e2e_steps"] / m["e2e_pyr_s"], "unit": "frames/s",
                                        "h2d_bytes_per_step": m["h2d_bytes"], "d2h_bytes_per_step": m["d2h_pyr_bytes"],
                                        "surface": "the reference's 5-argument call (src/ORBextractor.cc:620-621): imagePyramid levels >= 1 of "
                                                   "every frame come back too (level 0 is the caller's own input)"}
        if side is not None:
            line["configs"] = side
        # single-frame synchronous latency of the reference-facing call (sdorb_extract: host image in, results on the host), the
        # way Frame.cc:195 uses the extractor; output buffers allocated once, as a C++ caller's are -- outside every timed region
        ex1 = api.ORBextractor(*params, device=D.local, max_width=w, max_height=h, max_batch=1)
        lat = {}
        for want in (False, True):
            call = ex1.single_frame_call(w, h, want_pyramid=want)
            for i in range(10):
                call(host_probe[i])
            ts = []
            for i in range(60):
                t0 = time.perf_counter()
                call(host_probe[i % len(host_probe)])
                ts.append(time.perf_counter() - t0)
            lat["with_pyramid_ms" if want else "keypoints_only_ms"] = float(np.median(ts) * 1e3)
        lat["with_pyramid_over_keypoints_only"] = lat["with_pyramid_ms"] / lat["keypoints_only_ms"]
        ex1.close()
        line["single_frame_latency"] = lat
        if world == 1 and not args.no_cpu:
            cores = os.cpu_count() or 1
            cx, kind, what = cpu_extractor(params)
            sample_n = max(16 * cores, 128)
            sample = make_frames(sample_n, w, h)
            cpu_extract_rate(cx, sample[:2 * cores], cores)
            fps_n, _ = cpu_extract_rate(cx, sample, cores)
            fps_1, _ = cpu_extract_rate(cx, sample[:24], 1)
            ham_n = cpu_match_rate(cores, max(cores, 8), np.random.default_rng(0))
            cv2_n = cv2_primitives_rate(params, sample[:max(2 * cores, 32)], cores)
            line["cpu_baseline"] = {"value": fps_n, "unit": "frames/s", "cores": cores, "kind": kind,
                                    "sample": "first %d frames of the same batch (%d per thread), %d threads frame-parallel; %s" % (
                                        sample_n, sample_n // cores, cores, what),
                                    "value_1core": fps_1, "sample_1core": "first 24 frames, 1 thread (the reference's execution model)",
                                    "hamming_pairs_per_s": ham_n,
                                    "cv2_primitives_only": {"value": cv2_n, "unit": "frames/s", "cores": cores,
                                                            "what": "OpenCV 4.13 SIMD resize + per-cell FAST + GaussianBlur only (no culling / "
                                                                    "orientation / descriptors): upper bound for the reference linked against a real OpenCV"}}
        print(json.dumps(line))
    D.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C3_tum_640x480_1000kp_8lv", choices=sorted(WORKLOADS))
    ap.add_argument("--frames", type=int, default=0, help="frames per step over all GPUs (default: the workload's)")
    ap.add_argument("--pass-frames", type=int, default=2048,
                    help="frames per internal pass of the library (max_batch) in the resident leg: longer launches lose less to "
                         "launch gaps and partial last waves")
    ap.add_argument("--e2e-pass-frames", type=int, default=768,
                    help="max_batch of the handle of the end-to-end leg: the host pipeline overlaps copies and kernels pass by pass")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-side", action="store_true", help="skip the side configurations (C2 / C5 at N = 1, weak scaling at N > 1)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
