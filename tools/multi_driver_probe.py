#!/usr/bin/env python3
"""The C-level multi-GPU driver (sdorb_extract_batch_multi) from ONE process: a pinned host batch of 4096 frames 640x480 over all
visible GPUs (one handle and one host thread per GPU inside the library), end to end, against the same batch on GPU 0 alone.
Prints frames/s and whether the results are byte-identical."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sdslam_b200 import api  # noqa: E402


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    ngpu = torch.cuda.device_count()
    params = (1000, 1.2, 8, 20)
    host = torch.from_numpy(bench.make_frames(nf, 640, 480)).pin_memory()
    cap = 1000
    outs = []
    for devs in ([0], list(range(ngpu))):
        exs = [api.ORBextractor(*params, device=d, max_width=640, max_height=480, max_batch=768) for d in devs]
        cap = exs[0].max_keypoints
        k = torch.zeros((nf, cap, 7), dtype=torch.float32).pin_memory()
        d_ = torch.zeros((nf, cap, 32), dtype=torch.uint8).pin_memory()
        c = torch.zeros(nf, dtype=torch.int32).pin_memory()
        kn = k.numpy().view(api.KP_DTYPE).reshape(nf, cap)
        api.extract_batch_multi(exs, host, kn, d_.numpy(), c.numpy())  # warm-up (allocations, first-touch)
        ts = []
        for _ in range(3):
            t = time.perf_counter()
            api.extract_batch_multi(exs, host, kn, d_.numpy(), c.numpy())
            ts.append(time.perf_counter() - t)
        outs.append((kn.tobytes(), d_.numpy().tobytes(), c.numpy().tobytes()))
        print("%d GPU(s), one process, %d frames through sdorb_extract_batch_multi: %.0f frames/s end to end (best of 3: %.2f ms)" % (
            len(devs), nf, nf / min(ts), min(ts) * 1e3), flush=True)
        for e in exs:
            e.close()
    print("multi-GPU result byte-identical to one GPU:", outs[0] == outs[1])
    return 0 if outs[0] == outs[1] else 1


if __name__ == "__main__":
    sys.exit(main())
