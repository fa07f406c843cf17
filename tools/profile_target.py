#!/usr/bin/env python3
"""The program the ncu captures of a round run against (tools/profile_r2.sh): 512 resident 640x480 frames through one library
pass three times (the bench workload's kernels at the pass size the captures of round 1 used), then the matcher kernels on the
descriptors that pass produced: 256 frame pairs of 1000 x 1000 (match_kernel), ComputeDistinctiveDescriptors, and the single-frame
call.  Short on purpose: ncu replays every captured kernel ~40 times."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from sdslam_b200 import api  # noqa: E402


def main():
    nf = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    dev = torch.device("cuda:0")
    imgs = torch.from_numpy(bench.make_frames(nf, 640, 480)).to(dev)
    ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=nf)
    cap = ex.max_keypoints
    kps = torch.zeros((nf, cap, 7), dtype=torch.float32, device=dev)
    desc = torch.zeros((nf, cap, 32), dtype=torch.uint8, device=dev)
    cnt = torch.zeros(nf, dtype=torch.int32, device=dev)
    st = torch.cuda.Stream(dev)
    torch.cuda.set_stream(st)
    for _ in range(3):
        ex.extract_batch_device(imgs, kps, desc, cnt, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ex.batch_status()
    npairs = nf // 2
    dA, dB = desc[0::2].contiguous(), desc[1::2].contiguous()
    nA, nB = cnt[0::2].contiguous(), cnt[1::2].contiguous()
    out = torch.zeros((npairs, cap, 4), dtype=torch.int32, device=dev)
    for _ in range(3):
        ex.match_batch(dA, nA, dB, nB, out=out, device=True, stream=st.cuda_stream)
    nsets = nf * (cap // 16)
    off = torch.arange(0, 16 * nsets + 1, 16, dtype=torch.int32, device=dev)
    bi = torch.zeros(nsets, dtype=torch.int32, device=dev)
    bm = torch.zeros(nsets, dtype=torch.int32, device=dev)
    ex.distinctive_batch(desc.reshape(-1, 32), off, bi, bm, device=True, stream=st.cuda_stream)
    torch.cuda.synchronize()
    ex1 = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=1)
    call = ex1.single_frame_call(640, 480, want_pyramid=True)
    host = imgs[:4].cpu().numpy()
    for i in range(4):
        call(host[i])
    print("profile target ok: %d frames, %.1f keypoints per frame, %d accepted matches" % (nf, float(cnt.float().mean()), int(out[..., 3].sum())))
    ex.close()
    ex1.close()


if __name__ == "__main__":
    main()
