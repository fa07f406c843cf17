#!/usr/bin/env python3
"""Prints the host pipeline's timeline (SDORB_PIPE_TRACE=1) of one 4096-frame end-to-end call:  python tools/pipe_trace.py [frames] [max pass]"""
import os
import sys

os.environ["SDORB_PIPE_TRACE"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import bench  # noqa: E402
from sdslam_b200 import api  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
mp = int(sys.argv[2]) if len(sys.argv) > 2 else 768
imgs = torch.from_numpy(bench.make_frames(nf, 640, 480)).pin_memory()
ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=mp)
cap = ex.max_keypoints
k = torch.zeros((nf, cap, 7), dtype=torch.float32).pin_memory()
d = torch.zeros((nf, cap, 32), dtype=torch.uint8).pin_memory()
c = torch.zeros(nf, dtype=torch.int32).pin_memory()
for i in range(3):  # the third timeline is the warm one
    ex.extract_batch_host(imgs, k, d, c)
