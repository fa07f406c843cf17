#!/usr/bin/env python3
"""Per-stage device timing of the ORB-SLAM2-style mode (row f1) on a resident batch (development aid)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdslam_b200 import api, synth  # noqa: E402

dev = torch.device("cuda:0")
imgs = torch.from_numpy(np.concatenate([synth.frames(16, 640, 480)] * 32)).to(dev)
ex = api.ORBextractor(1000, 1.2, 8, 20, minThFAST=7, max_width=640, max_height=480, max_batch=512)
cap = ex.max_keypoints
k = torch.zeros((512, cap, 7), dtype=torch.float32, device=dev)
d = torch.zeros((512, cap, 32), dtype=torch.uint8, device=dev)
c = torch.zeros(512, dtype=torch.int32, device=dev)
for _ in range(2):
    ex.extract_batch_device(imgs, k, d, c)
torch.cuda.synchronize()
ex.set_profiling(True)
ex.stage_times(reset=True)
for _ in range(5):
    ex.extract_batch_device(imgs, k, d, c)
torch.cuda.synchronize()
ms, _ = ex.stage_times()
print({a: round(b / 5, 3) for a, b in ms.items()}, "keypoints/frame", float(c.float().mean()))
ex.close()
