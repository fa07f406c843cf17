#!/bin/bash
# compute-sanitizer memcheck of the whole extraction + matching path on a small batch (one tool per call).
# NOTE: on this pool gpurun answers "compute-sanitizer is closed" (2026-10-18); kept for pools where it is open.
set -e
cd "$(dirname "$0")/.."
compute-sanitizer --tool "${1:-memcheck}" --error-exitcode 3 python - <<'PY'
import numpy as np
from sdslam_b200 import api, synth
ex = api.ORBextractor(1000, 1.2, 8, 20, max_width=640, max_height=480, max_batch=3)
imgs = synth.frames(4, 640, 480)
k, d, c = ex.extract_batch_host(imgs)
kk, dd, pyr = ex(synth.rects(1, 333, 257))
m = ex.match_batch(d[0:1], c[0:1].astype(np.int32), d[1:2], c[1:2].astype(np.int32))
g = ex.match_batch(d[0:1], c[0:1].astype(np.int32), d[1:2], c[1:2].astype(np.int32), greedy=True)
hm = ex.hamming_matrix(d[0, :100], d[1, :90])
e = ex.debug_nth_element(((np.arange(500, dtype=np.uint32) << 8) | (np.arange(500, dtype=np.uint32) * 7 % 50 + 1)), 100)
print("sanitizer run ok", c.tolist(), len(kk), int(m["accepted"].sum()), int(g["accepted"].sum()), hm.shape)
PY
