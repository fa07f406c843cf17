#!/bin/bash
# FAST tile height sweep (run on the GPU box): rebuilds the library with -DSDORB_FAST_TH=<rows> and times the resident pass.
HEIGHTS=${1:-"60 48 40 32"}
for th in $HEIGHTS; do
  SDORB_NVCC_EXTRA="-DSDORB_FAST_TH=$th" python sdslam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed for $th"; continue; }
  grep -A3 "fast_tiles_kernel" sdslam_b200/build/ptxas.log | grep -E "Used" | head -1
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stages or golden or sweep or c5" 2>&1 | tail -1
  tools/ab_probe.sh SDORB_FAST_TH_BUILT "$th" "2048"
done
python sdslam_b200/build.py --force > /dev/null 2>&1
