#!/bin/bash
# A/B of compile-time switches of the library (run on the GPU box): tools/define_probe.sh "<flags of variant 1>" "<flags of variant 2>" ...
# Each variant is built with SDORB_NVCC_EXTRA=<flags>, checked with the staged parity tests and timed on the resident pass.
for flags in "$@"; do
  SDORB_NVCC_EXTRA="$flags" python sdslam_b200/build.py --force > /dev/null 2>&1 || { echo "build failed for [$flags]"; continue; }
  grep -A3 "fast_tiles_kernel" sdslam_b200/build/ptxas.log | grep -E "Used" | head -1
  python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "stages or golden or sweep or c5" 2>&1 | tail -1
  tools/ab_probe.sh SDORB_BUILT_WITH "[$flags]" "2048" | sed 's/  e2e.*stages/  stages/'
done
python sdslam_b200/build.py --force > /dev/null 2>&1
