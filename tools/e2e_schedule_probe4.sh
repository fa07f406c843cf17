#!/bin/bash
# pass growth of the host pipeline on the other workloads
for wl in C2_euroc_752x480_1000kp_8lv C5_1080p_4000kp_12lv C3_tum_640x480_1000kp_8lv; do
  for g in 125 112 105; do
    extra=""; [ "$wl" = "C5_1080p_4000kp_12lv" ] && extra="--frames 512 --pass-frames 256 --e2e-pass-frames 128"
    SDORB_PIPE_GROWTH=$g python bench.py --workload $wl --no-cpu --no-side --steps 3 --e2e-steps 6 $extra 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-30s growth %s  resident %.0f  e2e %.0f  e2e_with_pyramid %.0f  clocks %s' % ('$wl', '$g', d['value'], d['e2e']['value'], d['e2e_with_pyramid']['value'], d['clocks']))"
  done
done
