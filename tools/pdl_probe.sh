#!/bin/bash
# Resident throughput with programmatic dependent launch switched on for passes up to N frames (SDORB_PDL_MAX_FRAMES), at two pass sizes.
for pf in 512 2048; do
  for pdl in 0 64 100000; do
    SDORB_PDL_MAX_FRAMES=$pdl python bench.py --no-cpu --no-side --steps 6 --pass-frames $pf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pass_frames $pf  pdl_max_frames $pdl : resident %.0f frames/s  (with stage events %.2f ms/step, without %.2f)  e2e %.0f  e2e+pyr %.0f  latency %s' % (d['value'], d['details']['ms_per_step_with_stage_events'], d['ms_per_step'], d['e2e']['value'], d['e2e_with_pyramid']['value'], d['single_frame_latency']))"
  done
done
