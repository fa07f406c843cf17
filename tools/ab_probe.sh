#!/bin/bash
# A/B of one environment switch of the library on the resident / e2e / latency legs: tools/ab_probe.sh VAR "v1 v2 ..." [pass sizes]
VAR=$1; VALS=$2; PASSES=${3:-"512 2048"}
for pf in $PASSES; do
  for v in $VALS; do
    env $VAR=$v python bench.py --no-cpu --no-side --steps 6 --pass-frames $pf 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
s=d['stages']
print('pass_frames $pf  $VAR=$v : resident %.0f frames/s (%.3f ms/step)  e2e %.0f  stages pyr %.3f fast %.3f sel %.3f blur %.3f desc %.3f  latency %.3f / %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], s['pyramid']['ms_per_step'], s['fast']['ms_per_step'], s['select']['ms_per_step'], s['blur']['ms_per_step'], s['describe']['ms_per_step'], d['single_frame_latency']['keypoints_only_ms'], d['single_frame_latency']['with_pyramid_ms']))"
  done
done
