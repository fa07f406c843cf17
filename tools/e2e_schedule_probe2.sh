#!/bin/bash
run() {
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-cpu --no-side --steps 3 --e2e-steps 6 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-44s resident %.0f  e2e %.0f frames/s  e2e_with_pyramid %.0f' % ('$label', d['value'], d['e2e']['value'], d['e2e_with_pyramid']['value']))"
}
for g in 105 108 110 112 115 120; do run "growth $g" SDORB_PIPE_GROWTH=$g --; done
for g in 108 112; do run "growth $g, first pass 48" SDORB_PIPE_GROWTH=$g SDORB_PIPE_MIN=48 --; done
for g in 108 112; do run "growth $g, first pass 64" SDORB_PIPE_GROWTH=$g SDORB_PIPE_MIN=64 --; done
for g in 108 112; do run "growth $g, max pass 1024" SDORB_PIPE_GROWTH=$g -- --e2e-pass-frames 1024; done
run "growth 110, first pass 128" SDORB_PIPE_GROWTH=110 SDORB_PIPE_MIN=128 --
run "growth 110 again" SDORB_PIPE_GROWTH=110 --
