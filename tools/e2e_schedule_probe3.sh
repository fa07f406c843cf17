#!/bin/bash
run() {
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-cpu --no-side --steps 3 --e2e-steps 6 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-44s resident %.0f  e2e %.0f frames/s  e2e_with_pyramid %.0f' % ('$label', d['value'], d['e2e']['value'], d['e2e_with_pyramid']['value']))"
}
run "default (1.12)" X=0 --
run "dual lanes" SDORB_PIPE_DUAL=1 --
run "dual lanes, growth 1.05" SDORB_PIPE_DUAL=1 SDORB_PIPE_GROWTH=105 --
run "dual lanes, const 256" SDORB_PIPE_DUAL=1 SDORB_PIPE_CONST=256 --
run "dual lanes, const 384" SDORB_PIPE_DUAL=1 SDORB_PIPE_CONST=384 --
run "dual lanes, growth 1.12, first 128, max 512" SDORB_PIPE_DUAL=1 SDORB_PIPE_MIN=128 -- --e2e-pass-frames 512
run "first 128, max 512" SDORB_PIPE_MIN=128 -- --e2e-pass-frames 512
run "PDL on passes <= 1024" SDORB_PDL_MAX_FRAMES=1024 --
