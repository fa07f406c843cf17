#!/bin/bash
# End-to-end leg of the bench under different host-pipeline schedules (run on the GPU box).  One line per setting.
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" python bench.py --no-cpu --no-side --steps 3 --e2e-steps 6 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('%-44s resident %.0f  e2e %.0f frames/s  e2e_with_pyramid %.0f' % ('$label', d['value'], d['e2e']['value'], d['e2e_with_pyramid']['value']))"
}
run "default (growth 1.25, max pass 768)" X=0 --
run "growth 1.10" SDORB_PIPE_GROWTH=110 --
run "growth 1.01" SDORB_PIPE_GROWTH=101 --
run "taper" SDORB_PIPE_TAPER=1 --
run "growth 1.10 + taper" SDORB_PIPE_GROWTH=110 SDORB_PIPE_TAPER=1 --
run "const 192" SDORB_PIPE_CONST=192 --
run "const 256" SDORB_PIPE_CONST=256 --
run "const 384" SDORB_PIPE_CONST=384 --
run "const 512" SDORB_PIPE_CONST=512 --
run "max pass 512, growth 1.10" SDORB_PIPE_GROWTH=110 -- --e2e-pass-frames 512
run "max pass 384, growth 1.10" SDORB_PIPE_GROWTH=110 -- --e2e-pass-frames 384
run "const 256, 4 slots" SDORB_PIPE_CONST=256 SDORB_PIPE_SLOTS=4 --
