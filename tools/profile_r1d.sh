#!/bin/bash
# ncu captures behind profiles/r1d_* (run on the GPU box through gpurun; every ncu run follows a plain run that exited 0)
set -x
CMD="python bench.py --steps 2 --warmup 3 --frames 512 --pass-frames 512 --e2e-pass-frames 512 --no-cpu --e2e-steps 1"
timeout 300 $CMD > gpurun_out/ncu_plain_r1d.json 2> gpurun_out/ncu_plain_r1d.err || exit 1
if [ "$1" != "rest" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1d.csv $CMD > gpurun_out/ncu_r1d_a.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"resize_level|fast_tiles_kernel|gather_cells_kernel" -c 10 -f -o gpurun_out/prof_r1d $CMD > gpurun_out/ncu_r1d_b.log 2>&1
fi
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"select_kernel|blur_all_kernel|describe_kernel" -c 3 -f -o gpurun_out/prof_r1d_2 $CMD > gpurun_out/ncu_r1d_c.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"octree_kernel" -c 1 -f -o gpurun_out/prof_r1d_3 $CMD > gpurun_out/ncu_r1d_d.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
