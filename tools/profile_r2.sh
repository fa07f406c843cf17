#!/bin/bash
# ncu captures behind profiles/r2_* (run on the GPU box through gpurun).  Every ncu run follows a plain run of the same command
# that exited 0.  Outputs: gpurun_out/launches_r2.csv (per-launch device time, cold cache: compare SHARES), gpurun_out/prof_r2*.ncu-rep
# (ncu --set full).  tools/merge_profiles.py r2 turns them into profiles/r2_ncu_full_summary.csv + profiles/traffic.json.
set -x
CMD="python tools/profile_target.py 512"
timeout 300 $CMD > gpurun_out/profile_target_plain.log 2>&1 || { tail -5 gpurun_out/profile_target_plain.log; exit 1; }
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_r2_a.log 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"resize_level|fast_tiles_kernel|gather_cells_kernel|select_kernel|blur_all_kernel|describe_kernel" -s 12 -c 12 -f -o gpurun_out/prof_r2 $CMD > gpurun_out/ncu_r2_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"match_kernel|distinctive_kernel" -s 1 -c 3 -f -o gpurun_out/prof_r2_2 $CMD > gpurun_out/ncu_r2_c.log 2>&1
ls -la gpurun_out/*r2*.ncu-rep
tail -3 gpurun_out/ncu_r2_b.log
