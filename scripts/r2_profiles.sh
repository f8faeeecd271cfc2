#!/bin/bash
# round 2 ncu captures (one GPU).  Outputs under gpurun_out/; summaries are copied to profiles/ by scripts/summarize_profiles.py
cd "$(dirname "$0")/.."
python scripts/profile_kernels.py 64 12 > gpurun_out/r2_prof_plain.log 2>&1 || exit 1
tail -2 gpurun_out/r2_prof_plain.log
# launch list of a bench-like step (serialised, cold-cache per-launch times)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_c1.csv python scripts/profile_kernels.py 64 4 > gpurun_out/r2_ncu_launch.log 2>&1
for k in rows_kernel aux_kernel stats_kernel scan_kernel col_hyper_kernel logpdf_kernel simulate_kernel; do
  skip=0
  [ "$k" = rows_kernel ] && skip=12
  [ "$k" = aux_kernel ] && skip=12
  [ "$k" = stats_kernel ] && skip=25
  [ "$k" = scan_kernel ] && skip=12
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -o gpurun_out/r02_$k python scripts/profile_kernels.py 64 12 > gpurun_out/r2_ncu_$k.log 2>&1
  tail -1 gpurun_out/r2_ncu_$k.log
done
