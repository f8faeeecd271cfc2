#!/bin/bash
# round 2, experiment 1: parity tests, rows probe at fixed speculation levels, one ncu capture
cd "$(dirname "$0")/.."
timeout 1100 python -m pytest tests/test_gpu_rows.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_transition.py -x -q > gpurun_out/r2_t2.log 2>&1
tail -15 gpurun_out/r2_t2.log
for lvl in a 0 1 2 3 4 5; do
  if [ "$lvl" = a ]; then unset CGPM_B200_ROWS_LVL; else export CGPM_B200_ROWS_LVL=$lvl; fi
  echo "=== LVL $lvl" >> gpurun_out/r2_probe2.log
  timeout 200 env CGPM_B200_ROWS_PROF=1 python scripts/probe_rows.py 100000 50 60 4 4 >> gpurun_out/r2_probe2.log 2>&1
done
unset CGPM_B200_ROWS_LVL
grep -E "LVL|sweep 3|move rate|^\{" gpurun_out/r2_probe2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 2 -c 1 -o gpurun_out/r2_rows_v1 python scripts/probe_rows.py 100000 50 60 4 3 > gpurun_out/r2_ncu1.log 2>&1
tail -3 gpurun_out/r2_ncu1.log
