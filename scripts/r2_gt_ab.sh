#!/bin/bash
# A/B of the global-table one-row scoring (score_candidate_g): GPU tests, then the configs[2] shape at 100k rows with the
# shared tables capped (wide views on global tables) against the previous build (scripts/ubench/libcgpm_b200_base.so)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q ) > gpurun_out/gt_tests.log 2>&1
tail -4 gpurun_out/gt_tests.log
export CGPM_B200_ROWS_SMEM_KB=64
CGPM_B200_LIB=$PWD/scripts/ubench/libcgpm_b200_base.so timeout 300 python scripts/probe_c2.py 8 100000 5 full > gpurun_out/gt_probe_base.log 2>&1
timeout 300 python scripts/probe_c2.py 8 100000 5 full > gpurun_out/gt_probe_new.log 2>&1
grep -h "^sweep" gpurun_out/gt_probe_base.log | cut -c1-330
grep -h "^sweep" gpurun_out/gt_probe_new.log | cut -c1-330
