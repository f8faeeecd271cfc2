#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CGPM_B200_ROWS_DL=256 PROBE_PSEED=4242 PROBE_SEED=11 timeout 400 python scripts/probe_c2.py 8 1000000 5 full > gpurun_out/p_c2_dl256.log 2>&1
cut -c1-1500 gpurun_out/p_c2_dl256.log
