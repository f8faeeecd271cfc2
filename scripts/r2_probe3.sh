#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export CGPM_B200_LIB=$PWD/cgpm_b200/libcgpm_b200_phase.so
for cfg in "1 8" "1 4" "2 8"; do echo "== D K = $cfg"; timeout 120 python scripts/probe_movers.py $cfg 32 3 2>&1 | tail -6; done > gpurun_out/p_movers.log 2>&1
cat gpurun_out/p_movers.log
