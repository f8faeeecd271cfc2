#!/bin/bash
cd "$(dirname "$0")/.."
for cfg in "1 8" "2 8" "1 3" "4 6"; do echo "== D K = $cfg"; python scripts/probe_movers.py $cfg 32 3 2>&1 | tail -3; done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 2 -c 1 -o gpurun_out/r02_movers python scripts/probe_movers.py 1 8 32 3 > gpurun_out/r2_ncu_movers.log 2>&1
tail -2 gpurun_out/r2_ncu_movers.log
