#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 240 python scripts/probe_steps.py 15 25 > gpurun_out/p_steps.log 2>&1
tail -30 gpurun_out/p_steps.log
timeout 420 python scripts/probe_c2.py 8 1000000 5 full > gpurun_out/p_c2.log 2>&1
tail -30 gpurun_out/p_c2.log
