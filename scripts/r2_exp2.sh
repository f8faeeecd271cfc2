#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_gpu_rows.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_transition.py tests/test_gpu_crossload.py tests/test_gpu_api.py -x -q > gpurun_out/r2_t3.log 2>&1
tail -15 gpurun_out/r2_t3.log
timeout 200 env CGPM_B200_ROWS_PROF=1 python scripts/probe_rows.py 100000 50 60 4 4 > gpurun_out/r2_probe3.log 2>&1
grep -E "sweep|move rate|^\{|final" gpurun_out/r2_probe3.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err
tail -c 3000 gpurun_out/r2_bench1.json; tail -5 gpurun_out/r2_bench1.err
