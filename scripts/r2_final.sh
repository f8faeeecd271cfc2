#!/bin/bash
# round 2, final code: GPU test suite, the driver's bench command, launch list + full capture of the rows kernel (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests -m gpu -x -q --durations=12 ) > gpurun_out/f_tests.log 2>&1
tail -22 gpurun_out/f_tests.log
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
tail -3 gpurun_out/f_bench.err
timeout 300 python scripts/profile_kernels.py 64 12 > gpurun_out/f_prof_plain.log 2>&1 || { tail -5 gpurun_out/f_prof_plain.log; exit 1; }
tail -1 gpurun_out/f_prof_plain.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02f_launches_c2.csv python scripts/profile_kernels.py 64 4 > gpurun_out/f_ncu_launch.log 2>&1
tail -1 gpurun_out/f_ncu_launch.log
timeout 400 ncu --set full --clock-control none --import-source on -k regex:rows_kernel -s 12 -c 1 -o gpurun_out/r02f_rows_kernel python scripts/profile_kernels.py 64 12 > gpurun_out/f_ncu_rows.log 2>&1
tail -1 gpurun_out/f_ncu_rows.log
ls -la gpurun_out
