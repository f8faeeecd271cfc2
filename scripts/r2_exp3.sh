#!/bin/bash
cd "$(dirname "$0")/.."
timeout 1500 python -m pytest tests/test_gpu_rows.py tests/test_gpu_edge.py tests/test_gpu_golden.py tests/test_gpu_transition.py tests/test_gpu_crossload.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2_t4.log 2>&1
tail -8 gpurun_out/r2_t4.log
python scripts/probe_bench_rows.py 12 2>&1 | tail -16
python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['kernel_ms_per_step'], d['config']['clusters_per_view_mean'], d['config']['clusters_per_view_max'], d['e2e']['ms_per_step'], d['value'])"
