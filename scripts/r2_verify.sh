#!/bin/bash
# full GPU test suite + the driver's bench command (one GPU)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q --durations=15 ) > gpurun_out/v_tests.log 2>&1
tail -25 gpurun_out/v_tests.log
( time python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
tail -3 gpurun_out/v_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/v_bench.json').read().strip().splitlines()[0])
print(d['ms_per_step'], d['kernel_ms_per_step'], d['value'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e_resident']['ms_per_step'])
print(d['roofline_compute']['cycles_per_row_critical_cta'], d['config'].get('clusters_per_view_mean'), d['cpu_baseline'])
for k,v in d.get('extra',{}).items(): print(k, json.dumps(v)[:600])
PY
