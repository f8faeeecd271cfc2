#!/bin/bash
# the driver's N = 2 launch with fewer steps: does configs[2] through ShardedEngine fit the time limit?
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
( time timeout 800 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 4 --warmup 2 ) > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err
tail -5 gpurun_out/n2_bench.err
cut -c1-1500 gpurun_out/n2_bench.json
