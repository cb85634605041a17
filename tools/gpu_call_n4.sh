#!/bin/bash
# 4-GPU run of bench.py: weak-scaling headline + the strong-scaling config-4 sweep
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r02h_bench_n4.json 2> gpurun_out/bench_n4.err
tail -c 900 gpurun_out/r02h_bench_n4.json; tail -3 gpurun_out/bench_n4.err
