#!/bin/bash
# 2-GPU run of bench.py (weak-scaling headline + the strong-scaling config-4 sweep leg)
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err
tail -c 2500 gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/bench_n1b.json 2> gpurun_out/bench_n1b.err; tail -c 600 gpurun_out/bench_n1b.json
