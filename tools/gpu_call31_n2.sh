#!/bin/bash
# N = 2 bench of HEAD: weak-scaling headline + the strong-scaling sweep with its tensor gather on NCCL
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02h_bench_n2.json 2> gpurun_out/bench31_n2.err
tail -c 900 gpurun_out/r02h_bench_n2.json; tail -3 gpurun_out/bench31_n2.err
