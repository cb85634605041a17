#!/bin/bash
# 2-GPU run of bench.py at the final state (weak-scaling headline + the strong-scaling config-4 sweep leg), and the host
# mirror with the engines' internal half-batch split off (the mirror pipelines two engines itself)
set -u
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/r02g_bench_n2.json 2> gpurun_out/bench28_n2.err
tail -c 1500 gpurun_out/r02g_bench_n2.json; tail -3 gpurun_out/bench28_n2.err
echo "== PG_SPLIT=1"; PG_SPLIT=1 python tools/mirror_sweep.py 30 2>&1 | tee gpurun_out/mirror_sweep28_split1.txt | cut -c1-260
