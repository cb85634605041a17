#!/bin/bash
# node-boundary sub-blocks with the collapsed recurrence (PG_DEAD_BOUNDARY): A/B against the same source without it, register
# caps (launch bounds: 4 / 5 CTAs per SM), unroll 4
set -u
mkdir -p gpurun_out
for v in nodeadb deadb_mb5 deadb_mb4 deadb_u4; do
echo "== $v"; PG_LIB=$PWD/ab_build/libpg_$v.so python tools/kernel_times.py 2>&1 | tee gpurun_out/kt29_$v.txt
done
echo "== default (no cap)"; python tools/kernel_times.py 2>&1 | tee gpurun_out/kt29_deadb.txt
PG_LIB=$PWD/ab_build/libpg_deadb_mb5.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests29_mb5.txt 2>&1; tail -n 2 gpurun_out/tests29_mb5.txt
PG_LIB=$PWD/ab_build/libpg_deadb_mb5.so timeout 300 python tools/gpu_fuzz.py 400 48 94 > gpurun_out/gpu_fuzz29.txt 2>&1; tail -n 1 gpurun_out/gpu_fuzz29.txt
