#!/bin/bash
# checkpoint interval decoupled from the traceback tile size (TS = 16): CK = 16 / 32 / 64
set -u
mkdir -p gpurun_out
python tools/kernel_times.py > gpurun_out/kt24_ck16.txt 2>&1; cat gpurun_out/kt24_ck16.txt
for v in ck32 ck64; do
PG_LIB=$PWD/ab_build/libpg_$v.so python tools/kernel_times.py > gpurun_out/kt24_$v.txt 2>&1; echo $v; cat gpurun_out/kt24_$v.txt
PG_LIB=$PWD/ab_build/libpg_$v.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests24_$v.txt 2>&1; tail -n 2 gpurun_out/tests24_$v.txt
done
PG_LIB=$PWD/ab_build/libpg_ck32.so timeout 300 python tools/gpu_fuzz.py 300 48 92 > gpurun_out/gpu_fuzz24_ck32.txt 2>&1; tail -n 1 gpurun_out/gpu_fuzz24_ck32.txt
