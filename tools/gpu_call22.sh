#!/bin/bash
# partial tile recomputation in the traceback + CK=32 and boundary-loop unroll variants on top of the lean node events
set -u
mkdir -p gpurun_out
python tools/kernel_times.py > gpurun_out/kt22_base.txt 2>&1; cat gpurun_out/kt22_base.txt
for v in ck32 fu1 fu2; do
PG_LIB=$PWD/ab_build/libpg_$v.so python tools/kernel_times.py > gpurun_out/kt22_$v.txt 2>&1; echo $v; cat gpurun_out/kt22_$v.txt
done
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests22.txt 2>&1; tail -3 gpurun_out/tests22.txt
PG_LIB=$PWD/ab_build/libpg_ck32.so timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests22_ck32.txt 2>&1; tail -3 gpurun_out/tests22_ck32.txt
PG_LIB=$PWD/ab_build/libpg_ck32.so timeout 300 python tools/gpu_fuzz.py 300 48 91 > gpurun_out/gpu_fuzz22_ck32.txt 2>&1; tail -1 gpurun_out/gpu_fuzz22_ck32.txt
