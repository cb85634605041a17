#!/bin/bash
# lean node events (entry words + prefetched seeds): A/B against the same source built with PG_LEAN_EVENTS=0, GPU fuzz, suite
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build21.txt 2>&1
PG_LIB=$PWD/ab_build/libpg_nolean.so python tools/kernel_times.py > gpurun_out/kt21_nolean.txt 2>&1; cat gpurun_out/kt21_nolean.txt
python tools/kernel_times.py > gpurun_out/kt21_lean.txt 2>&1; cat gpurun_out/kt21_lean.txt
timeout 300 python tools/gpu_fuzz.py 400 48 77 > gpurun_out/gpu_fuzz21.txt 2>&1; tail -2 gpurun_out/gpu_fuzz21.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests21.txt 2>&1; tail -3 gpurun_out/tests21.txt
