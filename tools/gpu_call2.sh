#!/bin/bash
# round 2, call 2: honest per-phase kernel times on all workload shapes (base vs speculative build), an ncu
# --set full capture of the speculative build's kernels on config 2 (serialised, PG_SPLIT=1), and the new GPU tests
set -u
mkdir -p gpurun_out
python tools/kernel_times.py > gpurun_out/kt_base.txt 2>&1; cat gpurun_out/kt_base.txt
PG_LIB=ab_build/libpg_spec_dead.so python tools/kernel_times.py > gpurun_out/kt_spec.txt 2>&1; cat gpurun_out/kt_spec.txt
PG_LIB=ab_build/libpg_spec_prune.so python tools/kernel_times.py > gpurun_out/kt_prune.txt 2>&1; cat gpurun_out/kt_prune.txt
PG_SPLIT=1 PG_LIB=ab_build/libpg_spec_dead.so timeout 600 ncu --set full --clock-control none --import-source on \
   -k regex:'pg_fill_kernel|pg_trace_kernel' -s 3 -c 3 -f -o gpurun_out/r02a_spec python tools/profile_run.py > gpurun_out/ncu_spec.log 2>&1
tail -3 gpurun_out/ncu_spec.log
ls -la gpurun_out/*.ncu-rep
timeout 900 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tests2.txt 2>&1; tail -5 gpurun_out/tests2.txt
