#!/bin/bash
# is config 5's gain from CK = 32 the halved scratch (24 GB limit -> chunking)?  CK = 16 / 32 with 24 / 64 / 100 GB of scratch
set -u
mkdir -p gpurun_out
for gb in 24 64 100; do
echo "CK=16 scratch $gb GB"; PG_SCRATCH_GB=$gb python tools/kernel_times.py config5 2>&1 | tee -a gpurun_out/kt25.txt
echo "CK=32 scratch $gb GB"; PG_SCRATCH_GB=$gb PG_LIB=$PWD/ab_build/libpg_ck32.so python tools/kernel_times.py config5 2>&1 | tee -a gpurun_out/kt25.txt
done
