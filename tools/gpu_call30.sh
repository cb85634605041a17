#!/bin/bash
# the rare general merge of node_event_pre out of line (__noinline__), with and without a register cap; boundary loop unroll 2
set -u
mkdir -p gpurun_out
for v in inline out_mb5 out_fu2; do
echo "== $v"; PG_LIB=$PWD/ab_build/libpg_$v.so python tools/kernel_times.py config2 config4_share 2>&1 | tee gpurun_out/kt30_$v.txt
done
echo "== outline (no cap)"; python tools/kernel_times.py config2 config4_share 2>&1 | tee gpurun_out/kt30_outline.txt
