#!/bin/bash
# round 2, call 4: occupancy of the fill kernel on many-node graphs: warps per CTA chosen by residency (new default) vs 4
# (old), and the tables-in-HBM instantiation
set -u
mkdir -p gpurun_out
for v in "" "PG_FILL_WARPS_RT=4" "PG_FILL_WARPS_RT=2" "PG_FORCE_TABG=1"; do
  echo "== $v" | tee -a gpurun_out/kt4.txt
  env $v python tools/kernel_times.py config4_share config3 config2 >> gpurun_out/kt4.txt 2>&1
done
cat gpurun_out/kt4.txt
