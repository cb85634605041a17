#!/bin/bash
# profile of the short-node shape after the lean node events: one DP step of the config-4 share under ncu --set full,
# per-SASS execution counts of the forward fill launch and of the traceback
set -u
mkdir -p gpurun_out /tmp/ncu
M="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none --import-source on -k regex:'pg_' -s 8 -c 8 -f -o /tmp/ncu/config4_step python tools/profile_run.py config4_share > gpurun_out/ncu23.log 2>&1; tail -1 gpurun_out/ncu23.log
python tools/ncu_step_summary.py /tmp/ncu/config4_step.ncu-rep gpurun_out/r02g_config4_step_ncu.json "config-4 share (1 250 vcf2paragraph-shaped sites, 96k reads), PG_SPLIT=1, one DP step (tools/profile_run.py config4_share), lean node events"
ncu -i /tmp/ncu/config4_step.ncu-rep --page source --csv --kernel-name regex:pg_fill_kernel --launch-skip 0 --launch-count 1 > /tmp/ncu/fill_src.csv 2>/dev/null; gzip -c /tmp/ncu/fill_src.csv > gpurun_out/r02g_config4_fill_fwd_source.csv.gz
ncu -i /tmp/ncu/config4_step.ncu-rep --page source --csv --kernel-name regex:pg_trace_kernel > /tmp/ncu/trace_src.csv 2>/dev/null; gzip -c /tmp/ncu/trace_src.csv > gpurun_out/r02g_config4_trace_source.csv.gz
ncu -i /tmp/ncu/config4_step.ncu-rep --page raw --csv > gpurun_out/r02g_config4_step_raw.csv 2>/dev/null
ls -la gpurun_out | tail -8
python tools/kernel_times.py config4_share
