#!/bin/bash
# one complete DP step of config 2 under ncu --set full (9 launches since the traceback of the reads that wait for a
# second-round fill is a launch of its own): the r02g capture window started one launch early
set -u
mkdir -p gpurun_out /tmp/ncu
M="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none --import-source on -k regex:'pg_' -s 9 -c 9 -f -o /tmp/ncu/config2_step python tools/profile_run.py config2 > gpurun_out/ncu39.log 2>&1; tail -1 gpurun_out/ncu39.log
python tools/ncu_step_summary.py /tmp/ncu/config2_step.ncu-rep gpurun_out/r02h_step_ncu.json "config 2 (10 000 reads), PG_SPLIT=1, one complete DP step of HEAD: forward fill, plan, pair, paired reversed-graph fill, plan, pair, second-round fill, traceback of the settled reads, traceback of the pending reads (tools/profile_run.py config2; ncu --set full --clock-control none); lean node events, CK = 32, TS = 16"
ncu -i /tmp/ncu/config2_step.ncu-rep --page source --csv --kernel-name regex:pg_trace_kernel --launch-skip 0 --launch-count 1 > /tmp/ncu/trace_src.csv 2>/dev/null; gzip -c /tmp/ncu/trace_src.csv > gpurun_out/r02h_trace_source.csv.gz
python - <<'P'
import json
d = json.load(open("gpurun_out/r02h_step_ncu.json"))
for l in d["launches"]:
    print(l["kernel"][:40], l["grid"], l["block"], round(l["duration_us"], 1), "inst %.3g" % l["inst_executed"], "alu%% %.1f issue%% %.1f regs %d" % (l["alu_pipe_pct_of_peak_active"], l["issue_active_pct"], l["registers"]))
P
