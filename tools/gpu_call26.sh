#!/bin/bash
# final state of round 2 (lean node events, CK = 32 / TS = 16, 64 GB scratch): ncu captures, launch list, sanitizers, GPU suite,
# both bench arms
set -u
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build26.txt 2>&1; tail -1 gpurun_out/build26.txt
M="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none --import-source on -k regex:'pg_' -s 8 -c 8 -f -o /tmp/ncu/config2_step python tools/profile_run.py config2 > gpurun_out/ncu26a.log 2>&1; tail -1 gpurun_out/ncu26a.log
python tools/ncu_step_summary.py /tmp/ncu/config2_step.ncu-rep gpurun_out/r02g_step_ncu.json "config 2 (10 000 reads), PG_SPLIT=1, one DP step: forward fill, plan, pair, paired reversed-graph fill, plan, pair, second-round fill, traceback (tools/profile_run.py config2; ncu --set full --clock-control none); lean node events, CK = 32, TS = 16"
ncu -i /tmp/ncu/config2_step.ncu-rep --page source --csv --kernel-name regex:pg_fill_kernel --launch-skip 0 --launch-count 1 > /tmp/ncu/fill_src.csv 2>/dev/null; gzip -c /tmp/ncu/fill_src.csv > gpurun_out/r02g_fill_fwd_source.csv.gz
ncu -i /tmp/ncu/config2_step.ncu-rep --page source --csv --kernel-name regex:pg_trace_kernel > /tmp/ncu/trace_src.csv 2>/dev/null; gzip -c /tmp/ncu/trace_src.csv > gpurun_out/r02g_trace_source.csv.gz
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none -k regex:'pg_' -s 8 -c 8 -f -o /tmp/ncu/config4_step python tools/profile_run.py config4_share > gpurun_out/ncu26b.log 2>&1; tail -1 gpurun_out/ncu26b.log
python tools/ncu_step_summary.py /tmp/ncu/config4_step.ncu-rep gpurun_out/r02g_config4_step_ncu.json "config-4 share (1 250 vcf2paragraph-shaped sites, 96k reads), PG_SPLIT=1, one DP step (tools/profile_run.py config4_share); lean node events, CK = 32, TS = 16"
cp gpurun_out/r02g_step_ncu.json gpurun_out/r02g_config4_step_ncu.json profiles/
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02g_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu26d.log 2>&1
python tools/kernel_times.py > gpurun_out/r02g_kernel_times.txt 2>&1; cat gpurun_out/r02g_kernel_times.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests26.txt 2>&1; tail -n 2 gpurun_out/tests26.txt
python bench.py --impl reference > gpurun_out/r02g_bench_reference_arm.json 2> gpurun_out/bench26_ref.err; cut -c1-200 gpurun_out/r02g_bench_reference_arm.json
python bench.py > gpurun_out/r02g_bench_n1.json 2> gpurun_out/bench26.err; cut -c1-1200 gpurun_out/r02g_bench_n1.json
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool python tools/sanitize_check.py > /tmp/ncu/sanitize_$tool.txt 2>&1
  (echo "== $tool"; grep -E "SUMMARY|sanitize|Hazard|hazard|error" /tmp/ncu/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -20) >> gpurun_out/r02g_sanitizer.txt
done
cat gpurun_out/r02g_sanitizer.txt
timeout 300 python tools/gpu_fuzz.py 600 48 93 > gpurun_out/r02g_gpu_fuzz.txt 2>&1; tail -n 1 gpurun_out/r02g_gpu_fuzz.txt
python tools/mirror_sweep.py 30 > gpurun_out/r02g_mirror_sweep.txt 2>&1; tail -n 12 gpurun_out/r02g_mirror_sweep.txt
