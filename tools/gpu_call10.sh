#!/bin/bash
# round 2, call 10: sanitizer logs (truncated), ncu captures exported to text on the box (the reports themselves are too
# big to bring back together), timeline, launch list
set -u
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build10.txt 2>&1; tail -1 gpurun_out/build10.txt
for tool in synccheck racecheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_check.py > /tmp/ncu/sanitize_$tool.txt 2>&1
  head -150 /tmp/ncu/sanitize_$tool.txt > gpurun_out/r02d_sanitize_$tool.txt; echo "..." >> gpurun_out/r02d_sanitize_$tool.txt; grep -E "SUMMARY|sanitize|Hazard|hazard|error" /tmp/ncu/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -40 >> gpurun_out/r02d_sanitize_$tool.txt
  tail -3 /tmp/ncu/sanitize_$tool.txt
done
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_check.py > /tmp/ncu/sanitize_memcheck.txt 2>&1; grep -E "SUMMARY|sanitize" /tmp/ncu/sanitize_memcheck.txt > gpurun_out/r02d_sanitize_memcheck.txt; cat gpurun_out/r02d_sanitize_memcheck.txt
M="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none --import-source on -k regex:'pg_' -s 8 -c 8 -f -o /tmp/ncu/config2_step python tools/profile_run.py config2 > gpurun_out/ncu10a.log 2>&1; tail -1 gpurun_out/ncu10a.log
python tools/ncu_step_summary.py /tmp/ncu/config2_step.ncu-rep gpurun_out/r02d_step_ncu.json "config 2 (10 000 reads), PG_SPLIT=1, one DP step: forward fill, plan, pair, paired reversed-graph fill, plan, pair, second-round fill, traceback (tools/profile_run.py config2; ncu --set full --clock-control none)"
ncu -i /tmp/ncu/config2_step.ncu-rep --page raw --csv > gpurun_out/r02d_config2_step_raw.csv 2>/dev/null
ncu -i /tmp/ncu/config2_step.ncu-rep --page source --csv --kernel-name regex:pg_fill_kernel --launch-skip 0 --launch-count 1 > /tmp/ncu/fill_src.csv 2>/dev/null; gzip -c /tmp/ncu/fill_src.csv > gpurun_out/r02d_fill_fwd_source.csv.gz
ncu -i /tmp/ncu/config2_step.ncu-rep --page source --csv --kernel-name regex:pg_trace_kernel > /tmp/ncu/trace_src.csv 2>/dev/null; gzip -c /tmp/ncu/trace_src.csv > gpurun_out/r02d_trace_source.csv.gz
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none -k regex:'pg_' -s 8 -c 8 -f -o /tmp/ncu/config4_step python tools/profile_run.py config4_share > gpurun_out/ncu10b.log 2>&1; tail -1 gpurun_out/ncu10b.log
python tools/ncu_step_summary.py /tmp/ncu/config4_step.ncu-rep gpurun_out/r02d_config4_step_ncu.json "config-4 share (1 250 vcf2paragraph-shaped sites, 94k reads), PG_SPLIT=1, one DP step (tools/profile_run.py config4_share)"
ncu -i /tmp/ncu/config4_step.ncu-rep --page raw --csv > gpurun_out/r02d_config4_step_raw.csv 2>/dev/null
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none -k regex:'pg_path_warp_kernel|pg_kmer_kernel|pg_path_index' -c 6 -f -o /tmp/ncu/front python tools/profile_run.py config2 > gpurun_out/ncu10c.log 2>&1; tail -1 gpurun_out/ncu10c.log
python tools/ncu_step_summary.py /tmp/ncu/front.ncu-rep gpurun_out/r02d_front_stages_ncu.json "exact-match and k-mer stage kernels on config 2 (tools/profile_run.py config2)"
ncu -i /tmp/ncu/front.ncu-rep --page raw --csv > gpurun_out/r02d_front_stages_raw.csv 2>/dev/null
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu10d.log 2>&1
PG_DEBUG_TIMELINE=1 python -c "
import sys; sys.path.insert(0,'.')
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges); blob, off = ctx.pack_reads(reads, pinned=True)
for _ in range(4): ctx.align_packed(blob, off)
" 2> gpurun_out/r02d_timeline.txt; tail -2 gpurun_out/r02d_timeline.txt
python tools/kernel_times.py > gpurun_out/kt10.txt 2>&1; cat gpurun_out/kt10.txt
timeout 200 python tools/gpu_fuzz.py 60 150 31 > gpurun_out/fuzz10.txt 2>&1; tail -1 gpurun_out/fuzz10.txt
du -sh gpurun_out
