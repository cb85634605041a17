#!/bin/bash
# round 2, call 9: evidence for the final kernels: ncu --set full of one DP step on config 2 and on config-4-shaped sites
# (+ ALU-pipe instruction counts), the path / k-mer kernels, the launch list of bench.py, compute-sanitizer, the
# two-stream timeline, then bench.py
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build9.txt 2>&1; tail -1 gpurun_out/build9.txt
M="--metrics sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_lsu.sum"
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none --import-source on -k regex:'pg_' -s 8 -c 8 -f -o gpurun_out/r02d_config2_step python tools/profile_run.py config2 > gpurun_out/ncu9a.log 2>&1; tail -2 gpurun_out/ncu9a.log
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none -k regex:'pg_' -s 8 -c 8 -f -o gpurun_out/r02d_config4_step python tools/profile_run.py config4_share > gpurun_out/ncu9b.log 2>&1; tail -2 gpurun_out/ncu9b.log
PG_SPLIT=1 timeout 900 ncu --set full $M --clock-control none -k regex:'pg_path_warp_kernel|pg_kmer_kernel|pg_path_index' -c 6 -f -o gpurun_out/r02d_front_stages python tools/profile_run.py config2 > gpurun_out/ncu9c.log 2>&1; tail -2 gpurun_out/ncu9c.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02d_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-legs > gpurun_out/ncu9d.log 2>&1; tail -c 300 gpurun_out/ncu9d.log
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_check.py > gpurun_out/sanitize_$tool.txt 2>&1; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|sanitize" gpurun_out/sanitize_$tool.txt | tail -4
done
PG_DEBUG_TIMELINE=1 python -c "
import sys; sys.path.insert(0,'.')
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges); blob, off = ctx.pack_reads(reads, pinned=True)
for _ in range(4): ctx.align_packed(blob, off)
" 2> gpurun_out/r02d_timeline.txt; tail -4 gpurun_out/r02d_timeline.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or idempotent or imported" > gpurun_out/tests9.txt 2>&1; tail -3 gpurun_out/tests9.txt
( time python bench.py ) > gpurun_out/bench9.json 2> gpurun_out/bench9.err; tail -c 400 gpurun_out/bench9.json; tail -3 gpurun_out/bench9.err
ls -la gpurun_out/*.ncu-rep
