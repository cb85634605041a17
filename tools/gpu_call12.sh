#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build12.txt 2>&1
for tool in synccheck racecheck memcheck; do
  timeout 900 compute-sanitizer --tool $tool python tools/sanitize_check.py > /tmp/sanitize_$tool.txt 2>&1
  (echo "== compute-sanitizer --tool $tool python tools/sanitize_check.py"; grep -E "SUMMARY|sanitize|Race reported|Barrier error" /tmp/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -12) >> gpurun_out/r02e_sanitizer.txt
done
cat gpurun_out/r02e_sanitizer.txt
python tools/kernel_times.py config2 config4_share > gpurun_out/kt12.txt 2>&1; cat gpurun_out/kt12.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests12.txt 2>&1; tail -4 gpurun_out/tests12.txt
( time python bench.py ) > gpurun_out/bench12.json 2> gpurun_out/bench12.err; tail -c 300 gpurun_out/bench12.json; tail -3 gpurun_out/bench12.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench12_ref.json 2> gpurun_out/bench12_ref.err; tail -c 200 gpurun_out/bench12_ref.json
