#!/bin/bash
# compute-sanitizer on HEAD (lean node events on short-node graphs, R = 32 geometry included)
set -u
mkdir -p gpurun_out /tmp/ncu
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build33.txt 2>&1
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_check.py > /tmp/ncu/sanitize_$tool.txt 2>&1
  (echo "== $tool"; grep -E "SUMMARY|sanitize|Hazard|hazard|error" /tmp/ncu/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -20) >> gpurun_out/r02h_sanitizer.txt
done
cat gpurun_out/r02h_sanitizer.txt
