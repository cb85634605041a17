#!/bin/bash
# group barrier after the traceback's parallel op logging: racecheck again, GPU suite, kernel times
set -u
mkdir -p gpurun_out
rm -f gpurun_out/r02h_sanitizer.txt
for tool in memcheck racecheck synccheck; do
  timeout 600 compute-sanitizer --tool $tool python tools/sanitize_check.py > /tmp/sanitize_$tool.txt 2>&1
  (echo "== $tool"; grep -E "SUMMARY|sanitize|Hazard|hazard|error" /tmp/sanitize_$tool.txt | sort | uniq -c | sort -rn | head -20) >> gpurun_out/r02h_sanitizer.txt
done
cat gpurun_out/r02h_sanitizer.txt
python tools/kernel_times.py 2>&1 | tee gpurun_out/r02h_kernel_times.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests35.txt 2>&1; tail -n 2 gpurun_out/tests35.txt
timeout 300 python tools/gpu_fuzz.py 400 48 95 > gpurun_out/r02h_gpu_fuzz.txt 2>&1; tail -n 1 gpurun_out/r02h_gpu_fuzz.txt
