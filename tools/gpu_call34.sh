#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python tools/sanitize_check.py > gpurun_out/racecheck34.txt 2>&1
grep -n -B2 -A12 "azard" gpurun_out/racecheck34.txt | head -80
