#!/bin/bash
# round 2, call 5: new node-table layout (byte-packed rows, 128-bit stores, registers for the just-finished predecessor)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build5.txt 2>&1; tail -2 gpurun_out/build5.txt
python tools/kernel_times.py > gpurun_out/kt5.txt 2>&1; cat gpurun_out/kt5.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests5.txt 2>&1; tail -6 gpurun_out/tests5.txt
timeout 300 python tools/gpu_fuzz.py 60 300 21 > gpurun_out/fuzz5.txt 2>&1; tail -2 gpurun_out/fuzz5.txt
