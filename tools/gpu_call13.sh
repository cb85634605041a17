#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build13.txt 2>&1
python tools/kernel_times.py > gpurun_out/kt13.txt 2>&1; cat gpurun_out/kt13.txt
timeout 600 python -m pytest tests/test_gpu_configs.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tests13.txt 2>&1; tail -2 gpurun_out/tests13.txt
timeout 200 python tools/gpu_fuzz.py 60 150 41 > gpurun_out/fuzz13.txt 2>&1; tail -1 gpurun_out/fuzz13.txt
