#!/bin/bash
# host mirror as a software pipeline: knob sweep with host phase times; late-round threshold check
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build15.txt 2>&1
nproc
python tools/mirror_sweep.py 30 2>&1 | tee gpurun_out/mirror_sweep15.txt
python tools/kernel_times.py 2>&1 | tee gpurun_out/kt15.txt
timeout 600 python -m pytest tests/test_host_mirror.py tests/test_bench_mirror.py tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/tests15.txt 2>&1; tail -2 gpurun_out/tests15.txt
