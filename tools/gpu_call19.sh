#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build19.txt 2>&1
python tools/mirror_sweep.py 40 2>&1 | tee gpurun_out/mirror_sweep19.txt
timeout 600 python -m pytest tests/test_host_mirror.py tests/test_bench_mirror.py -m gpu -x -q > gpurun_out/tests19.txt 2>&1; tail -2 gpurun_out/tests19.txt
python bench.py > gpurun_out/bench19.json 2> gpurun_out/bench19.err; cut -c1-600 gpurun_out/bench19.json
