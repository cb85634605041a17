#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build17.txt 2>&1
python tools/mirror_sweep.py 30 2>&1 | tee gpurun_out/mirror_sweep17.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests17.txt 2>&1; tail -2 gpurun_out/tests17.txt
