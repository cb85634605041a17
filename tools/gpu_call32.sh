#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build32.txt 2>&1
python tools/mirror_split_sweep.py 30 2>&1 | tee gpurun_out/mirror_split32.txt
