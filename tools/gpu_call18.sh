#!/bin/bash
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build18.txt 2>&1
python tools/mirror_sweep.py 40 2>&1 | tee gpurun_out/mirror_sweep18.txt
