#!/bin/bash
# round 2, call 3: the whole GPU suite on the default (speculative) build incl. the full-size config tests, then bench.py
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build3.txt 2>&1; tail -2 gpurun_out/build3.txt
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/tests3.txt 2>&1; tail -6 gpurun_out/tests3.txt
( time python bench.py ) > gpurun_out/bench3.json 2> gpurun_out/bench3.err; tail -c 3000 gpurun_out/bench3.json; tail -5 gpurun_out/bench3.err
( time python bench.py --impl reference --steps 5 --warmup 1 ) > gpurun_out/bench3_ref.json 2> gpurun_out/bench3_ref.err; tail -c 1500 gpurun_out/bench3_ref.json; tail -3 gpurun_out/bench3_ref.err
nproc
