#!/bin/bash
# re-entry check of HEAD: GPU suite, bench (both arms), kernel times per shape, A/B: W=16 geometry on the short-node
# shapes, checkpoints every 32 steps
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build20.txt 2>&1
python tools/kernel_times.py > gpurun_out/kt20_base.txt 2>&1; cat gpurun_out/kt20_base.txt
PG_GEOM_W=16 python tools/kernel_times.py > gpurun_out/kt20_w16.txt 2>&1; cat gpurun_out/kt20_w16.txt
PG_LIB=$PWD/ab_build/libpg_ck32.so python tools/kernel_times.py > gpurun_out/kt20_ck32.txt 2>&1; cat gpurun_out/kt20_ck32.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/tests20.txt 2>&1; tail -3 gpurun_out/tests20.txt
python bench.py --impl reference > gpurun_out/bench20_ref.json 2> gpurun_out/bench20_ref.err; cut -c1-300 gpurun_out/bench20_ref.json
python bench.py > gpurun_out/bench20.json 2> gpurun_out/bench20.err; cut -c1-900 gpurun_out/bench20.json
