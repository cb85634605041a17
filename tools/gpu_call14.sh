#!/bin/bash
# late second-round reversed-graph fills (PG_LATE_ROUND) + speculative op-word download: A/B and parity
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build14.txt 2>&1
for v in 0 1 0 1; do
  echo "PG_LATE_ROUND=$v"; PG_LATE_ROUND=$v python tools/kernel_times.py 2>&1 | tee -a gpurun_out/kt14_$v.txt
done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/tests14.txt 2>&1; tail -2 gpurun_out/tests14.txt
timeout 200 python tools/gpu_fuzz.py 60 150 51 > gpurun_out/fuzz14.txt 2>&1; tail -1 gpurun_out/fuzz14.txt
for v in 0 1; do
  PG_LATE_ROUND=$v python bench.py --no-cpu-baseline --no-extra-legs > gpurun_out/bench14_$v.json 2> gpurun_out/bench14_$v.err; cat gpurun_out/bench14_$v.json
done
