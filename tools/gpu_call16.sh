#!/bin/bash
# where the host time of upload + run goes in the mirror; index build anomaly
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build16.txt 2>&1
python - <<'PY' > gpurun_out/hostlaps16.txt 2>&1
import os, subprocess, sys, tempfile
sys.path.insert(0, os.getcwd())
from paragraph_b200 import synth
import bench
nodes, edges, reads = bench.workload(0)
with tempfile.TemporaryDirectory() as tmp:
    f2 = os.path.join(tmp, "config2.txt")
    synth.write_workload_file(f2, [("DEL", nodes, edges, reads)])
    for minr in (1 << 30, 4096):
        env = dict(os.environ, PG_DEBUG_HOST="1", PGB_PIPELINE_MIN_READS=str(minr))
        r = subprocess.run(["tools/cpp/bench_mirror", f2, "alignReads", "3", "3", "16", "0"], capture_output=True, text=True, env=env)
        print("=== min_reads", minr)
        print(r.stdout)
        print("\n".join(r.stderr.strip().split("\n")[-80:]))
PY
tail -60 gpurun_out/hostlaps16.txt
python tools/index_time.py 2>&1 | tee gpurun_out/index_time16.txt
