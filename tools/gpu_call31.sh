#!/bin/bash
# final check of HEAD: smoke, GPU suite, bench at N = 1 (both arms) and N = 2 (tensor gather of the sweep on NCCL), host phases
# of the SitePipeline leg
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke31.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests31.txt 2>&1; tail -n 2 gpurun_out/tests31.txt
python bench.py --impl reference > gpurun_out/r02h_bench_reference_arm.json 2> gpurun_out/bench31_ref.err; cut -c1-200 gpurun_out/r02h_bench_reference_arm.json
python bench.py > gpurun_out/r02h_bench_n1.json 2> gpurun_out/bench31.err; cut -c1-700 gpurun_out/r02h_bench_n1.json
python - <<'P' 2>&1 | tee gpurun_out/r02h_pipeline_phases.txt
import os, subprocess, sys, tempfile
sys.path.insert(0, ".")
from paragraph_b200 import synth
sw = synth.packed_sweep(seed=4, n_sites=1250)
with tempfile.TemporaryDirectory() as tmp:
    f4 = os.path.join(tmp, "config4_share.txt")
    synth.write_workload_file(f4, synth.sweep_as_site_list(sw))
    for th in (1, 16):
        r = subprocess.run(["tools/cpp/bench_mirror", f4, "pipeline", "5", "2", str(th), "0"], capture_output=True, text=True)
        print("threads", th, r.stdout.strip(), "|", r.stderr.strip().split("\n")[-1])
P
