#!/bin/bash
# R = 32 geometry (reads up to 1024 bp) on the device, index-build timing, kernel times (nothing else may have moved)
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build27.txt 2>&1; tail -1 gpurun_out/build27.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests27.txt 2>&1; tail -n 3 gpurun_out/tests27.txt
python tools/kernel_times.py 2>&1 | tee gpurun_out/kt27.txt
python tools/index_time.py 2>&1 | tee gpurun_out/r02g_index_time.txt
python - <<'P' 2>&1 | tee gpurun_out/r02g_long_reads.txt
import sys, time; sys.path.insert(0, ".")
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R
rng = np.random.default_rng(3)
nodes, edges = synth.del_graph(rng, 1100, 300)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
for rl in (512, 1024):
    reads = synth.simulate_reads(rng, nodes, edges, 2000, read_len=rl, sub=0.01)
    blob, off = ctx.pack_reads(reads)
    for _ in range(2): ctx.align_packed(blob, off)
    t0 = time.perf_counter(); rec, ops = ctx.align_packed(blob, off); dt = time.perf_counter() - t0
    s = ctx.stats()
    exp = R.ref_align_batch(nodes, edges, reads[:200], threads=8)
    got = ctx.align(reads[:200])
    bad = sum(1 for g, e in zip(got, exp) if {k: g[k] for k in e} != e)
    print("read length %d: 2000 reads in %.2f ms (fill %.2f ms, trace %.2f ms) = %.0f reads/s; 200 reads vs the compiled reference: %d mismatches" % (rl, dt * 1e3, s["fill_ms"], s["trace_ms"], 2000 / dt, bad))
P
