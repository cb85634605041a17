"""Differential fuzz of the R = 32 geometry on the device (reads of 513..1024 bp through the C-ABI) against the compiled
reference when it is there, else the oracle.  usage: gpu_fuzz_long.py [n_graphs] [reads_per_graph] [seed]   (under gpurun)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200
nr = int(sys.argv[2]) if len(sys.argv) > 2 else 8
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
ctx = capi.Context(0)
n = bad = hi = 0
t0 = time.time()
for gi in range(ng):
    alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
    nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 7)), max_len=int(rng.choice([5, 60, 300, 700])), alphabet=alpha)
    reads = [r[:1024] for r in synth.fuzz_reads(rng, nodes, edges, nr, min_len=513, max_len=1024) if len(r) > 0]
    isrev = [i & 1 for i in range(len(reads))]
    flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
    if R.have_ref():
        exp = R.ref_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags, threads=8)
    else:
        exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    got = ctx.align(reads, is_rev=isrev, flags=flags)
    for g, e in zip(got, exp):
        g = dict(g)
        st = g.pop("status", 0); g.pop("clipped", None)
        bad += (g != e or st != 0)
    n += len(reads)
    hi = max([hi] + [e["score"] for e in exp])
print("GPU LONG-READ FUZZ (R = 32 geometry, 513..1024 bp) against %s: %d reads over %d graphs, %d mismatches, top score %d, %.0f s"
      % ("oracle/_ref" if R.have_ref() else "the oracle", n, ng, bad, hi, time.time() - t0))
