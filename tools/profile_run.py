"""Minimal driver for ncu / compute-sanitizer: a workload through the C-ABI, DP only (2 runs), then the cascade with the
exact-match stage in front (1 run) and with the k-mer stage as well (1 run).
Kernel launch order of a DP run with PG_SPLIT=1: fill (forward graph), plan, pair, fill (paired reversed graph), plan, pair,
fill (second round, strided), trace = 8 launches.
usage: profile_run.py [config2 | config3 | config4_share | config5] [n_reads for config2]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth

name = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].isdigit() else "config2"
ctx = capi.Context(0)
if name == "config2":
    n = int(sys.argv[-1]) if sys.argv[-1].isdigit() else 10000
    nodes, edges, reads = synth.config2(seed=42, n_reads=n)
    ctx.add_graph(nodes, edges)
    ctx.set_paths(0, [[0, 1, 2], [0, 2]])
    blob, off = ctx.pack_reads(reads, pinned=True)
    sites = None
else:
    sw = synth.packed_sweep(seed=4, n_sites=1250) if name == "config4_share" else synth.packed_sweep(seed=3, n_sites=1000, kinds=("DEL", "INS"), shaped=False)
    ctx.add_graphs(sw["graphs"])
    for i, (nodes, edges) in enumerate(sw["graphs"]):
        ctx.set_paths(i, synth.haplotype_paths(nodes, edges, limit=8))
    blob, off, sites = sw["blob"], sw["off"], sw["site"]
for _ in range(2):
    ctx.align_packed(blob, off, sites)
ctx.set_stages(32, True, True)
rec, ops = ctx.align_packed(blob, off, sites)
print("path stage:", ctx.path_stats(), "records", len(rec), "ops", len(ops))
ctx.set_kmer_stage(16)
rec, ops = ctx.align_packed(blob, off, sites)
print("path + kmer stage:", ctx.path_stats(), ctx.kmer_stats(), "by stage:", np.bincount(rec["mapped_by"], minlength=6).tolist())
ctx.close()
