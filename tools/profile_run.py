"""Minimal driver for ncu / compute-sanitizer: config 2 through the C-ABI, DP only (2 runs) then the cascade with the
exact-match stage in front (1 run).  Kernel launch order: fill, trace, fill, trace, path, fill, trace.
usage: profile_run.py [n_reads]"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from paragraph_b200 import capi, synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
nodes, edges, reads = synth.config2(seed=42, n_reads=n)
ctx = capi.Context(0)
ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads, pinned=True)
for _ in range(2):
    ctx.align_packed(blob, off)
ctx.set_stages(32, True, True)
rec, ops = ctx.align_packed(blob, off)
print("path stage:", ctx.path_stats(), "records", len(rec), "ops", len(ops))
ctx.close()
