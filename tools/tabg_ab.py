import os, sys
sys.path.insert(0, "/root/repo")
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads, pinned=True)
ts = []
for _ in range(8):
    ctx.align_packed(blob, off); s = ctx.stats(); ts.append((s["fill_ms"], s["trace_ms"]))
print("PG_FORCE_TABG=%s PG_NO_TMA=%s fill %.3f trace %.3f" % (os.environ.get("PG_FORCE_TABG"), os.environ.get("PG_NO_TMA"), min(t[0] for t in ts), min(t[1] for t in ts)))
