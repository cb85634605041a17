import os, sys
sys.path.insert(0, "/root/repo")
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads, pinned=True)
ctx.set_stages(32, True, True)
ts = []
for _ in range(8):
    ctx.align_packed(blob, off); ts.append(ctx.path_stats()["path_ms"])
print("PG_PATH_HOST_INDEX=%s path_ms min %.4f median %.4f" % (os.environ.get("PG_PATH_HOST_INDEX", "0"), min(ts), sorted(ts)[4]))
