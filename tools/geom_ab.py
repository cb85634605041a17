"""A/B of the lane geometries (W lanes per task) on config 2: parity vs oracle + kernel times (run under gpurun)."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
exp = R.OracleGraph(nodes, edges).align_batch(reads[:400])
rng = np.random.default_rng(5)
fz = []
for gi in range(60):
    n2, e2 = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=["ACGT", "ACGTN"][gi & 1])
    rd = [r[:250] for r in synth.fuzz_reads(rng, n2, e2, 16, max_len=[160, 250][gi % 2])]
    fz.append((n2, e2, rd, R.OracleGraph(n2, e2).align_batch(rd)))
for W in (32, 16, 8):
    os.environ["PG_GEOM_W"] = str(W)
    ctx = capi.Context(0)
    bad = 0
    for (n2, e2, rd, ex) in fz:
        ctx.clear_graphs(); ctx.add_graph(n2, e2)
        got = ctx.align(rd)
        for g, e in zip(got, ex):
            st = g.pop("status"); g.pop("clipped", None)
            bad += (g != e or st != 0)
    ctx.clear_graphs(); ctx.add_graph(nodes, edges)
    got = ctx.align(reads[:400])
    for g, e in zip(got, exp):
        st = g.pop("status"); g.pop("clipped", None)
        bad += (g != e or st != 0)
    blob, off = ctx.pack_reads(reads)
    for _ in range(3):
        ctx.align_packed(blob, off)
    ts = []
    for _ in range(5):
        ctx.align_packed(blob, off)
        s = ctx.stats()
        ts.append((s["fill_ms"], s["trace_ms"]))
    f = min(t[0] for t in ts); t = min(t[1] for t in ts)
    print("W=%2d  mismatches=%d  fill %.3f ms  trace %.3f ms  total %.3f ms -> %.2f Mreads/s" % (W, bad, f, t, f + t, 10 / (f + t)), flush=True)
    ctx.close()
