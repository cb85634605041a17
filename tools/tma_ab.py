import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
code = r'''
import os, sys
sys.path.insert(0, %r)
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads)
for _ in range(3): ctx.align_packed(blob, off)
ts = []
for _ in range(6):
    ctx.align_packed(blob, off); s = ctx.stats(); ts.append((s["fill_ms"], s["trace_ms"]))
f = min(t[0] for t in ts); t = min(t[1] for t in ts)
print("NO_TMA=%%s W=%%s fill %%.3f trace %%.3f total %%.3f ms" %% (os.environ.get("PG_NO_TMA","0"), os.environ.get("PG_GEOM_W","32"), f, t, f+t), flush=True)
''' % ROOT
for notma in ("0", "1"):
    for w in ("32", "16"):
        subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PG_NO_TMA=notma, PG_GEOM_W=w))
