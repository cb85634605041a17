"""Time the exact-match index build (device vs host) on the multi-site shapes.  Run under gpurun."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import numpy as np
from paragraph_b200 import capi, synth
for name, seed, n, kinds in (("config3", 3, 1000, ("DEL", "INS")), ("config4/8", 4, 1250, ("DEL", "INS", "DUP", "INV"))):
    sites = synth.sites(seed=seed, n_sites=n, kinds=kinds, max_reads=4)
    ctx = capi.Context(0)
    reads, sids = [], []
    for (_, nodes, edges, rds) in sites:
        sid = ctx.add_graph(nodes, edges); reads += rds; sids += [sid] * len(rds)
    ctx.align(reads[:8], sites=sids[:8])
    ctx.set_stages(32, True, True)
    t0 = time.perf_counter(); ctx.align(reads[:8], sites=sids[:8]); dt = time.perf_counter() - t0
    print(name, "PG_PATH_HOST_INDEX=%%s" %% os.environ.get("PG_PATH_HOST_INDEX", "0"), "first cascade call %%.1f ms, index build %%.1f ms" %% (dt * 1e3, ctx.path_stats()["index_build_ms"]), flush=True)
    ctx.close()
''' % ROOT
for h in ("0", "1"):
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PG_PATH_HOST_INDEX=h, PG_DEBUG_TIMING="1"))
