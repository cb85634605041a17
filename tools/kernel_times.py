"""Fill / traceback kernel-phase times of the named workloads (synth.workload) with the kernels run back to back on
one stream (PG_SPLIT=1: the default two-stream half-batch overlap stretches per-phase event times), plus the default
run's whole-batch time.  usage: kernel_times.py [workload ...]   (under gpurun; honours PG_LIB)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from paragraph_b200 import capi, synth

def measure(name):
    sites = synth.workload(name)
    reads, sids, cells = synth.flatten_sites(sites)
    out = {}
    for split in ("1", None):
        if split:
            os.environ["PG_SPLIT"] = split
        ctx = capi.Context(0)
        os.environ.pop("PG_SPLIT", None)
        for (_, nodes, edges, _) in sites:
            ctx.add_graph(nodes, edges)
        blob, off = ctx.pack_reads(reads, pinned=True)
        ctx.upload(blob, off, sids)
        for _ in range(2):
            ctx.run()
        torch.cuda.synchronize()
        f, t, w = [], [], []
        for _ in range(5):
            t0 = time.perf_counter()
            ctx.run()
            ctx.download()
            w.append(time.perf_counter() - t0)
            s = ctx.stats()
            f.append(s["fill_ms"]); t.append(s["trace_ms"])
        out[split or "default"] = (float(np.median(f)), float(np.median(t)), float(np.median(w)) * 1e3)
        ctx.close()
    f, t, w1 = out["1"]
    _, _, w = out["default"]
    print("%-14s sites=%d reads=%d  split=1: fill %.3f ms (%.2f Tcell/s) trace %.3f ms, run+download %.2f ms | default: run+download %.2f ms (%.2f Mreads/s)"
          % (name, len(sites), len(reads), f, cells / f / 1e9, t, w1, w, len(reads) / w / 1e3), flush=True)

if __name__ == "__main__":
    for nm in (sys.argv[1:] or ["config2", "config3", "config4_share", "config5"]):
        measure(nm)
