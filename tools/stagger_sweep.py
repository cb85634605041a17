"""Chunk pipelines inside a batch: PG_SPLIT (number of chunks) x PG_STAGGER (first chunk in per cent of the others) x PG_PRIO
(priority of the auxiliary stream) on the named workloads; run + download of a resident batch, median of 7, plus a result
digest per combination (all must agree).  usage: stagger_sweep.py [workload ...]   (under gpurun)"""
import hashlib, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from paragraph_b200 import capi, synth

def digest(rec, ops):
    h = hashlib.sha1()
    for f in ("graph_pos", "score", "unique", "chose_reverse", "status", "query_clipped", "cigar_len"):
        h.update(np.ascontiguousarray(rec[f]).tobytes())
    o = np.asarray(ops)
    for x in rec[:: max(1, len(rec) // 4000)]:
        h.update(o[int(x["cigar_off"]):int(x["cigar_off"]) + int(x["cigar_len"])].tobytes())
    return h.hexdigest()[:10]

names = sys.argv[1:] or ["config2", "config3", "config4_share", "config5"]
combos = [(2, 0, 0), (2, 50, 0), (2, 0, 1), (2, 50, 1), (3, 0, 0), (3, 50, 0), (4, 0, 0), (4, 50, 0), (4, 50, 1), (4, 50, 2), (6, 50, 0), (8, 50, 0), (4, 30, 0), (4, 70, 0)]
for name in names:
    sites = synth.workload(name)
    reads, sids, cells = synth.flatten_sites(sites)
    digs = set()
    for split, stag, prio in combos:
        os.environ["PG_SPLIT"], os.environ["PG_STAGGER"], os.environ["PG_PRIO"] = str(split), str(stag), str(prio)
        ctx = capi.Context(0)
        for (_, nodes, edges, _) in sites:
            ctx.add_graph(nodes, edges)
        blob, off = ctx.pack_reads(reads, pinned=True)
        ctx.upload(blob, off, sids)
        for _ in range(2):
            ctx.run()
        torch.cuda.synchronize()
        w = []
        for _ in range(7):
            t0 = time.perf_counter()
            ctx.run()
            rec, ops = ctx.download()
            w.append(time.perf_counter() - t0)
        d = digest(rec, ops)
        digs.add(d)
        print("%-14s split %d stagger %3d prio %d: run+download %.3f ms (%.2f Mreads/s) digest %s" % (name, split, stag, prio, np.median(w) * 1e3, len(reads) / np.median(w) / 1e6, d), flush=True)
        ctx.close()
    print(name, "DIGESTS", "AGREE" if len(digs) == 1 else "DIFFER", flush=True)
