"""Throughput + sampled parity on BASELINE.json configs[2..4] shapes (run under gpurun).
  config 3: 1k mixed DEL/INS sites <= 500 bp, synthetic 30x 150 bp, one multi-site batch on 1 GPU
  config 4 (per-GPU share): DEL/INS/DUP/INV sites
  config 5: INV/DUP graphs with 1-10 kb variant nodes, 1k reads/site"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from paragraph_b200 import capi, synth
from oracle import refbind as R

def run(ctx, name, sites, check_sites=6):
    ctx.clear_graphs()
    reads, sids, spans, cells = [], [], [], 0
    for (_, nodes, edges, rds) in sites:
        sid = ctx.add_graph(nodes, edges)
        spans.append((len(reads), len(reads) + len(rds)))
        reads += rds
        sids += [sid] * len(rds)
        cells += 4 * sum(len(r) for r in rds) * sum(len(n) for n in nodes)
    blob, off = ctx.pack_reads(reads)
    st = np.ascontiguousarray(sids, dtype=np.int32)
    ctx.align_packed(blob, off, st)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(3):
        rec, ops = ctx.align_packed(blob, off, st)
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 3
    s = ctx.stats()
    kms = s["fill_ms"] + s["trace_ms"]
    # sampled parity
    bad = 0; nchk = 0
    idx = np.linspace(0, len(sites) - 1, check_sites).astype(int)
    for i in idx:
        _, nodes, edges, rds = sites[i]
        a, b = spans[i]
        exp = R.OracleGraph(nodes, edges).align_batch(rds[:64])
        for j, e in enumerate(exp):
            x = rec[a + j]
            got = (int(x["graph_pos"]), int(x["score"]), bool(x["unique"]), capi.format_cigar(x, ops))
            if got != (e["pos"], e["score"], e["unique"], e["cigar"]) or x["status"]:
                bad += 1
            nchk += 1
    print("%-8s sites=%d reads=%d  e2e %.1f ms (%.2f Mreads/s)  kernels %.2f ms (%.2f Mreads/s, %.2f Tcell/s)  parity %d/%d bad"
          % (name, len(sites), len(reads), dt * 1e3, len(reads) / dt / 1e6, kms, len(reads) / kms / 1e3, cells / kms / 1e9, bad, nchk), flush=True)
    # the same batch through the cascade (exact-match stage in front, k = 32; second chance as the default filters cause)
    ctx.set_stages(32, True, True)
    t0 = time.perf_counter()
    ctx.align_packed(blob, off, st)  # first call: builds the k-mer index of every site (on the device)
    torch.cuda.synchronize(); first = time.perf_counter() - t0
    t0 = time.perf_counter()
    for _ in range(3):
        rec2, ops2 = ctx.align_packed(blob, off, st)
    torch.cuda.synchronize(); dt2 = (time.perf_counter() - t0) / 3
    ps, s2 = ctx.path_stats(), ctx.stats()
    ctx.set_stages(0, True, False)
    print("%-8s   cascade: e2e %.1f ms (%.2f Mreads/s), %d of %d reads by the exact-match stage (%.3f ms), fill+trace %.2f ms; "
          "first call %.1f ms of which index build %.1f ms (%d sites)"
          % ("", dt2 * 1e3, len(reads) / dt2 / 1e6, ps["mapped"], len(reads), ps["path_ms"], s2["fill_ms"] + s2["trace_ms"],
             first * 1e3, ps["index_build_ms"], len(sites)), flush=True)
    return bad

def main():
    ctx = capi.Context(0)
    bad = 0
    bad += run(ctx, "config3", synth.sites(seed=3, n_sites=1000, kinds=("DEL", "INS")))
    bad += run(ctx, "config4/8", synth.sites(seed=4, n_sites=1250, kinds=("DEL", "INS", "DUP", "INV")))
    rng = np.random.default_rng(5)
    big = []
    for i in range(24):
        kind = "INV" if i % 2 else "DUP"
        n = int(rng.integers(1000, 10001))
        nodes, edges = (synth.inv_graph(rng, 500, n) if kind == "INV" else synth.dup_graph(rng, n + 500, n))
        big.append((kind, nodes, edges, synth.simulate_reads(rng, nodes, edges, 1000, alternate=False)))
    bad += run(ctx, "config5", big, check_sites=3)
    print("TOTAL BAD", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
