"""Large differential fuzz of the CUDA path (through the C-ABI) against the oracle, run under gpurun.
usage: gpu_fuzz.py [n_graphs] [reads_per_graph] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

estage = []


def cascade(graphs, reads, sites, isrev, flags, gexp, k, second):
    """CompositeAligner(path, gssw) from the two oracles (CompositeAligner.cpp:78-176), per site"""
    global estage
    out, estage = list(gexp), ["gssw"] * len(gexp)
    by_site = {}
    for i, s in enumerate(sites):
        by_site.setdefault(s, []).append(i)
    for s, idx in by_site.items():
        nodes, edges = graphs[s]
        pexp, _ = R.OraclePathIndex(nodes, edges, k).align_batch([reads[i] for i in idx])
        og = None
        for i, p in zip(idx, pexp):
            if not p["mapped"]:
                continue
            if second and not p["unique"]:
                og = og or R.OracleGraph(nodes, edges)
                out[i] = og.align_batch([p["bases"]], is_rev=[isrev[i]], flags=flags)[0]
                estage[i] = "gssw2" if p["graph_reverse"] else "gssw"
            else:
                out[i] = {key: p[key] for key in ("pos", "score", "unique", "mapq", "graph_reverse", "bases", "cigar")}
                estage[i] = "path"
    return out


def main():
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    nr = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 99
    rng = np.random.default_rng(seed)
    ctx = capi.Context(0)
    bad = tot = 0
    t0 = time.time()
    for batch in range(0, ng, 200):
        ctx.clear_graphs()
        reads, sites, exp, isrev = [], [], [], []
        ctx_graphs = {}
        flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
        maxlen = int(rng.choice([160, 160, 250, 320, 512]))
        for gi in range(min(200, ng - batch)):
            alpha = ["ACGT", "ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 5))]
            nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 10)), max_len=int(rng.choice([5, 20, 60, 200, 600])), alphabet=alpha)
            minlen = 8 if maxlen <= 250 else int(rng.choice([8, 200]))
            rd = [r[:maxlen] for r in synth.fuzz_reads(rng, nodes, edges, nr, min_len=minlen, max_len=maxlen)]
            rv = [int(x) for x in rng.integers(0, 2, size=len(rd))]
            sid = ctx.add_graph(nodes, edges)
            ctx_graphs[sid] = (nodes, edges)
            reads += rd; sites += [sid] * len(rd); isrev += rv
            exp += R.OracleGraph(nodes, edges).align_batch(rd, is_rev=rv, flags=flags)
        # every other batch: the exact-match stage (grm::PathAligner) in front, with / without the filter's second chance
        k = int(rng.choice([0, 0, 8, 16, 32]))
        second = bool(rng.integers(0, 2))
        ctx.set_stages(k, True, second)
        if k:
            exp = cascade(ctx_graphs, reads, sites, isrev, flags, exp, k, second)
        got = ctx.align(reads, sites=sites, is_rev=isrev, flags=flags)
        for i, (g, e) in enumerate(zip(got, exp)):
            st = g.pop("status"); cl = g.pop("clipped"); stage = g.pop("stage", None)
            ok = (g == e and st == 0)
            if ok and stage is not None:
                ok = stage == estage[i]
            if ok and (flags & 1) and e["cigar"]:
                ok = R.oracle_bad_align(e["cigar"], 0.8)[1] == cl
            if not ok:
                bad += 1
                if bad <= 5:
                    print("MISMATCH flags=%x" % flags, reads[i], "\n got", g, st, cl, "\n exp", e)
        tot += len(exp)
        print("batch %d: reads %d bad %d (%.0fs)" % (batch // 200, tot, bad, time.time() - t0), flush=True)
    print("GPU FUZZ TOTAL reads=%d mismatches=%d" % (tot, bad))
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
