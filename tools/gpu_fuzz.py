"""Large differential fuzz of the CUDA path (through the C-ABI) against the oracle, run under gpurun.
usage: gpu_fuzz.py [n_graphs] [reads_per_graph] [seed]"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

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
        flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
        maxlen = int(rng.choice([160, 160, 250, 320, 512]))
        for gi in range(min(200, ng - batch)):
            alpha = ["ACGT", "ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 5))]
            nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 10)), max_len=int(rng.choice([5, 20, 60, 200, 600])), alphabet=alpha)
            minlen = 8 if maxlen <= 250 else int(rng.choice([8, 200]))
            rd = [r[:maxlen] for r in synth.fuzz_reads(rng, nodes, edges, nr, min_len=minlen, max_len=maxlen)]
            rv = [int(x) for x in rng.integers(0, 2, size=len(rd))]
            sid = ctx.add_graph(nodes, edges)
            reads += rd; sites += [sid] * len(rd); isrev += rv
            exp += R.OracleGraph(nodes, edges).align_batch(rd, is_rev=rv, flags=flags)
        got = ctx.align(reads, sites=sites, is_rev=isrev, flags=flags)
        for i, (g, e) in enumerate(zip(got, exp)):
            st = g.pop("status"); cl = g.pop("clipped")
            ok = (g == e and st == 0)
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
