"""Differential fuzz: oracle/pg_oracle.c (restatement) vs oracle/_ref (unmodified reference).
Usage: python tools/fuzz_oracle_vs_ref.py [n_graphs] [reads_per_graph] [seed] [variant]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import refbind as R
from paragraph_b200 import synth

def main():
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    nr = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    variant = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    R.set_fill_variant(variant)
    rng = np.random.default_rng(seed)
    cells = bad = nreads = 0
    t0 = time.time()
    for gi in range(ng):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 5))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=alpha)
        reads = synth.fuzz_reads(rng, nodes, edges, nr)
        ref_res = R.ref_align_batch(nodes, edges, reads, is_rev=[i & 1 for i in range(len(reads))])
        og = R.OracleGraph(nodes, edges)
        orc_res = og.align_batch(reads, is_rev=[i & 1 for i in range(len(reads))])
        for i, (a, b) in enumerate(zip(ref_res, orc_res)):
            nreads += 1
            if a != b:
                bad += 1
                if bad < 10:
                    print("ALIGN MISMATCH", nodes, edges, reads[i], "\n ref", a, "\n orc", b)
        # raw matrices on the forward graph
        rg = R.RefGssw(nodes, edges)
        for r in reads[: max(3, nr // 5)]:
            ru = r.upper()
            x = rg.fill_trace(ru)
            y = og.fill_trace(ru)
            same = x["cigar"] == y["cigar"] and x["pos"] == y["pos"] and x["score"] == y["score"]
            if variant == 0:
                same = same and (x["stats"] == y["stats"]).all()
            if y["max_node"] >= 0:
                same = same and x["max_node"] == y["max_node"]
            for (h1, e1, f1), (h2, e2, f2) in zip(x["mats"], y["mats"]):
                cells += h1.size
                same = same and (h1 == h2).all()
                if variant == 0:
                    same = same and (e1 == e2).all() and (f1 == f2).all()
            if not same:
                bad += 1
                if bad < 10:
                    print("FILL MISMATCH", nodes, edges, ru, x["stats"].tolist(), y["stats"].tolist(), x["cigar"], y["cigar"], x["max_node"], y["max_node"])
        rg.close(); og.close()
    print(f"graphs={ng} reads={nreads} cells={cells} mismatches={bad} time={time.time()-t0:.1f}s")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
