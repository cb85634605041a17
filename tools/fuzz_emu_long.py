"""Differential fuzz of the R = 32 geometry (reads of 513..1024 bp) in the CPU lane emulator against the compiled reference\n(oracle/_ref).  usage: fuzz_emu_long.py <seed> <seconds>"""
import sys, os, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from oracle import refbind as R
from paragraph_b200 import synth
import emubind
rng = np.random.default_rng(int(sys.argv[1])); T = float(sys.argv[2])
bad = n = 0; t0 = time.time(); hi = 0
while time.time() - t0 < T:
    alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
    nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 7)), max_len=int(rng.choice([5, 60, 300, 700])), alphabet=alpha)
    reads = [r[:1024] for r in synth.fuzz_reads(rng, nodes, edges, 4, min_len=513, max_len=1024) if len(r) > 0]
    isrev = [i & 1 for i in range(len(reads))]
    flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
    exp = R.ref_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
    got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
    for g, e in zip(got, exp):
        g = dict(g); g.pop("status", None); g.pop("clipped", None)
        if g != e: bad += 1
    n += len(reads); hi = max([hi] + [e["score"] for e in exp])
print("LONG-READ EMULATOR FUZZ (R = 32 geometry, reads of 513..1024 bp) vs the compiled reference, seed %s: %d reads, %d mismatches, top score %d" % (sys.argv[1], n, bad, hi))
