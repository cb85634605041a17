"""Differential fuzz: CPU lane emulator of the CUDA path (tests/emu) vs the reference (oracle/_ref) or,
when the reference is absent, the oracle restatement.  Usage: fuzz_emu.py [n_graphs] [reads] [seed] [maxlen] [W]"""
import sys, os, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import refbind as R
from paragraph_b200 import synth
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import emubind

def main():
    ng = int(sys.argv[1]) if len(sys.argv) > 1 else 100
    nr = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    seed = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    maxlen = int(sys.argv[4]) if len(sys.argv) > 4 else 160
    emubind.set_geometry(int(sys.argv[5]) if len(sys.argv) > 5 else 32)
    rng = np.random.default_rng(seed)
    bad = n = tiles = 0
    t0 = time.time()
    for gi in range(ng):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 5))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=alpha)
        reads = [r[:maxlen] for r in synth.fuzz_reads(rng, nodes, edges, nr, max_len=maxlen)]
        isrev = [i & 1 for i in range(len(reads))]
        flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
        if R.have_ref():
            exp = R.ref_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
        else:
            exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
        got, nt = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
        tiles += nt
        for i, (a, b) in enumerate(zip(exp, got)):
            n += 1
            st = b.pop("status"); b.pop("clipped", None)
            if a != b or st != 0:
                bad += 1
                if bad <= 8:
                    print("MISMATCH flags=%x status=%d" % (flags, st), nodes, edges, reads[i], "\n exp", a, "\n got", b)
    print(f"graphs={ng} reads={n} mismatches={bad} tiles/read={tiles/max(n,1):.1f} time={time.time()-t0:.1f}s")
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
