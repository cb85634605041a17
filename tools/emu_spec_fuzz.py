"""Differential fuzz of the speculative "no gap alive" blocks (pg_core.cuh: lane_step_dead) in the CPU lane emulator against the
compiled reference (oracle/_ref): config-2 batches, DEL/INS/DUP/INV sites, vcf2paragraph-shaped long deletions, random
bubble graphs over several alphabets; all geometries and flag sets.  usage: emu_spec_fuzz.py <seed> <seconds> [mode: 1 = speculative blocks (default), 2 = plus the pruning experiment]"""
import sys, time
import os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import emubind
from paragraph_b200 import synth
from oracle import refbind as R
from conftest import strip_status
emubind.set_spec(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
R.set_fill_variant(0)
rng = np.random.default_rng(int(sys.argv[1]))
t0 = time.time(); n = 0; bad = 0; it = 0
while time.time() - t0 < float(sys.argv[2]):
    it += 1
    kind = it % 4
    if kind == 0:
        nodes, edges, reads = synth.config2(seed=int(rng.integers(1, 1 << 30)), n_reads=12)
        isrev = [int(rng.integers(0, 2)) for _ in reads]
    elif kind == 1:
        s = synth.sites(int(rng.integers(1, 1 << 30)), 1, kinds=(["DEL", "INS", "DUP", "INV"][int(rng.integers(0, 4))],), max_reads=12)[0]
        nodes, edges, reads = s[1], s[2], s[3]
        isrev = [i & 1 for i in range(len(reads))]
    elif kind == 2:
        nodes, edges = synth.long_del_graph(rng)
        reads = synth.simulate_reads(rng, nodes, edges, 12, indel_frac=0.3)
        isrev = [i & 1 for i in range(len(reads))]
    else:
        alpha = ["ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200, 600])), alphabet=alpha)
        reads = [r[:250] for r in synth.fuzz_reads(rng, nodes, edges, 12, max_len=int(rng.choice([60, 160, 250])))]
        isrev = [i & 1 for i in range(len(reads))]
    flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 1, 3, 5, 7]))
    w = int(rng.choice([32, 32, 16, 8]))
    emubind.set_geometry(w)
    exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
    got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
    if strip_status(got) != exp:
        bad += 1
        print("MISMATCH iteration", it, "kind", kind, "w", w, "flags", flags, flush=True)
    n += len(reads)
db = emubind.dead_boundary_stats()
print("boundary sub-blocks run dead %d, attempted and redone %d" % tuple(db), flush=True)
st = emubind.spec_stats(); tot = sum(st)
print("SPEC FUZZ seed %s: %d reads in %d batches, %d mismatching batches; blocks dead %.1f%% redone %.1f%% alive %.1f%% boundary %.1f%%" % ((sys.argv[1], n, it, bad) + tuple(100.0 * x / tot for x in st)), flush=True)
