"""Differential fuzz of the fill's lean node events on the device: graphs with runs of 1-3 bp nodes, chain links, merges with and
without the node just finished, several sources (synth.short_node_graphs), many sites per batch, against the compiled reference
when it is there, else the oracle.  usage: gpu_fuzz_short_nodes.py [n_batches] [graphs_per_batch] [seed]   (under gpurun)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

nb = int(sys.argv[1]) if len(sys.argv) > 1 else 50
gpb = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rng = np.random.default_rng(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
ctx = capi.Context(0)
n = bad = 0
t0 = time.time()
for b in range(nb):
    ctx.clear_graphs()
    reads, sites, exp = [], [], []
    flags = int(rng.choice([0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 3, 7]))
    for nodes, edges in synth.short_node_graphs(rng, gpb):
        sid = ctx.add_graph(nodes, edges)
        rd = [r[:150] for r in synth.fuzz_reads(rng, nodes, edges, 16, max_len=150) if len(r) > 0]
        reads += rd
        sites += [sid] * len(rd)
        exp += R.ref_align_batch(nodes, edges, rd, flags=flags, threads=8) if R.have_ref() else R.OracleGraph(nodes, edges).align_batch(rd, flags=flags)
    got = ctx.align(reads, sites=sites, flags=flags)
    for g, e in zip(got, exp):
        g = dict(g)
        st = g.pop("status", 0); g.pop("clipped", None)
        bad += (g != e or st != 0)
    n += len(reads)
print("GPU SHORT-NODE FUZZ (lean node events) against %s: %d reads over %d graphs in %d batches, %d mismatches, %.0f s"
      % ("oracle/_ref" if R.have_ref() else "the oracle", n, nb * gpb, nb, bad, time.time() - t0))
