"""Small run for compute-sanitizer (memcheck / racecheck / synccheck): golden cases + a multi-site fuzz batch."""
import json, glob, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
ctx = capi.Context(0)
n = 0
for p in sorted(glob.glob(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "site_*.json")))[:6]:
    c = json.load(open(p))
    ctx.clear_graphs(); ctx.add_graph(c["nodes"], [tuple(e) for e in c["edges"]])
    got = ctx.align(c["reads"][:40], is_rev=(c["is_rev"] or [0] * len(c["reads"]))[:40], flags=c["flags"])
    for g, e in zip(got, c["expected"][:40]):
        g.pop("status"); g.pop("clipped")
        assert g == e
    n += len(got)
rng = np.random.default_rng(1)
ctx.clear_graphs()
reads, sites = [], []
for k in range(12):
    nodes, edges = synth.bubble_graph(rng, n_nodes=[3, 9, 40][k % 3], max_len=40)
    rd = [r[:160] for r in synth.fuzz_reads(rng, nodes, edges, 6)]
    sid = ctx.add_graph(nodes, edges); reads += rd; sites += [sid] * len(rd)
ctx.align(reads, sites=sites)
print("sanitize run ok:", n + len(reads), "reads")
# reads past the 8-bit limit (WIDE geometries) and the exact-match stage in front of the DP (both path kernels)
nodes, edges = synth.del_graph(rng, 400, 120)
ctx.clear_graphs(); ctx.add_graph(nodes, edges)
ctx.align(synth.simulate_reads(rng, nodes, edges, 24, read_len=300, sub=0.01) + synth.simulate_reads(rng, nodes, edges, 16, read_len=500, sub=0.01))
for second in (False, True):
    ctx.set_stages(32, True, second)
    got = ctx.align(synth.simulate_reads(rng, nodes, edges, 96, read_len=150, sub=0.003, indel_frac=0.0))
    print("cascade:", ctx.path_stats(), sum(g["stage"] == "path" for g in got), "of", len(got), "by the exact-match stage")
ctx.set_stages(32, False, False)
ctx.align(synth.simulate_reads(rng, nodes, edges, 32, read_len=150, sub=0.003, indel_frac=0.0))
ctx.set_stages(0, True, False)
# the k-mer stage (grm::KmerAligner) in front of the DP, alone and behind the exact-match stage
ctx.set_paths(0, [[0, 1, 2], [0, 2]])
ctx.set_kmer_stage(16)
got = ctx.align(synth.simulate_reads(rng, nodes, edges, 96, read_len=150, sub=0.01, indel_frac=0.05))
print("k-mer stage:", ctx.kmer_stats(), sum(g["stage"].startswith("kmer") for g in got), "of", len(got))
ctx.set_stages(32, True, True)
ctx.align(synth.simulate_reads(rng, nodes, edges, 64, read_len=150, sub=0.005, indel_frac=0.02))
ctx.set_stages(0, True, False)
ctx.set_kmer_stage(0)
print("sanitize extras ok")
# round 2, second session: every branch of the lean node events (runs of 1-3 bp nodes, merges with and without the node just
# finished, several sources), many sites in one batch; and the R = 32 geometry (reads of 513 .. 1024 bp)
ctx.clear_graphs()
reads, sites = [], []
for nodes, edges in synth.short_node_graphs(rng, 10):
    sid = ctx.add_graph(nodes, edges)
    rd = [r[:150] for r in synth.fuzz_reads(rng, nodes, edges, 8, max_len=150)]
    reads += rd
    sites += [sid] * len(rd)
ctx.align(reads, sites=sites)
nodes, edges = synth.del_graph(rng, 700, 200)
ctx.clear_graphs(); ctx.add_graph(nodes, edges)
ctx.align(synth.simulate_reads(rng, nodes, edges, 6, read_len=700, sub=0.01) + synth.simulate_reads(rng, nodes, edges, 4, read_len=1024, sub=0.01))
print("sanitize lean events + 1024 bp ok")
