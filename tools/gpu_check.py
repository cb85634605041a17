"""First-contact GPU check (run under gpurun): parity of the CUDA path vs the oracle on fuzz batches and
a config-2 sample, then a rough timing of the 10k-read config-2 batch."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from paragraph_b200 import capi, synth
from oracle import refbind as R

def compare(got, exp, tag):
    bad = 0
    for i, (g, e) in enumerate(zip(got, exp)):
        st = g.pop("status"); g.pop("clipped", None)
        if g != e or st:
            bad += 1
            if bad <= 5:
                print(tag, "MISMATCH", i, "status", st, "\n got", g, "\n exp", e)
    print(tag, "reads", len(exp), "mismatches", bad, flush=True)
    return bad

def main():
    quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
    ctx = capi.Context(0)
    print(capi.load().pg_version().decode())
    bad = 0
    rng = np.random.default_rng(5)
    # fuzz: many small graphs in ONE multi-site batch
    allreads, allsites, exp = [], [], []
    for gi in range(40 if quick else 300):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 5))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=alpha)
        reads = [r[:160] for r in synth.fuzz_reads(rng, nodes, edges, 20)]
        sid = ctx.add_graph(nodes, edges)
        allreads += reads
        allsites += [sid] * len(reads)
        exp += R.OracleGraph(nodes, edges).align_batch(reads)
    got = ctx.align(allreads, sites=allsites)
    bad += compare(got, exp, "fuzz-multisite")
    # long reads (R=8 instantiation)
    ctx.clear_graphs()
    nodes, edges = synth.del_graph(rng, 300, 120)
    ctx.add_graph(nodes, edges)
    reads = synth.simulate_reads(rng, nodes, edges, 200, read_len=250, sub=0.03, indel_frac=0.3)
    bad += compare(ctx.align(reads), R.OracleGraph(nodes, edges).align_batch(reads), "long-reads")
    # config 2 sample vs oracle
    nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    ns = 300 if quick else 2000
    t0 = time.time()
    if R.have_ref():
        exp = R.ref_align_batch(nodes, edges, reads[:ns], threads=os.cpu_count())
        print("reference (%d threads): %.1f reads/s" % (os.cpu_count(), ns / (time.time() - t0)))
    else:
        exp = R.OracleGraph(nodes, edges).align_batch(reads[:ns])
    got = ctx.align(reads)
    bad += compare(got[:ns], exp, "config2")
    # timing
    blob, off = ctx.pack_reads(reads)
    for it in range(3):
        t0 = time.time()
        rec, ops = ctx.align_packed(blob, off)
        dt = time.time() - t0
        st = ctx.stats()
        print("config2 10k reads: e2e %.2f ms (%.0f reads/s)  fill %.3f ms  trace %.3f ms  -> kernels %.0f reads/s"
              % (dt * 1e3, len(reads) / dt, st["fill_ms"], st["trace_ms"],
                 len(reads) / ((st["fill_ms"] + st["trace_ms"]) * 1e-3)), flush=True)
    print("TOTAL MISMATCHES", bad)
    return 1 if bad else 0

if __name__ == "__main__":
    sys.exit(main())
