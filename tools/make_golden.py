"""Generate tests/golden/*.json from the UNMODIFIED reference (oracle/_ref/libpgref.so, built from
/root/reference by oracle/Makefile).  Run in the build container; the fixtures are committed so that the
GPU box (which has no /root/reference) can check against them.   python tools/make_golden.py"""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np
from oracle import refbind as R
from paragraph_b200 import synth

OUT = os.path.join(os.path.dirname(__file__), "..", "tests", "golden")


def case(name, nodes, edges, reads, is_rev=None, flags=R.AF_ALL, note=""):
    res = R.ref_align_batch(nodes, edges, reads, is_rev=is_rev, flags=flags, threads=8)
    # the reference's decodeGraphAlignment + read filters (NonUniq, BadAlign at 0.8) on its own alignments
    filt = R.ref_filter_batch(nodes, edges, [len(r) for r in reads], [x["pos"] for x in res],
                              [x["unique"] for x in res], [x["cigar"] for x in res], 0.8).tolist()
    d = dict(name=name, note=note, nodes=nodes, edges=[list(e) for e in edges], reads=reads,
             is_rev=list(map(int, is_rev)) if is_rev is not None else None, flags=flags, expected=res,
             filters=filt, filters_note="per read: [decode_ok, query_clipped, nonuniq_filtered, badalign_filtered(0.8)]")
    with open(os.path.join(OUT, name + ".json"), "w") as f:
        json.dump(d, f, indent=0, separators=(",", ":"))
    print(name, len(reads), "reads")


def main():
    if not R.have_ref():
        raise SystemExit("oracle/_ref/libpgref.so missing: run `make -C oracle` where /root/reference exists")
    os.makedirs(OUT, exist_ok=True)
    # 1. the reference's own unit-test vectors (src/c++/test/test_paragraph_parts.cpp:52-144)
    case("ref_unit_paragraphtest",
         ["AAAAAAAAAAA", "TTTTTTTT", "GGGGGGGG", "AAAAAAAAAAA"], [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3)],
         ["AAAAAAAATTTTCTTTAAAAAAAA", "TTTTTTAAAGAAAATTTTTTT", "AAAAAGCGGGGGGAAAAAA", "AAAAGCGGGGGGAAAAAA",
          "TTTTTTCCCCCCGCTTTTT", "AAAAAAAAAAAAAAAAAAA"],
         note="graph+reads of ParagraphTest (test_paragraph_parts.cpp:52-89); expected pos/cigar/score/mapq/strand "
              "equal the JSON asserted at :113-144")
    # 2. config 2 sample (BASELINE.json configs[1]): 3-node DEL graph, 150 bp reads
    nodes, edges, reads = synth.config2(seed=42, n_reads=400)
    case("config2_sample", nodes, edges, reads, note="synth.config2(seed=42)[:400]")
    # 3. INS / DUP / INV / long-del (vcf2paragraph-shaped) sites
    rng = np.random.default_rng(1234)
    for kind in ("INS", "DUP", "INV"):
        nodes, edges = synth.site_graph(rng, kind, flank=200, sv_len=120)
        case("site_" + kind.lower(), nodes, edges, synth.simulate_reads(rng, nodes, edges, 120, alternate=False, indel_frac=0.1))
    nodes, edges = synth.long_del_graph(rng)
    case("site_longdel_two_sources", nodes, edges, synth.simulate_reads(rng, nodes, edges, 120, alternate=False))
    # 4. adversarial fuzz: IUPAC, N, lower case, 'U', '=' , repeats, random reads, all flag combinations
    for k, flags in enumerate([R.AF_ALL, 1, 3, 5, 7, 0]):
        nodes, edges = synth.bubble_graph(rng, n_nodes=5, max_len=80, alphabet="ACGTRYN" if k % 2 else "ACGT")
        reads = [r[:160] for r in synth.fuzz_reads(rng, nodes, edges, 60)]
        case("fuzz_flags_%x" % (flags & 0xF), nodes, edges, reads, is_rev=[i & 1 for i in range(len(reads))], flags=flags)
    # 5. edge cases: 1-base reads, unrelated reads (score ~0), all-N read, homopolymers, reads longer than the graph,
    #    long reads up to the 8-bit limit (250)
    nodes, edges = ["ACGTACGTAC", "GGGTTT", "ACGTACGTAC"], [(0, 1), (1, 2), (0, 2)]
    reads = ["A", "N", "C", "NNNNNNNN", "TTTTTTTTTTTTTTTTTTTT", "ACGTACGTACACGTACGTAC", "ACGTACGTACGGGTTTACGTACGTAC" * 3,
             "acgtacgtac", "ACGTNCGTAC", "GGGGGGGGGG", "U", "=", "ACGTACGTACGGGTTTACGTAC", "CATG" * 40]
    case("edge_small_graph", nodes, edges, reads)
    nodes, edges = synth.del_graph(rng, 300, 150)
    reads = synth.simulate_reads(rng, nodes, edges, 60, read_len=250, sub=0.03, indel_frac=0.3)
    reads += synth.simulate_reads(rng, nodes, edges, 20, read_len=161) + synth.simulate_reads(rng, nodes, edges, 20, read_len=160)
    case("long_reads_250", nodes, edges, reads, note="R=8 instantiation (161..250 bp) and the 160 bp boundary")
    nodes, edges = ["A" * 40, "A" * 30, "A" * 40], [(0, 1), (1, 2), (0, 2)]
    case("ties_homopolymer", nodes, edges, ["A" * 20, "A" * 50, "A" * 100, "AAAAACAAAAA", "T" * 30],
         note="maximal ties: many equal-scoring end cells / predecessor choices")


if __name__ == "__main__":
    main()
