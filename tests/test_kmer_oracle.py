"""grm::KmerAligner<K> (second stage of the reference's CompositeAligner cascade, src/c++/lib/grm/KmerAligner.cpp):
the oracle restatement (oracle/pg_oracle_kmer.c, incl. libstdc++'s heap algorithms the result depends on) is pinned
  (a) on the reference's own unit-test vectors (src/c++/test/test_kmeraligner.cpp:149-191: K = 10, 4-node graph,
      three paths, six reads -> position, CIGAR, score, strand, rewritten bases, MAPPED / BAD_ALIGN), and
  (b) against oracle/_ref = the UNMODIFIED KmerAligner.cpp compiled here, on seeded fuzz inputs for K = 10 and 16:
      repeats (ties in the candidate heap), N-filled source / sink nodes (soft clips), lower case, reads with 0-4
      mismatches, reads with repeated k-mers, many paths (heap eviction)."""
import numpy as np
import pytest

from oracle import refbind as R
from paragraph_b200 import synth

needs_ref = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference in this checkout)")

UNIT_NODES = ["AAAAAAAAAAA", "TTTTTTTT", "GGGGGGGG", "AAAAAAAAAAA"]
UNIT_EDGES = [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3)]
UNIT_PATHS = [[0, 1, 3], [0, 2, 3], [0, 3]]
UNIT_READS = ["AAAAAAAATTTTTTTTAAAAAAAA", "TTTTTTAAAAAAAATTTTTTT", "AAAAAGGGGGGGGAAAAAA", "AAAAGGGGGGGGAAAAAA",
              "TTTTTTCCCCCCCCTTTTT", "AAAAAAAAAAAAAAAAAAA"]
# test_kmeraligner.cpp:149-191 (graphPos is omitted from the JSON when 0)
UNIT_EXPECTED = [
    dict(status="mapped", bases="AAAAAAAATTTTTTTTAAAAAAAA", score=24, cigar="0[8M]1[8M]3[8M]", mapq=60, pos=3, unique=True, graph_reverse=False),
    dict(status="mapped", bases="AAAAAAATTTTTTTTAAAAAA", score=21, cigar="0[7M]1[8M]3[6M]", mapq=60, pos=4, unique=True, graph_reverse=True),
    dict(status="mapped", bases="AAAAAGGGGGGGGAAAAAA", score=19, cigar="0[5M]2[8M]3[6M]", mapq=60, pos=6, unique=True, graph_reverse=False),
    dict(status="mapped", bases="AAAAGGGGGGGGAAAAAA", score=18, cigar="0[4M]2[8M]3[6M]", mapq=60, pos=7, unique=True, graph_reverse=False),
    dict(status="mapped", bases="AAAAAGGGGGGGGAAAAAA", score=19, cigar="0[5M]2[8M]3[6M]", mapq=60, pos=6, unique=True, graph_reverse=True),
    # "this is the repeat one": the A-run fits at several offsets of the D path -> BAD_ALIGN, mapq 0
    dict(status="bad_align", bases="AAAAAAAAAAAAAAAAAAA", score=19, cigar="0[11M]3[8M]", mapq=0, pos=0, unique=False, graph_reverse=False),
]


def test_oracle_reproduces_the_reference_unit_test(built):
    got = R.OracleKmerIndex(UNIT_NODES, UNIT_EDGES, UNIT_PATHS, 10).align_batch(UNIT_READS)
    assert got == UNIT_EXPECTED


@needs_ref
def test_compiled_reference_reproduces_its_unit_test(built):
    got, cnt = R.ref_kmer_align_batch(UNIT_NODES, UNIT_EDGES, UNIT_PATHS, UNIT_READS, 10)
    assert got == UNIT_EXPECTED
    assert cnt == (6, 5)


def kmer_cases(rng, n_graphs, reads_per_graph=14):
    for gi in range(n_graphs):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN", "A"][int(rng.integers(0, 5))]
        nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 8)), max_len=int(rng.choice([6, 20, 60, 150])),
                                          alphabet=alpha)
        if gi % 5 == 2:  # N-filled source and sink, as vcf2paragraph writes them ("NNNNNNNNNN"): soft clips
            nodes = ["N" * 10] + nodes + ["N" * 10]
            edges = [(0, 1)] + [(a + 1, b + 1) for a, b in edges] + [(len(nodes) - 2, len(nodes) - 1)]
        if gi % 7 == 3:
            nodes = [s.lower() if rng.random() < 0.3 else s for s in nodes]
        paths = synth.haplotype_paths(nodes, edges, limit=int(rng.choice([1, 2, 4, 12])))
        if gi % 4 == 1 and paths:
            paths = paths + [paths[0]]  # the same path twice: equally good candidates that agree -> stays unique
        haps = ["".join(nodes[v] for v in p) for p in paths] or [synth.random_seq(rng, 40)]
        reads = []
        for _ in range(reads_per_graph):
            h = haps[int(rng.integers(0, len(haps)))]
            L = min(len(h), int(rng.integers(8, 161)))
            st = int(rng.integers(0, len(h) - L + 1))
            r = list(h[st:st + L])
            for _ in range(int(rng.choice([0, 0, 1, 2, 3, 4]))):
                p = int(rng.integers(0, L))
                r[p] = "ACGTN"[int(rng.integers(0, 5))]
            r = "".join(r)
            mode = rng.random()
            if mode < 0.08:
                r = synth.random_seq(rng, L)
            elif mode < 0.16 and L > 24:
                r = r[:12] + r[:12] + r[24:]  # a k-mer twice in the read
            if rng.random() < 0.5:
                r = synth.revcomp_exact(r)
            if rng.random() < 0.05:
                r = r.lower()
            reads.append(r or "A")
        yield nodes, edges, paths, reads


@needs_ref
@pytest.mark.parametrize("k", [10, 16])
def test_oracle_vs_compiled_reference_fuzz(built, k):
    rng = np.random.default_rng(1000 + k)
    n = mapped = bad = 0
    for nodes, edges, paths, reads in kmer_cases(rng, 400):
        isrev = [int(x) for x in rng.integers(0, 2, size=len(reads))]
        exp, cnt = R.ref_kmer_align_batch(nodes, edges, paths, reads, k, is_rev=isrev)
        got = R.OracleKmerIndex(nodes, edges, paths, k).align_batch(reads, is_rev=isrev)
        assert got == exp, (nodes, edges, paths, k, [(r, g, e) for r, g, e in zip(reads, got, exp) if g != e][:2])
        n += len(reads)
        mapped += sum(e["status"] == "mapped" for e in exp)
        bad += sum(e["status"] == "bad_align" for e in exp)
        assert cnt == (len(reads), sum(e["status"] == "mapped" for e in exp))
    assert mapped > n // 5 and bad > 20 and mapped + bad < n


def test_device_source_reproduces_the_reference_unit_test(built):
    """pg_kmer.cuh (the kernel's source compiled for the host, one lane per group) on the same vectors"""
    import emubind
    assert emubind.emu_kmer_align_batch(UNIT_NODES, UNIT_EDGES, UNIT_PATHS, UNIT_READS, 10) == UNIT_EXPECTED


@pytest.mark.parametrize("k", [10, 16, 5])
def test_device_source_vs_oracle_fuzz(built, k):
    import emubind
    rng = np.random.default_rng(2000 + k)
    n = 0
    for nodes, edges, paths, reads in kmer_cases(rng, 250):
        isrev = [int(x) for x in rng.integers(0, 2, size=len(reads))]
        exp = R.OracleKmerIndex(nodes, edges, paths, k).align_batch(reads, is_rev=isrev)
        got = emubind.emu_kmer_align_batch(nodes, edges, paths, reads, k, is_rev=isrev)
        assert got == exp, (nodes, edges, paths, k, [(r, g, e) for r, g, e in zip(reads, got, exp) if g != e][:2])
        n += len(reads)
    assert n > 3000
