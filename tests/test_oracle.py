"""The oracle (oracle/pg_oracle.c, a plain-C restatement of gssw + GraphAligner) is pinned against
(a) the reference's own unit-test vectors, (b) the committed fixtures generated from the unmodified reference,
(c) the unmodified reference itself (oracle/_ref) cell by cell, when it is available in this checkout."""
import numpy as np
import pytest

from conftest import golden_cases
from oracle import refbind as R
from paragraph_b200 import synth

# src/c++/test/test_paragraph_parts.cpp:113-144 (graphPos, graphCigar, graphAlignmentScore, graphMapq, isGraphReverseStrand)
REF_UNIT_EXPECTED = [
    (3, "0[8M]1[4M1X3M]3[8M]", 19, 60, False),
    (4, "0[7M]1[4M1X3M]3[6M]", 16, 60, True),
    (6, "0[5M]2[1M1X6M]3[6M]", 14, 60, False),
    (7, "0[4M]2[1M1X6M]3[6M]", 13, 60, False),
    (6, "0[5M]2[1M1X6M]3[6M]", 14, 60, True),
    (0, "0[11M]3[8M]", 19, 60, False),
]
REF_UNIT_BASES = ["AAAAAAAATTTTCTTTAAAAAAAA", "AAAAAAATTTTCTTTAAAAAA", "AAAAAGCGGGGGGAAAAAA", "AAAAGCGGGGGGAAAAAA",
                  "AAAAAGCGGGGGGAAAAAA", "AAAAAAAAAAAAAAAAAAA"]


def test_reference_unit_vectors(built):
    R.set_fill_variant(0)
    case = [c for c in golden_cases() if c["name"] == "ref_unit_paragraphtest"][0]
    got = R.OracleGraph(case["nodes"], case["edges"]).align_batch(case["reads"])
    for g, (pos, cigar, score, mapq, rev), bases in zip(got, REF_UNIT_EXPECTED, REF_UNIT_BASES):
        assert (g["pos"], g["cigar"], g["score"], g["mapq"], g["graph_reverse"], g["bases"]) == \
            (pos, cigar, score, mapq, rev, bases)
        assert g["unique"]


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_golden_fixtures(built, case):
    R.set_fill_variant(0)
    got = R.OracleGraph(case["nodes"], case["edges"]).align_batch(case["reads"], is_rev=case["is_rev"],
                                                                  flags=case["flags"])
    assert got == case["expected"]


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference in this checkout)")
def test_cell_by_cell_against_compiled_reference(built):
    """mH / mE / mF, per-node (score, ref_end, read_end), CIGAR: byte-identical to gssw itself."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(11)
    cells = 0
    for _ in range(120):
        alpha = ["ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60])), alphabet=alpha)
        reads = synth.fuzz_reads(rng, nodes, edges, 6)
        rg, og = R.RefGssw(nodes, edges), R.OracleGraph(nodes, edges)
        for r in reads:
            x, y = rg.fill_trace(r.upper()), og.fill_trace(r.upper())
            assert (x["stats"] == y["stats"]).all()
            assert (x["cigar"], x["pos"], x["score"]) == (y["cigar"], y["pos"], y["score"])
            for (h1, e1, f1), (h2, e2, f2) in zip(x["mats"], y["mats"]):
                assert (h1 == h2).all() and (e1 == e2).all() and (f1 == f2).all()
                cells += h1.size
        assert R.ref_align_batch(nodes, edges, reads) == og.align_batch(reads)
    assert cells > 100000


def _long_read_case(rng, max_len):
    """A bubble graph long enough for 251..max_len bp reads that mostly match (scores >= 251 -> gssw's 16-bit mode)."""
    alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
    nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 6)), max_len=int(rng.choice([300, 500, 700])),
                                      alphabet=alpha)
    reads = synth.fuzz_reads(rng, nodes, edges, 6, min_len=200, max_len=max_len)
    return nodes, edges, reads


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference in this checkout)")
def test_word_mode_cell_by_cell_against_compiled_reference(built):
    """Reads whose score reaches 251 make gssw redo the graph with 16-bit lanes (gssw.c:4001-4013): the restatement
    of gssw_sw_sse2_word (different lazy-F loop, end_ref starts at 0) must match mH / mE / mF, the per-node stats
    incl. is_byte, CIGAR -- and GraphAligner's uniqueness, which scans the 16-bit matrix through a uint8_t*."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(23)
    cells = word_fills = 0
    for _ in range(50):
        nodes, edges, reads = _long_read_case(rng, 500)
        rg, og = R.RefGssw(nodes, edges), R.OracleGraph(nodes, edges)
        for r in reads:
            x, y = rg.fill_trace(r.upper(), wide=True), og.fill_trace(r.upper(), wide=True)
            assert (x["stats"] == y["stats"]).all()
            assert (x["cigar"], x["pos"], x["score"]) == (y["cigar"], y["pos"], y["score"])
            word_fills += int(x["stats"][0, 3] == 0)
            for (h1, e1, f1), (h2, e2, f2) in zip(x["mats"], y["mats"]):
                assert (h1 == h2).all() and (e1 == e2).all() and (f1 == f2).all()
                cells += h1.size
        assert R.ref_align_batch(nodes, edges, reads) == og.align_batch(reads)
    assert word_fills > 40 and cells > 1000000


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference in this checkout)")
def test_word_mode_uniqueness_scan(built):
    """Top scores 251..255 and >= 256 on graphs where the read fits two nodes equally well: the reference's byte
    scan of the 16-bit matrix (GraphAligner.cpp:177-186) decides `unique`, not the number of nodes holding the top."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(5)
    seen = set()
    for L in (250, 251, 252, 253, 254, 255, 256, 257, 300):
        rep = synth.random_seq(rng, L)
        # two parallel copies of the same sequence between short flanks: every read has two equally good placements
        nodes = [synth.random_seq(rng, 20), rep, rep, synth.random_seq(rng, 20)]
        edges = [(0, 1), (0, 2), (1, 3), (2, 3)]
        reads = [rep, rep[: L - 3] + "A", nodes[0][-5:] + rep[: L - 5], synth.revcomp_exact(rep)]
        exp = R.ref_align_batch(nodes, edges, reads)
        assert exp == R.OracleGraph(nodes, edges).align_batch(reads)
        seen |= {(e["score"] >= 256, 251 <= e["score"] <= 255, e["unique"]) for e in exp}
    assert (True, False, True) in seen      # >= 256: the byte scan never matches -> "unique"
    assert any(s[1] for s in seen) and any(not s[0] and not s[1] and not s[2] for s in seen)
