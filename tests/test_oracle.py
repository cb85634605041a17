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


def test_byte_overflow_is_reported(built):
    """Reads that would leave gssw's 8-bit mode are an explicit oracle error, never a silent answer."""
    R.set_fill_variant(0)
    seq = synth.random_seq(np.random.default_rng(3), 400)
    g = R.OracleGraph([seq], [])
    with pytest.raises(RuntimeError):
        g.align_batch([seq[:300]])
