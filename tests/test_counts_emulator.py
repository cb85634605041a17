"""The counting stage's device code (paragraph_b200/csrc/pg_count.cuh) compiled for the host by tests/emu and run on
alignments made by the emulated kernels: must reproduce the oracle (oracle/pg_oracle_counts.c) exactly --
per-read verdict / supported nodes / edges / sequences and the three count tables."""
import numpy as np
import pytest

import emubind
from oracle import refbind as R
from test_counts_oracle import fuzz_site


def compare(e, o, n_reads, tag):
    assert (e["support"]["verdict"] == o["support"]["verdict"]).all(), tag
    assert (e["support"]["sequences"] == o["support"]["sequences"]).all(), tag
    assert (e["support"]["graph_reverse"] == o["support"]["graph_reverse"]).all(), tag
    for i in range(n_reads):
        assert R.support_sets(e["support"], e["path_words"], i) == R.support_sets(o["support"], o["path_words"], i), (tag, i)
    assert (e["node_counts"] == o["node_counts"]).all(), tag
    assert (e["edge_counts"] == o["edge_counts"]).all(), tag
    assert set(e["families"]) == set(o["families"]), tag
    for m in o["families"]:
        assert (e["families"][m] == o["families"][m]).all(), (tag, hex(m))


@pytest.mark.parametrize("use_filters", [True, False])
def test_emulated_counting_matches_oracle(built, use_filters):
    rng = np.random.default_rng(2024 + int(use_filters))
    for it in range(18):
        kind = ["DEL", "INS", "DUP", "INV", "bubble", "bubble"][it % 6]
        nodes, edges, masks, reads, is_rev, frag = fuzz_site(rng, kind)
        kw = dict(remove_nonuniq=bool(it % 5), bad_align_frac=[0.8, 0.5, 0.95][it % 3], use_filters=use_filters)
        e = emubind.emu_count_site(nodes, edges, masks, reads, is_rev, frag, **kw)
        g = R.OracleGraph(nodes, edges)
        al = g.align_batch(reads, is_rev=is_rev)
        g.close()
        o = R.oracle_count_site([len(s) for s in nodes], edges, masks, [len(r) for r in reads],
                                [a["pos"] for a in al], [a["unique"] for a in al], [a["cigar"] for a in al],
                                [a["graph_reverse"] for a in al], frag, **kw)
        compare(e, o, len(reads), (it, kind))


def test_family_table_overflow_is_reported(built):
    rng = np.random.default_rng(5)
    nodes, edges, masks, reads, is_rev, frag = fuzz_site(rng, "bubble")
    while len(set(m for m in masks if m)) < 2:
        nodes, edges, masks, reads, is_rev, frag = fuzz_site(rng, "bubble")
    o = emubind.emu_count_site(nodes, edges, masks, reads, is_rev, frag, family_slots=64)
    if len(o["families"]) > 1:
        with pytest.raises(RuntimeError):
            emubind.emu_count_site(nodes, edges, masks, reads, is_rev, frag, family_slots=1)
