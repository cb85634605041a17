"""GPU parity of the counting stage (pg_batch_count): read filters -> disambiguation -> fragment counts computed on
the device from the trace kernel's op words, against the oracle and the reference's golden vectors."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import refbind as R
from paragraph_b200 import capi, synth
from test_counts_oracle import check_tables, expected_tables, fuzz_site, load_phasing, phasing_inputs, unit_site

pytestmark = pytest.mark.gpu


def site_slices(shapes):
    nb, eb, out = 0, 0, []
    for nn, ne in shapes:
        out.append((nb, nb + nn, eb, eb + ne))
        nb += nn
        eb += ne
    return out


@pytest.mark.parametrize("use_filters", [True, False])
def test_multisite_counts_match_oracle(built, use_filters):
    rng = np.random.default_rng(99 + int(use_filters))
    sites = [fuzz_site(rng, ["DEL", "INS", "DUP", "INV", "bubble", "bubble"][k % 6]) for k in range(24)]
    ctx = capi.Context(0)
    try:
        all_reads, all_sites, all_rev, all_frag, fbase = [], [], [], [], 0
        for k, (nodes, edges, masks, reads, is_rev, frag) in enumerate(sites):
            sid = ctx.add_graph(nodes, edges)
            ctx.set_edge_labels(sid, masks)
            all_reads += reads
            all_sites += [sid] * len(reads)
            all_rev += is_rev
            all_frag += [fbase + f for f in frag]
            fbase += max(frag) + 1
        blob, off = ctx.pack_reads(all_reads)
        rec, ops = ctx.align_packed(blob, off, np.array(all_sites, dtype=np.int32))
        rec, ops = rec.copy(), ops.copy()
        kw = dict(remove_nonuniq=True, bad_align_frac=0.8, use_filters=use_filters)
        got = ctx.count(fragment=all_frag, is_rev=all_rev, **kw)
        assert len(got["path_words"]) == len(ops)
        r0 = 0
        for k, ((nodes, edges, masks, reads, is_rev, frag), (n0, n1, e0, e1)) in enumerate(
                zip(sites, site_slices([(len(s[0]), len(s[1])) for s in sites]))):
            g = R.OracleGraph(nodes, edges)
            al = g.align_batch(reads, is_rev=is_rev)
            g.close()
            o = R.oracle_count_site([len(s) for s in nodes], edges, masks, [len(r) for r in reads],
                                    [a["pos"] for a in al], [a["unique"] for a in al], [a["cigar"] for a in al],
                                    [a["graph_reverse"] for a in al], frag, **kw)
            sl = slice(r0, r0 + len(reads))
            assert (got["support"]["verdict"][sl] == o["support"]["verdict"]).all(), k
            assert (got["support"]["sequences"][sl] == o["support"]["sequences"]).all(), k
            assert (got["support"]["graph_reverse"][sl] == o["support"]["graph_reverse"]).all(), k
            for i in range(len(reads)):
                assert (got["support"]["path_off"][r0 + i] == rec["cigar_off"][r0 + i])
                assert R.support_sets(got["support"], got["path_words"], r0 + i) == \
                    R.support_sets(o["support"], o["path_words"], i), (k, i, al[i]["cigar"])
            assert (got["node_counts"][n0:n1] == o["node_counts"]).all(), k
            assert (got["edge_counts"][e0:e1] == o["edge_counts"]).all(), k
            fam = {m: v for (s, m), v in got["families"].items() if s == k}
            assert set(fam) == set(o["families"]), k
            for m in fam:
                assert (fam[m] == o["families"][m]).all(), (k, hex(m))
            r0 += len(reads)
    finally:
        ctx.close()


def test_counts_behind_the_cascade(built):
    """alignAndDisambiguate with path_sequence_matching on (the `paragraph` default): exact matches come from the
    exact-match stage (its own strand convention: graph_reverse = matched strand), non-unique ones get their second
    chance in gssw because the default filter chain rejects them, everything else is gssw -- then the same filter /
    disambiguation / counting.  Expected = the oracle's counting over the cascade of the two alignment oracles."""
    from test_gpu_parity import _cascade_expected
    rng = np.random.default_rng(321)
    kinds = ["DEL", "INS", "DUP", "INV", "bubble", "DEL"]
    ctx = capi.Context(0)
    try:
        ctx.set_stages(16, True, True)
        sites, all_reads, all_sites, all_rev, all_frag, fbase = [], [], [], [], [], 0
        for k in range(12):
            nodes, edges, masks, reads, is_rev, frag = fuzz_site(rng, kinds[k % 6])
            # make a good share of the reads exact so that the stage has something to map
            haps = synth.haplotypes(nodes, edges)
            for i in range(0, len(reads), 2):
                h = haps[int(rng.integers(0, len(haps)))]
                L = min(len(h), len(reads[i]))
                st = int(rng.integers(0, len(h) - L + 1))
                reads[i] = h[st:st + L] if rng.random() < 0.5 else synth.revcomp_exact(h[st:st + L])
            sites.append((nodes, edges, masks, reads, is_rev, frag))
            sid = ctx.add_graph(nodes, edges)
            ctx.set_edge_labels(sid, masks)
            all_reads += reads
            all_sites += [sid] * len(reads)
            all_rev += is_rev
            all_frag += [fbase + f for f in frag]
            fbase += max(frag) + 1
        blob, off = ctx.pack_reads(all_reads)
        rec, ops = ctx.align_packed(blob, off, np.array(all_sites, dtype=np.int32))
        assert (rec["mapped_by"] == 1).sum() > len(all_reads) // 8
        kw = dict(remove_nonuniq=True, bad_align_frac=0.8, use_filters=True)
        got = ctx.count(fragment=all_frag, is_rev=all_rev, **kw)
        r0 = 0
        for k, ((nodes, edges, masks, reads, is_rev, frag), (n0, n1, e0, e1)) in enumerate(
                zip(sites, site_slices([(len(s[0]), len(s[1])) for s in sites]))):
            al, _ = _cascade_expected(nodes, edges, reads, 16, is_rev, True, True)
            o = R.oracle_count_site([len(s) for s in nodes], edges, masks, [len(r) for r in reads],
                                    [a["pos"] for a in al], [a["unique"] for a in al], [a["cigar"] for a in al],
                                    [a["graph_reverse"] for a in al], frag, **kw)
            sl = slice(r0, r0 + len(reads))
            assert (got["support"]["verdict"][sl] == o["support"]["verdict"]).all(), k
            assert (got["support"]["graph_reverse"][sl] == o["support"]["graph_reverse"]).all(), k
            assert (got["support"]["sequences"][sl] == o["support"]["sequences"]).all(), k
            assert (got["node_counts"][n0:n1] == o["node_counts"]).all(), k
            assert (got["edge_counts"][e0:e1] == o["edge_counts"]).all(), k
            fam = {m: v for (s, m), v in got["families"].items() if s == k}
            assert set(fam) == set(o["families"]), k
            for m in fam:
                assert (fam[m] == o["families"][m]).all(), (k, hex(m))
            r0 += len(reads)
    finally:
        ctx.close()


def import_phasing(ctx, g, site, keep):
    rd, frag = phasing_inputs(g, keep)
    sid = ctx.add_graph(["A" * n for n in site["lens"]], site["edges"])
    ctx.set_edge_labels(sid, site["masks"])
    rec = np.zeros(len(rd), dtype=capi.RECORD_DTYPE)
    ops = []
    for i, a in enumerate(rd):
        w = capi.parse_cigar(a["cigar"])
        rec[i] = (a["pos"], a["score"], sum((x >> 3) & 0x1FFF for x in w if x & 7 == 5), int(a["unique"]), 0, 0, 0,
                  len(ops), len(w))
        ops += w
    ctx.import_alignments([a["len"] for a in rd], rec, np.array(ops, dtype=np.uint32))
    return rd, frag


def test_phasing_golden_on_device(built):
    """the reference's expected output for its phasing test data, fed through pg_batch_import + pg_batch_count"""
    g, site = load_phasing()
    ctx = capi.Context(0)
    try:
        rd, frag = import_phasing(ctx, g, site, lambda a: not a["verdict"].startswith("kmer"))
        got = ctx.count(fragment=frag, is_rev=[a["rev"] for a in rd])
        names, labs = site["names"], site["labs"]
        vname = {0: "MAPPED", 1: "nonuniq", 2: "bad_align", 3: "invalid"}
        for i, a in enumerate(rd):
            assert vname[int(got["support"]["verdict"][i])] == a["verdict"], (i, a["cigar"])
            if a["verdict"] != "MAPPED":
                continue
            nodes, edges, seqs = R.support_sets(got["support"], got["path_words"], i)
            assert sorted(names[x] for x in nodes) == sorted(a["nodes"]), i
            assert sorted(names[x] + "_" + names[y] for x, y in edges) == sorted(a["edges"]), i
            assert sorted(labs[k] for k in range(64) if seqs >> k & 1) == sorted(a["seqs"]), i
        got["families"] = {m: v for (s, m), v in got["families"].items()}
        check_tables(got, expected_tables(g, site))
    finally:
        ctx.close()


def test_paragraph_unit_golden_on_device(built):
    """ParagraphTest.Aligns (test_paragraph_parts.cpp:113-144): aligned and disambiguated (null filters) on the device"""
    case = json.load(open(os.path.join(GOLDEN_DIR, "counts_unit.json")))["ParagraphTest"]
    site = unit_site(case)
    ctx = capi.Context(0)
    try:
        sid = ctx.add_graph(site["seqs"], site["edges"])
        ctx.set_edge_labels(sid, site["masks"])
        rd = case["reads"]
        blob, off = ctx.pack_reads([a["bases"] for a in rd])
        rec, ops = ctx.align_packed(blob, off)
        for i, a in enumerate(rd):
            assert capi.format_cigar(rec[i], ops) == a["cigar"] and int(rec[i]["graph_pos"]) == a["pos"]
        got = ctx.count(use_filters=False)
        for i, a in enumerate(rd):
            nodes, edges, seqs = R.support_sets(got["support"], got["path_words"], i)
            assert sorted(site["names"][x] for x in nodes) == sorted(a["nodes"])
            assert sorted(site["names"][x] + "_" + site["names"][y] for x, y in edges) == sorted(a["edges"])
            assert sorted(site["labs"][k] for k in range(64) if seqs >> k & 1) == sorted(a["seqs"])
            assert int(got["support"]["graph_reverse"][i]) == int(a["rev"])
        # the same counts the reference build reports for these six fragments (LF 6, P1 2, Q1 3, RF 6 ...)
        assert got["node_counts"][:, 0].tolist() == [6, 2, 3, 6]
        assert got["edge_counts"][:, 0].tolist() == [2, 3, 1, 2, 3]
        assert {m: int(v[0, 0]) for (s, m), v in got["families"].items()} == \
            {1 << site["labs"].index("P"): 2, 1 << site["labs"].index("Q"): 3, 1 << site["labs"].index("D"): 1}
    finally:
        ctx.close()


def test_count_errors_and_edge_cases(built):
    rng = np.random.default_rng(3)
    nodes, edges = synth.del_graph(rng, 80, 40)
    reads = synth.simulate_reads(rng, nodes, edges, 40, read_len=60)
    ctx = capi.Context(0)
    try:
        with pytest.raises(capi.PgError):
            ctx.count()  # nothing has run yet
        s0 = ctx.add_graph(nodes, edges)
        s1 = ctx.add_graph(nodes, edges)
        blob, off = ctx.pack_reads(reads)
        sites = np.array([s0, s0, s1, s1] * 10, dtype=np.int32)  # mates (2k, 2k+1) share a site
        ctx.align_packed(blob, off, sites)
        # no labels: no families; every read its own fragment
        got = ctx.count()
        assert got["families"] == {}
        mapped = got["support"]["verdict"] == capi.V_MAPPED
        assert mapped.sum() > 20
        assert got["node_counts"][:, 0].sum() == got["node_counts"][:, 1].sum()  # one read per fragment
        assert (got["node_counts"][:, 1] == got["node_counts"][:, 2] + got["node_counts"][:, 3]).all()
        # a fragment may not span sites
        with pytest.raises(capi.PgError, match="spans two sites"):
            ctx.count(fragment=[0] * 40)
        with pytest.raises(capi.PgError, match="negative fragment"):
            ctx.count(fragment=[-1] * 40)
        # two labels on a DEL graph -> up to 3 families; one slot is not enough
        ctx.set_edge_labels(s0, synth.haplotype_labels(nodes, edges))
        ctx.set_edge_labels(s1, synth.haplotype_labels(nodes, edges))
        got = ctx.count(family_slots=8)
        assert len({m for (s, m) in got["families"] if s == s0}) >= 2
        with pytest.raises(capi.PgError, match="family_slots"):
            ctx.count(family_slots=1)
        # pairs as fragments: fragment totals count both mates
        pair = [i // 2 for i in range(40)]
        got = ctx.count(fragment=pair)
        both = (got["support"]["verdict"][0::2] == 0) & (got["support"]["verdict"][1::2] == 0)
        assert both.sum() > 5
        assert got["node_counts"][:3, 1].max() > got["node_counts"][:3, 0].max()  # READS > fragments
        # empty batch
        ctx.import_alignments([], np.zeros(0, dtype=capi.RECORD_DTYPE), np.zeros(0, dtype=np.uint32))
        got = ctx.count()
        assert got["node_counts"].sum() == 0 and len(got["support"]) == 0
    finally:
        ctx.close()
