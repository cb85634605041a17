"""The k-mer stage of the cascade on the B200 (pg_set_paths / pg_set_kmer_stage -> pg_kmer_kernel, one warp per read)
against the oracle restatement of grm::KmerAligner (pinned on the compiled reference, tests/test_kmer_oracle.py):
the reference's unit-test vectors, a fuzz with the stage alone, the cascade k-mer stage -> gssw (a read the stage
maps but not uniquely goes on with the bases it leaves behind), and the full cascade exact-match -> k-mer -> gssw."""
import numpy as np
import pytest

from conftest import strip_status
from oracle import refbind as R
from paragraph_b200 import capi, synth
from test_kmer_oracle import UNIT_EDGES, UNIT_EXPECTED, UNIT_NODES, UNIT_PATHS, UNIT_READS, kmer_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.set_kmer_stage(0)
    c.close()


def _as_expected(g):
    """capi.align dict -> the oracle's dict of a k-mer-stage result"""
    g = dict(g)
    st, stage = g.pop("status"), g.pop("stage")
    g.pop("clipped")
    if st == 3:
        return dict(status="unmapped")
    assert st == 0 and stage.startswith("kmer")
    g["status"] = "mapped" if g["unique"] else "bad_align"
    return g


def test_kmer_stage_reference_unit_vectors(ctx):
    ctx.clear_graphs()
    ctx.add_graph(UNIT_NODES, UNIT_EDGES)
    ctx.set_paths(0, UNIT_PATHS)
    ctx.set_kmer_stage(10)
    ctx.set_stages(0, False)
    try:
        got = [_as_expected(g) for g in ctx.align(UNIT_READS)]
        assert got == UNIT_EXPECTED
        assert ctx.kmer_stats()["attempted"] == 6 and ctx.kmer_stats()["mapped"] == 5
    finally:
        ctx.set_stages(0, True)
        ctx.set_kmer_stage(0)


@pytest.mark.parametrize("k", [16, 10])
def test_kmer_stage_alone_fuzz(ctx, k):
    rng = np.random.default_rng(3000 + k)
    n = mapped = bad = 0
    try:
        ctx.set_kmer_stage(k)
        ctx.set_stages(0, False)
        for nodes, edges, paths, reads in kmer_cases(rng, 150):
            reads = [r[:500] for r in reads]
            isrev = [int(x) for x in rng.integers(0, 2, size=len(reads))]
            exp = R.OracleKmerIndex(nodes, edges, paths, k).align_batch(reads, is_rev=isrev)
            ctx.clear_graphs()
            ctx.add_graph(nodes, edges)
            ctx.set_paths(0, paths)
            got = [_as_expected(g) for g in ctx.align(reads, is_rev=isrev)]
            assert got == exp, (nodes, edges, paths, k, [(r, g, e) for r, g, e in zip(reads, got, exp) if g != e][:2])
            n += len(reads)
            mapped += sum(e["status"] == "mapped" for e in exp)
            bad += sum(e["status"] == "bad_align" for e in exp)
            assert ctx.kmer_stats()["mapped"] == sum(e["status"] == "mapped" for e in exp)
        assert mapped > n // 5 and bad > 5
    finally:
        ctx.set_stages(0, True)
        ctx.set_kmer_stage(0)


def _cascade_expected(nodes, edges, paths, reads, isrev, k, path_k=0):
    """grm::CompositeAligner(path = path_k > 0, kmer, graph) with the default filter's NonUniq rule after the
    exact-match stage (CompositeAligner.cpp:78-176), from the three oracles."""
    og = R.OracleGraph(nodes, edges)
    ok = R.OracleKmerIndex(nodes, edges, paths, k)
    pexp = R.OraclePathIndex(nodes, edges, path_k).align_batch(reads)[0] if path_k else [None] * len(reads)
    out = []
    for i, r in enumerate(reads):
        flips, p = 0, pexp[i]
        if p is not None and p["mapped"]:
            if p["unique"]:
                d = {key: p[key] for key in ("pos", "score", "unique", "mapq", "graph_reverse", "bases", "cigar")}
                d["stage"] = "path"
                out.append(d)
                continue
            r = p["bases"]  # rejected as non-unique: the later stages see what PathAligner left behind
            flips += int(p["graph_reverse"])
        e = ok.align_batch([r], is_rev=[isrev[i]])[0]
        if e["status"] == "mapped":
            d = {key: e[key] for key in ("pos", "score", "unique", "mapq", "graph_reverse", "bases", "cigar")}
            d["stage"] = "kmer" + ("2" if flips else "")
            out.append(d)
            continue
        if e["status"] == "bad_align":
            flips += int(e["bases"] != r)
            r = e["bases"]
        d = og.align_batch([r], is_rev=[isrev[i]])[0]
        d["stage"] = "gssw" + ("" if flips == 0 else str(flips + 1))
        out.append(d)
    return out


@pytest.mark.parametrize("path_k", [0, 16])
def test_kmer_stage_in_the_cascade(ctx, path_k):
    R.set_fill_variant(0)
    rng = np.random.default_rng(3100 + path_k)
    stages = {}
    try:
        ctx.set_kmer_stage(16)
        ctx.set_stages(path_k, True, True)
        for nodes, edges, paths, reads in kmer_cases(rng, 120):
            reads = [r[:250] for r in reads]
            isrev = [int(x) for x in rng.integers(0, 2, size=len(reads))]
            exp = _cascade_expected(nodes, edges, paths, reads, isrev, 16, path_k)
            ctx.clear_graphs()
            ctx.add_graph(nodes, edges)
            ctx.set_paths(0, paths)
            got = strip_status(ctx.align(reads, is_rev=isrev))
            assert got == exp, (nodes, edges, paths, path_k, [(r, g, e) for r, g, e in zip(reads, got, exp) if g != e][:2])
            for e in exp:
                stages[e["stage"]] = stages.get(e["stage"], 0) + 1
        assert stages.get("kmer", 0) > 100 and stages.get("gssw", 0) > 100 and stages.get("gssw2", 0) > 0, stages
        if path_k:
            assert stages.get("path", 0) > 50, stages
    finally:
        ctx.set_stages(0, True)
        ctx.set_kmer_stage(0)


def test_kmer_stage_staged_api_and_errors(ctx):
    nodes, edges = synth.del_graph(np.random.default_rng(5), 120, 50)
    paths = [[0, 1, 2], [0, 2]]
    reads = synth.simulate_reads(np.random.default_rng(6), nodes, edges, 200, read_len=100, sub=0.01, indel_frac=0.02)
    try:
        ctx.clear_graphs()
        ctx.add_graph(nodes, edges)
        ctx.set_paths(0, paths)
        with pytest.raises(capi.PgError):
            ctx.set_paths(0, [[0, 0]])  # not a path of the graph
        with pytest.raises(capi.PgError):
            ctx.set_kmer_stage(17)
        ctx.set_kmer_stage(16)
        blob, off = ctx.pack_reads(reads)
        rec0, ops0 = ctx.align_packed(blob, off)
        rec0, ops0 = rec0.copy(), ops0.copy()
        ctx.upload(blob, off)
        for _ in range(3):  # the stage hands flipped bases to the DP: every run must start from the uploaded ones
            ctx.run()
        rec1, ops1 = ctx.download()
        assert (rec0 == rec1).all() if rec0.dtype.names is None else all((rec0[f] == rec1[f]).all() for f in rec0.dtype.names if f != "cigar_off")
        assert [capi.format_cigar(r, ops0) for r in rec0] == [capi.format_cigar(r, ops1) for r in rec1]
        assert (rec0["mapped_by"] == 3).sum() > 100 and (rec0["mapped_by"] == 0).sum() > 0
    finally:
        ctx.set_kmer_stage(0)
