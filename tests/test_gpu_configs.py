"""BASELINE.json configs 2-5 at FULL size on the B200, EVERY read compared with the compiled, unmodified reference
(oracle/_ref: gssw.c + GraphAligner.cpp, GraphAligner::alignRead semantics, src/c++/lib/grm/GraphAligner.cpp:308-404)
run on all host cores, plus every graph JSON the reference ships under share/test-data/paragraph.  Compared per read:
graph_pos, score, uniqueness (mapq), chosen strand, CIGAR string.  Bar: bit-exact, 0 differences.

  config 2      one 3-node DEL graph (500 bp flanks, D = 300), 10 000 x 150 bp reads
  config 3      1 000 mixed DEL / INS sites <= 500 bp, 30x 150 bp reads, one multi-site batch
  config 4      one GPU's share (1 250 sites) of the 10k-site DEL / INS / DUP / INV sweep, graphs shaped as
                vcf2paragraph makes them (<= 150 bp nodes after cutting, "X" source / sink, padding-base nodes)
  config 5      24 INV / DUP sites with 1-10 kb variant nodes, 1 000 reads each (chunked scratch, long-node path)
"""
import os

import numpy as np
import pytest

from conftest import ref_graph_shapes
from oracle import refbind as R
from paragraph_b200 import capi, synth

pytestmark = pytest.mark.gpu
STRIDE = 768  # bytes per reference CIGAR string in the comparison buffer


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


def _reference(sites):
    """-> (out6, cigar byte rows) from the compiled reference on all host cores; the pinned restatement only when
    oracle/_ref is absent (it is built in the build container and travels with the snapshot)."""
    pk = R.pack_sites(sites)
    if R.have_ref():
        return R.ref_align_sites_packed(pk, threads=os.cpu_count() or 8, cigar_stride=STRIDE)
    R.set_fill_variant(0)
    out = np.zeros((pk["n_reads"], 6), dtype=np.int32)
    cg = np.zeros((pk["n_reads"], STRIDE), dtype=np.uint8)
    i = 0
    for nodes, edges, rds in sites:
        for e in R.OracleGraph(nodes, edges).align_batch(rds):
            c = e["cigar"].encode()
            out[i] = (e["pos"], e["score"], e["unique"], e["mapq"], e["graph_reverse"], len(c))
            cg[i, :len(c)] = np.frombuffer(c, dtype=np.uint8)
            i += 1
    return out, cg


def _compare_every_read(ctx, site_list, flags=capi.AF_ALL):
    sites = [(n, e, r) for _, n, e, r in site_list]
    ctx.clear_graphs()
    for n, e, _ in sites:
        ctx.add_graph(n, e)
    reads, sids, _ = synth.flatten_sites(site_list)
    blob, off = ctx.pack_reads(reads)
    rec, ops = ctx.align_packed(blob, off, sids, flags)
    rec, ops = rec.copy(), ops.copy()
    exp, ecg = _reference(sites)
    assert (exp[:, 5] < STRIDE - 1).all(), "reference CIGAR longer than the comparison buffer"
    assert (rec["status"] == 0).all()
    bad = np.flatnonzero((rec["graph_pos"] != exp[:, 0]) | (rec["score"] != exp[:, 1])
                         | (rec["unique"] != exp[:, 2]) | (rec["chose_reverse"] != exp[:, 4]))
    assert bad.size == 0, (bad.size, int(bad[0]), rec[bad[0]], exp[bad[0]])
    nbad = 0
    for i in range(len(reads)):
        c = capi.format_cigar(rec[i], ops).encode()
        if len(c) != exp[i, 5] or c != ecg[i, :len(c)].tobytes():
            nbad += 1
            first = (i, c, ecg[i, :exp[i, 5]].tobytes()) if nbad == 1 else first
    assert nbad == 0, (nbad, first)
    return len(reads)


def test_config2_every_read_vs_reference(ctx):
    assert _compare_every_read(ctx, synth.workload("config2")) == 10000


def test_config3_every_read_vs_reference(ctx):
    n = _compare_every_read(ctx, synth.workload("config3"))
    assert n > 150000


def test_config4_share_every_read_vs_reference(ctx):
    w = synth.workload("config4_share")
    assert len(w) == 1250 and max(max(map(len, s[1])) for s in w) <= 300
    assert _compare_every_read(ctx, w) > 90000


def test_config5_every_read_vs_reference(ctx):
    w = synth.workload("config5")
    assert len(w) == 24 and max(max(map(len, s[1])) for s in w) > 8000
    assert _compare_every_read(ctx, w) == 24000


def test_config4_share_through_the_cascade(ctx):
    """The same vcf2paragraph-shaped sites with the exact-match stage in front (paragraph's default cascade): reads it
    does not map must come out of the DP exactly as without it."""
    w = synth.workload("config4_share", 0.2)
    reads, sids, _ = synth.flatten_sites(w)
    ctx.clear_graphs()
    for _, n, e, _ in w:
        ctx.add_graph(n, e)
    blob, off = ctx.pack_reads(reads)
    rec0, ops0 = ctx.align_packed(blob, off, sids)
    rec0, ops0 = rec0.copy(), ops0.copy()
    ctx.set_stages(32, True, False)
    try:
        rec1, ops1 = ctx.align_packed(blob, off, sids)
        rec1, ops1 = rec1.copy(), ops1.copy()
    finally:
        ctx.set_stages(0, True, False)
    dp = np.flatnonzero(rec1["mapped_by"] == 0)
    assert 0 < dp.size < len(reads)
    for f in ("graph_pos", "score", "unique", "chose_reverse", "status"):
        assert (rec0[f][dp] == rec1[f][dp]).all(), f
    for i in dp[:: max(1, dp.size // 5000)]:
        assert capi.format_cigar(rec0[i], ops0) == capi.format_cigar(rec1[i], ops1)


@pytest.mark.parametrize("shape", ref_graph_shapes(), ids=lambda g: g["source"].split("/")[-1])
def test_reference_graph_jsons(ctx, shape):
    """share/test-data/paragraph/*/*.json (long-del, pg-het-ins, pg-complex, haplo-complex, ...) as parity inputs:
    haplotype reads at 100 / 150 bp and adversarial reads."""
    nodes, edges = shape["nodes"], [tuple(e) for e in shape["edges"]]
    rng = np.random.default_rng(7 * len(nodes) + len(edges))
    reads = synth.simulate_reads(rng, nodes, edges, 300, read_len=150, alternate=False) \
        + synth.simulate_reads(rng, nodes, edges, 200, read_len=100, sub=0.03, indel_frac=0.2, alternate=False) \
        + synth.fuzz_reads(rng, nodes, edges, 200, max_len=250)
    assert _compare_every_read(ctx, [("json", nodes, edges, reads)]) == 700
