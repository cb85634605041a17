"""CPU side of the BASELINE.json workloads (configs 2-5) and of the reference's own graph JSONs: the generators make
graphs the reference accepts, shaped as its vcf2paragraph makes them; the pinned oracle restatement and the CPU lane
emulator of the device source agree with the compiled reference (oracle/_ref) on them.  The full-size, every-read
comparison on the B200 is tests/test_gpu_configs.py."""
import numpy as np
import pytest

import emubind
from conftest import ref_graph_shapes, strip_status
from oracle import refbind as R
from paragraph_b200 import synth

needs_ref = pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference)")


def _key(d):
    return (d["pos"], d["score"], d["unique"], d["graph_reverse"], d["cigar"])


@pytest.mark.parametrize("kind", ["DEL", "INS", "DUP", "INV"])
@pytest.mark.parametrize("sv_len", [60, 300, 301, 900])
def test_vcf_site_graph_shape(kind, sv_len):
    """vcf2paragraph shape (graphUtils.py:23-101, GraphInput.cpp:81-89): "X" source / sink, ids ascend along every
    edge, no node over 300 bp, a cut node keeps 150 + 150 bp, every node lies on a source -> sink path."""
    rng = np.random.default_rng(sv_len)
    nodes, edges = synth.vcf_site_graph(rng, kind, sv_len)
    assert nodes[0] == "X" and nodes[-1] == "X"
    assert all(f < t for f, t in edges) and len(set(edges)) == len(edges)
    assert all(0 < len(s) <= 300 for s in nodes)
    n = len(nodes)
    has_in = {t for _, t in edges}
    has_out = {f for f, _ in edges}
    assert has_in == set(range(1, n)) and has_out == set(range(n - 1))
    if sv_len > 300:
        assert sum(len(s) == 150 for s in nodes) >= 2  # the two kept ends of a cut node
        assert len(nodes) >= 7
    else:
        assert len(nodes) == (7 if kind == "INV" else 6)


def test_long_del_shape_matches_the_reference_json():
    """share/test-data/paragraph/long-del/chr4-21369091-21376907.json: 7 nodes, 8 edges, two source branches."""
    ref = ref_graph_shapes()[0]
    assert ref["source"].endswith("long-del/chr4-21369091-21376907.json")
    nodes, edges = synth.vcf_site_graph(np.random.default_rng(0), "DEL", 7800)
    assert sorted(map(len, nodes)) == sorted(map(len, ref["nodes"]))
    assert len(edges) == len(ref["edges"]) == 8
    indeg = lambda es, n: sorted(sum(1 for e in es if e[1] == i) for i in range(n))
    assert indeg(edges, len(nodes)) == indeg([tuple(e) for e in ref["edges"]], len(ref["nodes"]))


def test_named_workloads_are_stable():
    a = synth.workload("config4_share", 0.004)
    b = synth.workload("config4_share", 0.004)
    assert a == b and len(a) == 5
    reads, sids, cells = synth.flatten_sites(a)
    assert len(reads) == len(sids) == sum(len(x[3]) for x in a) and cells > 0
    assert [k for k, *_ in a] == ["DEL", "INS", "DUP", "INV", "DEL"]


@needs_ref
@pytest.mark.parametrize("shape", ref_graph_shapes(), ids=lambda g: g["source"].split("/")[-1])
def test_reference_graph_shapes_oracle_and_emulator_vs_ref(built, shape):
    """Every graph JSON the reference ships under share/test-data/paragraph (loaded as GraphInput.cpp:51-161 does):
    haplotype reads and adversarial reads; compiled reference == oracle restatement == lane emulator of the kernels."""
    R.set_fill_variant(0)
    nodes, edges = shape["nodes"], [tuple(e) for e in shape["edges"]]
    rng = np.random.default_rng(len(nodes) * 1000 + len(edges))
    reads = synth.simulate_reads(rng, nodes, edges, 24, read_len=100, alternate=False) \
        + synth.fuzz_reads(rng, nodes, edges, 24, max_len=150)
    exp = R.ref_align_batch(nodes, edges, reads, threads=4)
    assert R.OracleGraph(nodes, edges).align_batch(reads) == exp
    got, _ = emubind.emu_align_batch(nodes, edges, reads)
    assert strip_status(got) == exp


@needs_ref
def test_multi_site_reference_driver_equals_per_site_calls(built):
    """pgref_align_sites (threads pull whole sites, Workflow.cpp:108-146) == one pgref_align_batch per site."""
    sites = synth.workload("config4_share", 0.0064)  # 8 sites
    multi = R.ref_align_sites([(n, e, r[:40]) for _, n, e, r in sites], threads=4)
    single = []
    for _, n, e, r in sites:
        single += R.ref_align_batch(n, e, r[:40], threads=1)
    assert [_key(x) for x in multi] == [_key(x) for x in single]


@needs_ref
@pytest.mark.parametrize("name,scale,per_site", [("config3", 0.006, 48), ("config4_share", 0.0064, 48), ("config5", 0.1, 24)])
def test_workload_shapes_emulator_vs_ref(built, name, scale, per_site):
    """A few sites of every workload shape through the lane emulator (the device source on the CPU) against the
    compiled reference: the shapes the GPU tests run at full size."""
    sites = [(n, e, r[:per_site]) for _, n, e, r in synth.workload(name, scale)]
    exp = R.ref_align_sites(sites, threads=4)
    at = 0
    for n, e, r in sites:
        got, _ = emubind.emu_align_batch(n, e, r)
        assert [_key(x) for x in strip_status(got)] == [_key(x) for x in exp[at:at + len(r)]], name
        at += len(r)
