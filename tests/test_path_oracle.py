"""grm::PathAligner (exact-match stage of the CompositeAligner cascade): the C restatement
oracle/pg_oracle_path.c is pinned against the UNMODIFIED reference PathAligner.cpp + graph-tools KmerIndex.cpp
compiled into oracle/_ref (oracle/ref_path.cpp), and against committed fixtures generated from it."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN_DIR
from oracle import refbind as R
from paragraph_b200 import synth


def path_cases(rng, n_graphs, reads_per_graph=12):
    """Graphs + reads that exercise the stage: exact haplotype reads (both strands), reads with one mismatch,
    reads shorter than k, low-complexity graphs where k-mers repeat, lower case, N."""
    for gi in range(n_graphs):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 8)), max_len=int(rng.choice([3, 12, 40, 150])),
                                          alphabet=alpha)
        if gi % 7 == 3:
            nodes = [s.lower() if rng.random() < 0.3 else s for s in nodes]
        k = int(rng.choice([4, 8, 16, 32]))
        haps = synth.haplotypes(nodes, edges) or [synth.random_seq(rng, 60)]
        reads = []
        for _ in range(reads_per_graph):
            h = haps[int(rng.integers(0, len(haps)))]
            L = int(rng.integers(max(1, k - 3), max(k + 1, min(len(h), 160)) + 1))
            L = min(L, len(h))
            st = int(rng.integers(0, len(h) - L + 1))
            r = h[st:st + L]
            mode = rng.random()
            if mode < 0.25 and L > 2:
                p = int(rng.integers(0, L))
                r = r[:p] + "ACGT"[(("ACGT".find(r[p].upper()) + 1) % 4)] + r[p + 1:]
            elif mode < 0.3:
                r = synth.random_seq(rng, L)
            if rng.random() < 0.5:
                r = synth.revcomp_exact(r)
            reads.append(r)
        yield nodes, edges, reads, k


@pytest.mark.skipif(not R.have_ref(), reason="oracle/_ref not built (no /root/reference in this checkout)")
def test_restatement_matches_compiled_reference(built):
    rng = np.random.default_rng(41)
    n = mapped = multi = 0
    for nodes, edges, reads, k in path_cases(rng, 250):
        exp, ecnt = R.ref_path_align_batch(nodes, edges, reads, kmer_len=k)
        got, gcnt = R.OraclePathIndex(nodes, edges, k).align_batch(reads)
        assert got == exp, (nodes, edges, k)
        assert gcnt == ecnt
        n += len(reads)
        mapped += sum(e["mapped"] for e in exp)
        multi += sum(e["mapped"] and not e["unique"] for e in exp)
    assert mapped > n // 4 and multi > 0


def test_golden_fixtures(built):
    with open(os.path.join(GOLDEN_DIR, "path_aligner.json")) as f:
        doc = json.load(f)
    for case in doc["cases"]:
        got, cnt = R.OraclePathIndex(case["nodes"], [tuple(e) for e in case["edges"]], case["k"]).align_batch(case["reads"])
        assert got == case["expected"], case["name"]
        assert list(cnt) == case["counters"]


def test_device_code_on_host_matches_oracle(built):
    """paragraph_b200/csrc/pg_path.cuh (the kernel's source, compiled for the host by tests/emu) + the host-side index
    builder (pg_host.hpp: build_path_index) against the oracle restatement: same reads mapped, same position, CIGAR,
    strand, uniqueness, counters."""
    import emubind
    rng = np.random.default_rng(43)
    n = mapped = 0
    for nodes, edges, reads, k in path_cases(rng, 250):
        exp, ecnt = R.OraclePathIndex(nodes, edges, k).align_batch(reads)
        got, gcnt = emubind.emu_path_align_batch(nodes, edges, reads, kmer_len=k)
        assert got == exp, (nodes, edges, k)
        assert gcnt == ecnt
        n += len(reads)
        mapped += sum(e["mapped"] for e in exp)
    assert mapped > n // 4


def test_device_code_on_host_golden(built):
    import emubind
    with open(os.path.join(GOLDEN_DIR, "path_aligner.json")) as f:
        doc = json.load(f)
    for case in doc["cases"]:
        got, cnt = emubind.emu_path_align_batch(case["nodes"], [tuple(e) for e in case["edges"]], case["reads"], case["k"])
        assert got == case["expected"], case["name"]
        assert list(cnt) == case["counters"]


def test_device_index_sizing_covers_every_kmer_path(built):
    """The device-side index build sizes its tables from a DP count of the k-mer paths (pg_host.hpp count_kmer_paths);
    the count must equal a real enumeration (the kernel's depth-first walk, restated on the host) and the node-list bound
    must cover what the enumeration needs."""
    import ctypes as C
    import emubind
    lib = emubind.lib()
    lib.pgemu_count_kmer_paths.restype = C.c_longlong
    lib.pgemu_count_kmer_paths.argtypes = [C.c_int, C.c_char_p, C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_int32),
                                           C.POINTER(C.c_int32), C.c_int, C.POINTER(C.c_longlong)]
    rng = np.random.default_rng(61)
    for gi in range(300):
        nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 12)), max_len=int(rng.choice([2, 6, 20, 80])),
                                          p_edge=float(rng.choice([0.2, 0.5, 0.9])))
        k = int(rng.choice([3, 8, 16, 32, 64]))
        blob, off, ef, et = R.pack_graph(nodes, edges)
        out = (C.c_longlong * 3)()
        p = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        lists = lib.pgemu_count_kmer_paths(len(nodes), blob, p(off), len(edges), p(ef), p(et), k, out)
        assert out[0] == out[2], (nodes, edges, k)
        assert 0 <= lists <= out[1], (nodes, edges, k)
