"""CPU lane emulator of the CUDA path: tests/emu/pg_emu.cpp compiles the device source
(paragraph_b200/csrc/pg_core.cuh: wavefront step, node events, checkpoints, finalisation, strand choice,
tile traceback, CIGAR emission) for the host and drives 32 lanes in lock step.  It must agree with the
oracle / golden fixtures bit for bit; the -m gpu tests then check the same source running on the B200."""
import numpy as np
import pytest

import emubind
from conftest import golden_cases, strip_status
from oracle import refbind as R
from paragraph_b200 import synth


@pytest.fixture(autouse=True)
def _default_geometry():
    yield
    emubind.set_geometry(32)


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_emulator_golden(built, case):
    got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                     flags=case["flags"])
    assert strip_status(got) == case["expected"]


@pytest.mark.parametrize("w", [16, 8])
def test_emulator_other_geometries(built, w):
    """The kernels are templates over (W lanes per task, R rows per lane); W = 16 and 8 instantiations
    (PG_GEOM_W) must give the same bits as the default W = 32."""
    emubind.set_geometry(w)
    for case in golden_cases():
        got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                         flags=case["flags"])
        assert strip_status(got) == case["expected"], (w, case["name"])


def test_emulator_fuzz_vs_oracle(built):
    R.set_fill_variant(0)
    rng = np.random.default_rng(77)
    n = 0
    for _ in range(120):
        alpha = ["ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=alpha)
        reads = [r[:250] for r in synth.fuzz_reads(rng, nodes, edges, 12, max_len=int(rng.choice([60, 160, 250])))]
        flags = int(rng.choice([0xFFFFFFFF, 1, 3, 5, 7]))
        isrev = [i & 1 for i in range(len(reads))]
        exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
        got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
        assert strip_status(got) == exp
        n += len(reads)
    assert n > 1000


def test_emulator_rejects_long_reads(built):
    with pytest.raises(RuntimeError):
        emubind.emu_align_batch(["ACGT" * 300], [], ["ACGT" * 257])  # 1028 > 1024


def test_emulator_long_reads_16bit_mode(built):
    """251..512 bp reads (WIDE geometries: 16-bit checkpoints, 10-bit tile cells, region maxima).  Scores from 251 on
    put the reference into gssw's 16-bit mode (other lazy-F loop, byte-wise uniqueness scan of a 16-bit matrix)."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(31)
    n = hi = 0
    for _ in range(25):
        alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 6)), max_len=int(rng.choice([300, 500, 700])),
                                          alphabet=alpha)
        reads = synth.fuzz_reads(rng, nodes, edges, 6, min_len=200, max_len=512)
        isrev = [i & 1 for i in range(len(reads))]
        exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev)
        got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev)
        assert strip_status(got) == exp
        n += len(reads)
        hi += sum(e["score"] >= 251 for e in exp)
    assert hi > 30


def test_emulator_reads_up_to_1024(built):
    """513..1024 bp reads: the R = 32 geometry (32 rows per lane; tile cells H 11 bits, E / F 10 bits)."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(33)
    top = 0
    for it in range(10):
        if it % 2 == 0:
            nodes, edges = synth.del_graph(rng, 900, 300)
            reads = synth.simulate_reads(rng, nodes, edges, 3, read_len=[513, 640, 777, 1000, 1024][it // 2], sub=0.01, indel_frac=0.3)
        else:
            nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(2, 7)), max_len=600, alphabet=["ACGT", "ACGTN"][it % 4 == 1])
            reads = synth.fuzz_reads(rng, nodes, edges, 4, min_len=513, max_len=1024)
        isrev = [i & 1 for i in range(len(reads))]
        exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev)
        got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev)
        assert strip_status(got) == exp
        top = max([top] + [e["score"] for e in exp])
    assert top > 900


def long_read_uniqueness_cases(rng):
    """Two parallel copies of one sequence: every read has two equally good placements.  Top scores around the 8-bit
    limit (250 | 251..255 | >= 256) exercise the three branches of the uniqueness rule (pg_core.cuh n_top_rule)."""
    for L in (250, 251, 252, 253, 254, 255, 256, 257, 300, 400):
        rep = synth.random_seq(rng, L)
        nodes = [synth.random_seq(rng, 20), rep, rep, synth.random_seq(rng, 20)]
        edges = [(0, 1), (0, 2), (1, 3), (2, 3)]
        reads = [rep, rep[: L - 3] + "A", nodes[0][-5:] + rep[: L - 5], synth.revcomp_exact(rep), rep[: L // 2 + 3],
                 rep[L // 2 - 3:]]
        yield nodes, edges, reads


def test_emulator_16bit_uniqueness_rule(built):
    R.set_fill_variant(0)
    seen = set()
    for nodes, edges, reads in long_read_uniqueness_cases(np.random.default_rng(5)):
        exp = R.OracleGraph(nodes, edges).align_batch(reads)
        got, _ = emubind.emu_align_batch(nodes, edges, reads)
        assert strip_status(got) == exp
        seen |= {(min(max(e["score"], 250), 256), e["unique"]) for e in exp}
    assert (256, True) in seen and (250, False) in seen and any(250 < s < 256 for s, _ in seen)


def test_rev_plan_exhaustive(built):
    """The kernels fill only the reversed-graph halves rev_plan (pg_core.cuh) asks for.  Every combination of forward /
    reversed results, score order and flags: the lazily evaluated strand decision equals the full one, within two
    rounds, and the plan saves work (fewer than two halves per case on average)."""
    import ctypes as C
    lib = emubind.lib()
    lib.pgemu_rev_plan_check.restype = C.c_int
    cases, halves = C.c_long(0), C.c_long(0)
    assert lib.pgemu_rev_plan_check(C.byref(cases), C.byref(halves)) == 0
    assert cases.value == 9 * 3 * 3 * 3 * 3 * 3 and halves.value < cases.value


def test_emulator_speculative_dead_blocks(built):
    """pg_core.cuh: lane_step_dead -- blocks of 8 steps run with the collapsed recurrence (no E / F) while no gap is
    alive, checked afterwards and redone in full when a t > gap_open showed up (what pg_fill_kernel does when built
    with PG_SPEC_DEAD=1).  Results must not change: golden fixtures, a fuzz over bubble graphs / alphabets / flags
    (2-letter alphabets keep gaps alive almost everywhere, 4-letter ones exercise the dead blocks and the redo), all three
    geometries, and a config-2-shaped batch whose block statistics are the expected ones."""
    emubind.set_spec(0)  # the plain fill (kernels built with -DPG_SPEC_DEAD=0, kept for A/B)
    for case in golden_cases():
        got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                         flags=case["flags"])
        assert strip_status(got) == case["expected"], ("plain", case["name"])
    emubind.set_spec(1)
    try:
        before = emubind.spec_stats()
        for w in (32, 16, 8):
            emubind.set_geometry(w)
            for case in golden_cases():
                got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                                 flags=case["flags"])
                assert strip_status(got) == case["expected"], (w, case["name"])
        emubind.set_geometry(32)
        R.set_fill_variant(0)
        rng = np.random.default_rng(4242)
        for _ in range(80):
            alpha = ["ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 4))]
            nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200, 600])), alphabet=alpha)
            reads = [r[:250] for r in synth.fuzz_reads(rng, nodes, edges, 10, max_len=int(rng.choice([60, 160, 250])))]
            flags = int(rng.choice([0xFFFFFFFF, 1, 3, 5, 7]))
            isrev = [i & 1 for i in range(len(reads))]
            exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
            got, _ = emubind.emu_align_batch(nodes, edges, reads, is_rev=isrev, flags=flags)
            assert strip_status(got) == exp
        mid = emubind.spec_stats()
        assert mid[0] > before[0] and mid[1] > before[1] and mid[2] > before[2]  # dead, redone and alive blocks all occurred
        # config 2 (the bench workload): how many blocks run dead there
        nodes, edges, reads = synth.config2(seed=42, n_reads=60)
        exp = R.OracleGraph(nodes, edges).align_batch(reads)
        got, _ = emubind.emu_align_batch(nodes, edges, reads)
        assert strip_status(got) == exp
        dead, redone, alive, boundary = [b - a for a, b in zip(mid, emubind.spec_stats())]
        total = dead + redone + alive + boundary
        assert dead > 0.5 * total and redone < 0.2 * total, (dead, redone, alive, boundary)
        # EXPERIMENT (mode 2, emulator only so far): live gaps that cannot reach the best score
        # so far (value + remaining read rows < best) are dropped at a block start, and gaps opened inside a dead block
        # are tolerated under the same bound -- results must still not change, and more blocks run dead
        emubind.set_spec(2)
        for case in golden_cases():
            got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                             flags=case["flags"])
            assert strip_status(got) == case["expected"], ("prune", case["name"])
        rng = np.random.default_rng(99)
        for _ in range(30):
            nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([20, 60, 200, 600])), alphabet="ACGT")
            reads = [r[:250] for r in synth.fuzz_reads(rng, nodes, edges, 10, max_len=int(rng.choice([60, 160, 250])))]
            exp = R.OracleGraph(nodes, edges).align_batch(reads)
            got, _ = emubind.emu_align_batch(nodes, edges, reads)
            assert strip_status(got) == exp
        before2 = emubind.spec_stats()
        nodes, edges, reads = synth.config2(seed=42, n_reads=60)
        got, _ = emubind.emu_align_batch(nodes, edges, reads)
        assert strip_status(got) == R.OracleGraph(nodes, edges).align_batch(reads)
        dead2, redone2, alive2, boundary2 = [b - a for a, b in zip(before2, emubind.spec_stats())]
        assert dead2 > dead and alive2 < alive, (dead, alive, dead2, alive2)
    finally:
        emubind.set_spec(1)  # the default, as in the kernels


def test_emulator_short_node_events(built):
    """every branch of the lean node events (entry words, prefetched seeds, second boundary inside a sub-block) vs the oracle"""
    rng = np.random.default_rng(2024)
    for w in (32, 16):
        emubind.set_geometry(w)
        try:
            for nodes, edges in synth.short_node_graphs(rng, 8):
                reads = [r[:150] for r in synth.fuzz_reads(rng, nodes, edges, 10, max_len=150)]
                got, _ = emubind.emu_align_batch(nodes, edges, reads)
                assert strip_status(got) == R.OracleGraph(nodes, edges).align_batch(reads), (w, nodes, edges)
        finally:
            emubind.set_geometry(32)


def test_emulator_dead_boundary_blocks(built):
    """Node-boundary sub-blocks with the collapsed recurrence (the kernels' PG_DEAD_BOUNDARY): events between collapsed
    steps, attempts that fail on a t > gap_open or on a live E merged in by an event; with and without, against the oracle."""
    rng = np.random.default_rng(4242)
    cases = [(n, e, [r[:150] for r in synth.fuzz_reads(rng, n, e, 8, max_len=150)]) for n, e in synth.short_node_graphs(rng, 6)]
    nodes, edges, reads = synth.config2(seed=5, n_reads=16)
    cases.append((nodes, edges, reads))
    for s in synth.sites(77, 4, kinds=("DEL", "INS", "DUP", "INV"), max_reads=8):
        cases.append((s[1], s[2], s[3]))
    exp = [R.OracleGraph(n, e).align_batch(r) for n, e, r in cases]
    try:
        for on in (1, 0):
            emubind.set_dead_boundary(on)
            before = emubind.dead_boundary_stats()
            for (n, e, r), x in zip(cases, exp):
                got, _ = emubind.emu_align_batch(n, e, r)
                assert strip_status(got) == x, (on, n, e)
            after = emubind.dead_boundary_stats()
            if on:
                assert after[0] - before[0] > 100 and after[1] - before[1] > 5  # both outcomes were exercised
            else:
                assert after == before
    finally:
        emubind.set_dead_boundary(0)  # the default, as in the kernels
