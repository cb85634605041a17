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
        emubind.emu_align_batch(["ACGT" * 100], [], ["ACGT" * 63])  # 252 > 250
