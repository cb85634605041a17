"""SURVEY.md 8f rank 1, first piece: the read filters paragraph applies right after alignment
(createReadFilter: NonUniq then BadAlign, src/c++/lib/paragraph/ReadFilter.cpp:73-90) evaluated from the engine's
record (unique flag + query_clipped) instead of decoding the CIGAR string on the host.
Expectations come from the reference's own decodeGraphAlignment + filter classes (tests/golden/*.json "filters")."""
import numpy as np
import pytest

import emubind
from conftest import golden_cases
from oracle import refbind as R
from paragraph_b200 import capi


def _expected(case):
    f = np.array(case["filters"], dtype=np.int64).reshape(-1, 4)
    return f


@pytest.mark.parametrize("case", [c for c in golden_cases() if c["flags"] & 1], ids=lambda c: c["name"])
def test_oracle_restates_bad_align(built, case):
    f = _expected(case)
    for i, e in enumerate(case["expected"]):
        if not f[i, 0]:
            continue  # the reference could not decode (empty CIGAR of a score-0 read): NonUniq removes it first
        bad, clipped = R.oracle_bad_align(e["cigar"], 0.8)
        assert (int(bad), clipped) == (int(f[i, 3]), int(f[i, 1])), (case["name"], i)


@pytest.mark.parametrize("case", [c for c in golden_cases() if c["flags"] & 1], ids=lambda c: c["name"])
def test_emulator_clipped_count(built, case):
    f = _expected(case)
    got, _ = emubind.emu_align_batch(case["nodes"], case["edges"], case["reads"], is_rev=case["is_rev"],
                                     flags=case["flags"])
    for i, g in enumerate(got):
        if f[i, 0]:
            assert g["clipped"] == f[i, 1], (case["name"], i)


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in golden_cases() if c["flags"] & 1], ids=lambda c: c["name"])
def test_gpu_filters_match_reference(built, case):
    f = _expected(case)
    ctx = capi.Context(0)
    try:
        ctx.add_graph(case["nodes"], [tuple(e) for e in case["edges"]])
        blob, off = ctx.pack_reads(case["reads"])
        rec, ops = ctx.align_packed(blob, off, flags=case["flags"])
        lens = np.diff(off)
        nonuniq, bad = capi.Context.read_filter(rec, lens, True, 0.8)
        ok = f[:, 0] == 1
        assert (rec["query_clipped"][ok] == f[ok, 1]).all()
        assert (nonuniq.astype(int) == f[:, 2]).all()
        assert (bad[ok].astype(int) == f[ok, 3]).all()
        # every CIGAR the engine emits for a uniquely mapped read must decode against the graph in the reference's
        # decodeGraphAlignment and consume exactly the read -- except gssw's own 'U' quirk (an alignment that starts on
        # a 'U' drops that base from the CIGAR, gssw.c:1662-1676, so the reference's CIGAR is one base short as well)
        for i in np.nonzero(rec["unique"] == 1)[0]:
            if "U" not in case["reads"][i].upper() and case["expected"][i]["cigar"]:
                assert ok[i], (case["name"], int(i))
    finally:
        ctx.close()
