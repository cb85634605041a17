"""Parity of the CUDA path, called through the C-ABI (paragraph_b200/libpgalign.so), against
(1) the committed golden fixtures generated from the unmodified reference,
(2) the oracle on seeded fuzz inputs (multi-site batches, all flag combinations, both row-tile sizes),
(3) size-independent properties at BASELINE.json's full size (config 2: 10 000 x 150 bp reads).
Bar: bit-exact (score, graph_pos, CIGAR string, uniqueness / mapq, chosen strand, rewritten bases)."""
import numpy as np
import pytest

from conftest import golden_cases, strip_status
from oracle import refbind as R
from paragraph_b200 import capi, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx(built):
    c = capi.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("case", golden_cases(), ids=lambda c: c["name"])
def test_golden_fixtures(ctx, case):
    ctx.clear_graphs()
    ctx.add_graph(case["nodes"], [tuple(e) for e in case["edges"]])
    got = ctx.align(case["reads"], is_rev=case["is_rev"], flags=case["flags"])
    assert strip_status(got) == case["expected"]


def test_fuzz_multisite_batch(ctx):
    """300 random DAGs in ONE launch (multi-site batch), adversarial reads, vs the oracle."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(2024)
    ctx.clear_graphs()
    reads, sites, exp = [], [], []
    for _ in range(300):
        alpha = ["ACGT", "AC", "ACGTN", "ACGTRYN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet=alpha)
        rds = [r[:160] for r in synth.fuzz_reads(rng, nodes, edges, 16)]
        sid = ctx.add_graph(nodes, edges)
        reads += rds
        sites += [sid] * len(rds)
        exp += R.OracleGraph(nodes, edges).align_batch(rds)
    got = strip_status(ctx.align(reads, sites=sites))
    bad = [i for i, (g, e) in enumerate(zip(got, exp)) if g != e]
    assert not bad, (len(bad), got[bad[0]], exp[bad[0]])


@pytest.mark.parametrize("flags", [0xFFFFFFFF, 0, 1, 3, 5, 7])
def test_flags_and_strands(ctx, flags):
    R.set_fill_variant(0)
    rng = np.random.default_rng(flags & 0xFF)
    nodes, edges = synth.inv_graph(rng, 120, 60)
    reads = synth.simulate_reads(rng, nodes, edges, 150, read_len=90, sub=0.03, indel_frac=0.2, alternate=False)
    isrev = [int(x) for x in rng.integers(0, 2, size=len(reads))]
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev, flags=flags)
    assert strip_status(ctx.align(reads, is_rev=isrev, flags=flags)) == exp


def test_long_reads_up_to_8bit_limit(ctx):
    """161..250 bp reads run the 8-rows-per-lane instantiation."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(9)
    nodes, edges = synth.del_graph(rng, 400, 200)
    reads = synth.simulate_reads(rng, nodes, edges, 300, read_len=250, sub=0.02, indel_frac=0.3)
    reads += synth.simulate_reads(rng, nodes, edges, 50, read_len=161)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    assert strip_status(ctx.align(reads)) == R.OracleGraph(nodes, edges).align_batch(reads)


def test_reads_past_the_8bit_limit(ctx):
    """251..1024 bp reads: WIDE geometries (R = 10 / 16 / 32), scores >= 251 = gssw's 16-bit mode in the reference."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(19)
    nodes, edges = synth.del_graph(rng, 1100, 300)
    for rl in (251, 300, 320, 321, 450, 512, 513, 700, 1024):
        reads = synth.simulate_reads(rng, nodes, edges, 60 if rl <= 512 else 24, read_len=rl, sub=0.02, indel_frac=0.3)
        ctx.clear_graphs()
        ctx.add_graph(nodes, edges)
        exp = R.OracleGraph(nodes, edges).align_batch(reads)
        assert strip_status(ctx.align(reads)) == exp
        assert rl < 300 or max(e["score"] for e in exp) >= 251
    n = hi = 0
    for _ in range(40):  # adversarial long reads on bubble graphs, mixed lengths in one batch
        alpha = ["ACGT", "ACGT", "AC", "ACGTN"][int(rng.integers(0, 4))]
        nodes, edges = synth.bubble_graph(rng, n_nodes=int(rng.integers(1, 6)), max_len=int(rng.choice([300, 500, 700])),
                                          alphabet=alpha)
        reads = synth.fuzz_reads(rng, nodes, edges, 8, min_len=100, max_len=512 if _ % 4 else 1024)
        isrev = [i & 1 for i in range(len(reads))]
        ctx.clear_graphs()
        ctx.add_graph(nodes, edges)
        exp = R.OracleGraph(nodes, edges).align_batch(reads, is_rev=isrev)
        assert strip_status(ctx.align(reads, is_rev=isrev)) == exp
        n += len(reads)
        hi += sum(e["score"] >= 251 for e in exp)
    assert hi > 40


def test_16bit_mode_uniqueness_rule(ctx):
    """GraphAligner scans the 16-bit matrix byte-wise (GraphAligner.cpp:177-186): top scores 251..255 and >= 256."""
    from test_emulator import long_read_uniqueness_cases
    R.set_fill_variant(0)
    seen = set()
    for nodes, edges, reads in long_read_uniqueness_cases(np.random.default_rng(5)):
        ctx.clear_graphs()
        ctx.add_graph(nodes, edges)
        exp = R.OracleGraph(nodes, edges).align_batch(reads)
        assert strip_status(ctx.align(reads)) == exp
        seen |= {(min(max(e["score"], 250), 256), e["unique"]) for e in exp}
    assert (256, True) in seen and (250, False) in seen and any(250 < s < 256 for s, _ in seen)


def test_long_nodes_many_checkpoints(ctx):
    """config-5 shape (kb-sized nodes): hundreds of checkpoints / tiles per read, chunked scratch."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(10)
    nodes, edges = synth.inv_graph(rng, 300, 3000)
    reads = synth.simulate_reads(rng, nodes, edges, 96, alternate=False, indel_frac=0.1)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    ctx.set_scratch_limit(64 << 20)  # force several chunks
    try:
        got = strip_status(ctx.align(reads))
    finally:
        ctx.set_scratch_limit(64 << 30)
    assert got == R.OracleGraph(nodes, edges).align_batch(reads)


@pytest.mark.parametrize("n_nodes", [24, 60, 300])
def test_many_nodes_graph(ctx, n_nodes):
    """24 nodes: 4 warps per CTA; 60: the seed tables force 1 warp per CTA; 300: they move to HBM."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(12 + n_nodes)
    nodes, edges = synth.bubble_graph(rng, n_nodes=n_nodes, max_len=40 if n_nodes < 100 else 12,
                                      p_edge=0.15 if n_nodes < 100 else 0.02)
    reads = [r[:160] for r in synth.fuzz_reads(rng, nodes, edges, 200)]
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    assert strip_status(ctx.align(reads)) == R.OracleGraph(nodes, edges).align_batch(reads)


def test_short_node_events(ctx):
    """Runs of 1-3 bp nodes, chain links, merges with and without the node just finished, several sources: every branch of
    the fill's lean node events (pg_core.cuh: entry_word / seed_prefetch / node_event_pre), many sites in one batch."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(2025)
    graphs = synth.short_node_graphs(rng, 40)
    ctx.clear_graphs()
    reads, sites, exp = [], [], []
    for nodes, edges in graphs:
        sid = ctx.add_graph(nodes, edges)
        rd = [r[:150] for r in synth.fuzz_reads(rng, nodes, edges, 24, max_len=150)]
        reads += rd
        sites += [sid] * len(rd)
        exp += R.OracleGraph(nodes, edges).align_batch(rd)
    assert strip_status(ctx.align(reads, sites=sites)) == exp


@pytest.mark.parametrize("n_nodes", [30, 120])
def test_many_nodes_graph_long_reads(ctx, n_nodes):
    """WIDE geometries (R = 10 / 16) with many-node graphs: the 4-word node tables in shared memory with fewer warps per
    CTA (30 nodes) and in HBM (120 nodes), with the exact-match stage in front for half of the runs."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(77 + n_nodes)
    nodes, edges = synth.bubble_graph(rng, n_nodes=n_nodes, max_len=60, p_edge=0.1 if n_nodes < 100 else 0.03)
    reads = synth.fuzz_reads(rng, nodes, edges, 60, min_len=150, max_len=320) + \
        synth.fuzz_reads(rng, nodes, edges, 40, min_len=321, max_len=512) + \
        synth.fuzz_reads(rng, nodes, edges, 16, min_len=513, max_len=1024)
    try:
        for k in (0, 16):
            for part in (reads[:60], reads[:100], reads):  # R = 10 batch, then an R = 16 batch, then an R = 32 batch
                ctx.clear_graphs()
                ctx.add_graph(nodes, edges)
                ctx.set_stages(k, True, True)
                if k:
                    exp, _ = _cascade_expected(nodes, edges, part, k, None, True, True)
                else:
                    exp = R.OracleGraph(nodes, edges).align_batch(part)
                assert strip_status(ctx.align(part)) == exp, (k, len(part))
    finally:
        ctx.set_stages(0, True, False)


def test_empty_batch_is_legal(ctx):
    ctx.clear_graphs()
    ctx.add_graph(["ACGT"], [])
    assert ctx.align([]) == []
    blob, off = np.zeros(0, dtype=np.uint8), np.zeros(1, dtype=np.int32)
    rec, ops = ctx.align_packed(blob, off)
    assert len(rec) == 0 and len(ops) == 0


def test_errors_are_loud(ctx):
    ctx.clear_graphs()
    with pytest.raises(capi.PgError):
        ctx.align(["ACGT"])  # no graph
    with pytest.raises(capi.PgError):
        ctx.add_graph(["ACGT", "ACGT"], [(1, 0)])  # breaks topological order
    with pytest.raises(capi.PgError):
        ctx.add_graph(["ACGT", ""], [(0, 1)])  # empty node
    ctx.add_graph(["ACGT" * 100], [])
    with pytest.raises(capi.PgError):
        ctx.align(["ACGT" * 257])  # 1028 bp > PG_MAX_READ_LEN
    with pytest.raises(capi.PgError):
        ctx.align(["ACGT"], sites=[5])  # unknown site


def _consumed(cigar):
    import re
    q = r = 0
    for n, op in re.findall(r"(\d+)([MXNIDS])", cigar):
        n = int(n)
        if op in "MXNIS":
            q += n
        if op in "MXND":
            r += n
    return q, r


def test_full_size_config2_properties(ctx):
    """BASELINE.json configs[1] at full size: 10 000 reads x 150 bp on the 3-node DEL graph.
    Checked against the oracle on a 500-read sample, and on all reads through size-independent properties:
    CIGAR consumes exactly the read; score == #M - 4 #X - gaps; run twice -> identical; aligning the
    reverse-complemented read reports the same alignment on the other strand."""
    R.set_fill_variant(0)
    nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    got = strip_status(ctx.align(reads))
    assert got[:500] == R.OracleGraph(nodes, edges).align_batch(reads[:500])
    import re
    for g, r in zip(got, reads):
        q, _ = _consumed(g["cigar"])
        assert q == len(r), g
        # flatten over nodes: a gap that continues across a node boundary is ONE gap (one open)
        flat = []
        for n, op in re.findall(r"(\d+)([MXNIDS])", g["cigar"]):
            if flat and flat[-1][1] == op:
                flat[-1][0] += int(n)
            else:
                flat.append([int(n), op])
        sc = sum({"M": n, "X": -4 * n, "N": 0, "S": 0}.get(op, -(6 + n - 1)) for n, op in flat)
        assert sc == g["score"], (sc, g)
        assert 0 <= g["score"] <= len(r)
    again = strip_status(ctx.align(reads))
    assert again == got
    rc = strip_status(ctx.align([synth.revcomp(r) for r in reads[:2000]]))
    for a, b in zip(got[:2000], rc):
        if a["unique"] and b["unique"]:
            assert (a["score"], a["pos"], a["cigar"]) == (b["score"], b["pos"], b["cigar"])
            assert a["graph_reverse"] != b["graph_reverse"]


@pytest.mark.parametrize("w", ["16", "8"])
def test_other_lane_geometries(built, w, monkeypatch):
    """W = 16 / 8 lanes per task (2 / 4 tasks per warp) give the same bits as the default W = 32."""
    R.set_fill_variant(0)
    monkeypatch.setenv("PG_GEOM_W", w)
    c = capi.Context(0)
    try:
        rng = np.random.default_rng(31)
        reads, sites, exp = [], [], []
        for k in range(40):
            nodes, edges = synth.bubble_graph(rng, max_len=int(rng.choice([5, 20, 60, 200])), alphabet="ACGTN")
            rds = [r[:250] for r in synth.fuzz_reads(rng, nodes, edges, 12, max_len=160)]
            sid = c.add_graph(nodes, edges)
            reads += rds
            sites += [sid] * len(rds)
            exp += R.OracleGraph(nodes, edges).align_batch(rds)
        assert strip_status(c.align(reads, sites=sites)) == exp
        nodes, edges, rds = synth.config2(seed=3, n_reads=301)  # odd count: a half-empty warp at the tail
        c.clear_graphs()
        c.add_graph(nodes, edges)
        assert strip_status(c.align(rds)) == R.OracleGraph(nodes, edges).align_batch(rds)
    finally:
        c.close()


def test_staged_api_matches_one_call(ctx):
    nodes, edges, reads = synth.config2(seed=1, n_reads=512)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    blob, off = ctx.pack_reads(reads)
    rec1, ops1 = ctx.align_packed(blob, off)
    ctx.upload(blob, off)
    ctx.run()
    rec2, ops2 = ctx.download()
    c1 = [capi.format_cigar(r, ops1) for r in rec1]
    c2 = [capi.format_cigar(r, ops2) for r in rec2]
    assert c1 == c2
    for f in ("graph_pos", "score", "unique", "chose_reverse", "status"):
        assert (rec1[f] == rec2[f]).all()
    assert ctx.stats()["kernel_launches"] >= 4


def test_two_contexts_in_two_host_threads(built):
    """A context is single-owner, the library is re-entrant across contexts (one aligner per chunk/thread in the
    reference, src/c++/lib/grm/Align.cpp:107-110): two host threads, two contexts, same device, same answers."""
    import threading
    R.set_fill_variant(0)
    nodes, edges, reads = synth.config2(seed=11, n_reads=600)
    exp = R.OracleGraph(nodes, edges).align_batch(reads)
    out = {}

    def work(k):
        c = capi.Context(0)
        try:
            c.add_graph(nodes, edges)
            for _ in range(3):
                out[k] = strip_status(c.align(reads[k::2]))
        finally:
            c.close()

    th = [threading.Thread(target=work, args=(k,)) for k in range(2)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert out[0] == exp[0::2] and out[1] == exp[1::2]


def test_runs_on_a_caller_stream(ctx):
    """pg_set_stream: kernels are launched on the caller's (torch) stream and ordered with its events."""
    import torch
    nodes, edges, reads = synth.config2(seed=12, n_reads=256)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    base = strip_status(ctx.align(reads))
    s = torch.cuda.Stream()
    ctx.set_stream(s.cuda_stream)
    try:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        blob, off = ctx.pack_reads(reads)
        ctx.upload(blob, off)
        e0.record(s)
        ctx.run()
        e1.record(s)
        rec, ops = ctx.download()
        assert e0.elapsed_time(e1) > 0.0
        got = [capi.format_cigar(r, ops) for r in rec]
        assert got == [b["cigar"] for b in base]
    finally:
        ctx.set_stream(None)


# ---------------------------------------------------------------- exact-match stage in front of the DP (pg_set_stages)
def _cascade_expected(nodes, edges, reads, k, isrev=None, graph_matching=True, second_chance=False):
    """grm::CompositeAligner(path=true, graph=graph_matching): PathAligner first, gssw for the reads it leaves unmapped
    (lib/grm/CompositeAligner.cpp:90-107, 146-170), from the two oracles.  second_chance: the read filter holds NonUniq,
    so a non-unique exact match is rejected right after the stage (:97-103) and gssw aligns the bases PathAligner left."""
    pexp, cnt = R.OraclePathIndex(nodes, edges, k).align_batch(reads)
    og = R.OracleGraph(nodes, edges)
    gexp = og.align_batch(reads, is_rev=isrev) if graph_matching else [None] * len(reads)
    out = []
    for i, (p, g) in enumerate(zip(pexp, gexp)):
        if p["mapped"] and second_chance and graph_matching and not p["unique"]:
            d = og.align_batch([p["bases"]], is_rev=None if isrev is None else [isrev[i]])[0]
            d["stage"] = "gssw2" if p["graph_reverse"] else "gssw"
        elif p["mapped"]:
            d = {key: p[key] for key in ("pos", "score", "unique", "mapq", "graph_reverse", "bases", "cigar")}
            d["stage"] = "path"
        elif g is not None:
            d = dict(g)
            d["stage"] = "gssw"
        else:
            d = None
        out.append(d)
    return out, cnt


def test_path_stage_then_dp(ctx):
    from test_path_oracle import path_cases
    R.set_fill_variant(0)
    rng = np.random.default_rng(47)
    n = by_path = again = 0
    try:
        for nodes, edges, reads, k in path_cases(rng, 120):
            isrev = [i & 1 for i in range(len(reads))]
            ctx.clear_graphs()
            ctx.add_graph(nodes, edges)
            for second in (False, True):
                ctx.set_stages(k, True, second)
                exp, cnt = _cascade_expected(nodes, edges, reads, k, isrev, second_chance=second)
                got = strip_status(ctx.align(reads, is_rev=isrev))
                assert got == exp, (nodes, edges, k, second)
                st = ctx.path_stats()
                assert (st["attempted"], st["anchored"], st["mapped"]) == cnt
                again += sum(e["stage"] != "path" and p["mapped"] for e, p in
                             zip(exp, R.OraclePathIndex(nodes, edges, k).align_batch(reads)[0])) if second else 0
            n += len(reads)
            by_path += sum(e["stage"] == "path" for e in exp)
        assert by_path > n // 4 and by_path < n and again > 0
    finally:
        ctx.set_stages(0, True)


def test_path_stage_alone_leaves_reads_unmapped(ctx):
    rng = np.random.default_rng(48)
    nodes, edges = synth.del_graph(rng, 200, 80)
    reads = synth.simulate_reads(rng, nodes, edges, 200, read_len=100, sub=0.004, indel_frac=0.0)
    try:
        ctx.clear_graphs()
        ctx.add_graph(nodes, edges)
        ctx.set_stages(32, False)
        exp, cnt = _cascade_expected(nodes, edges, reads, 32, graph_matching=False)
        got = ctx.align(reads)
        assert 0 < cnt[2] < len(reads)
        for g, e in zip(got, exp):
            if e is None:
                assert g["status"] == 3
            else:
                assert g.pop("status") == 0 and g.pop("clipped") == 0
                assert g == e
        with pytest.raises(capi.PgError):
            ctx.set_stages(0, False)  # no stage at all
    finally:
        ctx.set_stages(0, True)


def test_path_stage_config2_mix(ctx):
    """10k-read shape of the benchmark: most reads differ from the haplotypes by >= 1 base (1 % substitutions), the
    rest are exact -- both populations must come out as the cascade of the two oracles says, across multiple sites."""
    R.set_fill_variant(0)
    rng = np.random.default_rng(49)
    try:
        ctx.clear_graphs()
        ctx.set_stages(32, True)
        reads, sites, exp = [], [], []
        for s in range(6):
            nodes, edges = synth.del_graph(rng, 300, int(rng.integers(40, 200)))
            rd = synth.simulate_reads(rng, nodes, edges, 150, read_len=150, sub=0.002, indel_frac=0.002)
            sid = ctx.add_graph(nodes, edges)
            e, _ = _cascade_expected(nodes, edges, rd, 32)
            reads += rd
            sites += [sid] * len(rd)
            exp += e
        got = strip_status(ctx.align(reads, sites=sites))
        assert got == exp
        frac = sum(e["stage"] == "path" for e in exp) / len(exp)
        assert 0.3 < frac < 0.95
    finally:
        ctx.set_stages(0, True)


def test_path_index_host_build_gives_the_same(built, monkeypatch):
    """The k-mer index is built on the device by default (pg_path_index_kernel); PG_PATH_HOST_INDEX=1 selects the host
    build (pg_host.hpp: build_path_index, also the fallback after a hash collision and for k > 64).  Both against the
    oracle cascade, incl. a k-mer length the device build does not take."""
    from test_path_oracle import path_cases
    R.set_fill_variant(0)
    for host_index, k_override in (("1", None), ("0", None), ("0", 70)):
        monkeypatch.setenv("PG_PATH_HOST_INDEX", host_index)
        c = capi.Context(0)
        try:
            rng = np.random.default_rng(51)
            for nodes, edges, reads, k in path_cases(rng, 40):
                k = k_override or k
                c.clear_graphs()
                c.add_graph(nodes, edges)
                c.set_stages(k, True, True)
                exp, cnt = _cascade_expected(nodes, edges, reads, k, None, True, True)
                assert strip_status(c.align(reads)) == exp, (host_index, k, nodes, edges)
                st = c.path_stats()
                assert (st["attempted"], st["anchored"], st["mapped"]) == cnt
        finally:
            c.close()


def test_staged_api_is_idempotent_with_second_chance(ctx):
    """Upload once, run many times (the staged API's contract): with the second-chance rule the exact-match kernels
    hand reverse-complemented bases to the DP; a later run over the same upload must see the uploaded bases again
    (same records, same stage, same strand), as one pg_align_batch call gives."""
    from test_path_oracle import path_cases
    rng = np.random.default_rng(53)
    flipped = 0
    # an inversion bubble: a read inside the inverted segment matches one branch as given and the other one reverse-
    # complemented -- two full-length exact matches, not unique -> second chance in the DP
    lf, mid, rf = synth.random_seq(rng, 60), synth.random_seq(rng, 90), synth.random_seq(rng, 60)
    twin = ([lf, mid, synth.revcomp(mid), rf], [(0, 1), (0, 2), (1, 3), (2, 3)])
    twin_reads = [mid[i:i + 60] for i in range(0, 30, 5)] + [synth.revcomp(mid[i:i + 60]) for i in range(0, 30, 7)]
    try:
        for nodes, edges, reads, k in [(twin[0], twin[1], twin_reads, 16)] + list(path_cases(rng, 60)):
            ctx.clear_graphs()
            ctx.add_graph(nodes, edges)
            ctx.set_stages(k, True, True)
            blob, off = ctx.pack_reads(reads)
            rec0, ops0 = ctx.align_packed(blob, off)
            rec0, ops0 = rec0.copy(), ops0.copy()
            ctx.upload(blob, off)
            for _ in range(3):
                ctx.run()
            rec1, ops1 = ctx.download()
            for f in ("graph_pos", "score", "unique", "chose_reverse", "status", "mapped_by", "cigar_len"):
                assert (rec0[f] == rec1[f]).all(), (f, k)
            assert [capi.format_cigar(r, ops0) for r in rec0] == [capi.format_cigar(r, ops1) for r in rec1]
            flipped += int((rec0["mapped_by"] != 1).sum())
        assert flipped > 0  # reads that went on to the DP
    finally:
        ctx.set_stages(0, True)


def test_imported_batch_cannot_be_run_and_clear_graphs_drops_the_batch(ctx):
    nodes, edges, reads = synth.config2(seed=14, n_reads=64)
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    blob, off = ctx.pack_reads(reads)
    rec, ops = ctx.align_packed(blob, off)
    rec, ops = rec.copy(), ops.copy()
    ctx.import_alignments([len(r) for r in reads], rec, ops)
    with pytest.raises(capi.PgError):
        ctx.run()  # alignments were imported, not reads: nothing to align
    ctx.upload(blob, off)
    ctx.clear_graphs()  # the uploaded site ids name graphs that are gone
    ctx.add_graph(nodes[:1], [])
    with pytest.raises(capi.PgError):
        ctx.run()
    ctx.clear_graphs()
    ctx.add_graph(nodes, edges)
    rec2, ops2 = ctx.align_packed(blob, off)
    assert (rec2["score"] == rec["score"]).all()


def test_tma_staging_equals_the_l1_path(built, monkeypatch):
    """The column codes of a task are staged into shared memory with a bulk async copy (cp.async.bulk + mbarrier) while
    the profile is built; compute-sanitizer's racecheck cannot pair that warp-scoped barrier with its wait and reports
    "potential hazards".  The same batches with the staging switched off (PG_NO_TMA=1: codes through L1, node tables
    from HBM) must give bit-identical records and CIGARs -- repeatedly, on batches that keep every SM busy with
    several waves of CTAs, so that a real race between the copy and the first loads would show."""
    import hashlib
    rng = np.random.default_rng(77)
    sites = []
    for k in range(40):
        nodes, edges = synth.site_graph(rng, ["DEL", "INS", "DUP", "INV"][k % 4])
        sites.append((nodes, edges, synth.simulate_reads(rng, nodes, edges, 400, alternate=False)))
    reads = [r for _, _, rd in sites for r in rd]
    sids = np.ascontiguousarray([i for i, s in enumerate(sites) for _ in s[2]], dtype=np.int32)
    digests = {}
    for no_tma in ("0", "1"):
        monkeypatch.setenv("PG_NO_TMA", no_tma)
        c = capi.Context(0)
        try:
            c.add_graphs([(n, e) for n, e, _ in sites])
            blob, off = c.pack_reads(reads)
            ds = set()
            for _ in range(5):
                rec, ops = c.align_packed(blob, off, sids)
                h = hashlib.sha1()
                for f in ("graph_pos", "score", "unique", "chose_reverse", "status", "cigar_len"):
                    h.update(np.ascontiguousarray(rec[f]).tobytes())
                h.update("".join(capi.format_cigar(r, ops) for r in rec[::7]).encode())
                ds.add(h.hexdigest())
            assert len(ds) == 1
            digests[no_tma] = ds.pop()
        finally:
            c.close()
    assert digests["0"] == digests["1"]
