#!/usr/bin/env python
"""bench.py -- reads/sec aligned to graph (150 bp, DEL graph), the metric of BASELINE.json.

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torch.distributed.run)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (GraphAligner::alignRead semantics: 4 fills + traceback + uniqueness +
strand choice per read) over one batch = BASELINE.json configs[1]: one 3-node DEL graph (500 bp flanks,
D = 300), 10 000 synthetic 150 bp reads.  With N GPUs every rank aligns its own site of that shape (sites shard
across GPUs with no data-path collective -> weak scaling); value = reads of all ranks / max-over-ranks time.

  value  : kernels only, inputs already resident in HBM (pg_batch_run), timed with CUDA events on the launching
           stream, one event pair per step, a 256 MiB L2 flush between steps (untimed).
  e2e    : the same batch through the one-call C-ABI pg_align_batch with HOST buffers: pinned staging, H2D,
           kernels, D2H of records + CIGAR ops, inside the timed region.
  roofline / cpu_baseline : see DESIGN.md "Measurement".

Legs beside the headline (all in the one JSON line):
  configs           the other BASELINE.json shapes at full size on one GPU (N = 1 runs): config3 (1 000 DEL/INS sites),
                    config4_share (1 250 vcf2paragraph-shaped DEL/INS/DUP/INV sites = one GPU's eighth of config 4),
                    config5 (24 sites with 1-10 kb nodes): host-to-host reads/s, kernel-phase times, Tcell/s, and a
                    parity sample against the compiled reference (mismatches must be 0; every read is compared in
                    tests/test_gpu_configs.py)
  sweep_config4     every N: the FIXED 10 000-site config-4 sweep, LPT-sharded over the N ranks (strong scaling); the
                    timed pass registers the rank's graphs (pg_add_graphs), uploads, aligns, downloads and gathers
                    per-site summaries on rank 0
  e2e_mirror        N = 1: the C++ drop-in surface itself (pgb::grm::alignReads over vector<unique_ptr<Read>>, bases /
                    quals / CIGAR strings written back, MAPPED filter) and SitePipeline fed by a per-site producer
                    (tools/cpp/bench_mirror.cpp)
"""
import argparse
import glob
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from paragraph_b200 import synth  # noqa: E402

METRIC = "reads/sec to graph (150bp, DEL/INS); bit-exact score+CIGAR"
READS_PER_SITE = 10000
READ_LEN = 150
FLANK, DEL_LEN = 500, 300
G_COLS = 2 * FLANK + DEL_LEN
CELLS_PER_READ = 4 * READ_LEN * G_COLS  # 4 fills x L x G = 780 000
DPX_SLOTS_PER_CELL_PAIR = 5.0           # full recurrence: 4 half-rate DPX ops + 2 full-rate VIMNMX per packed cell pair
WORKLOAD = "config2: 3-node DEL graph (500 bp flanks, D=300), 10k synthetic 150 bp reads per GPU"
SWEEP_SITES = 10000


def step_profile():
    """Per-launch figures of ONE benchmark step from the newest committed ncu capture (profiles/r*_step_ncu.json, made
    by tools/ncu_step_summary.py from `ncu --set full` of this workload): ALU-pipe instructions and DRAM bytes of the
    step's fill launches.  The workload is fixed (seed 42), so the executed instruction counts are those of every step."""
    # (r<round><state>_step_ncu.json = the benchmark workload, config 2; other shapes carry their name: r02g_config4_step_ncu.json)
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r[0-9][0-9][a-z]_step_ncu.json")))
    if not files:
        return None
    try:
        doc = json.load(open(files[-1]))
        fills = [l for l in doc["launches"] if "pg_fill_kernel" in l["kernel"]]
        return dict(source="profiles/" + os.path.basename(files[-1]), what=doc.get("what"),
                    fill_launches=len(fills), fill_alu_inst=sum(l["alu_pipe_inst"] for l in fills),
                    fill_inst=sum(l["inst_executed"] for l in fills),
                    fill_dram_bytes=int(sum(l["dram_read_bytes"] + l["dram_write_bytes"] for l in fills)),
                    fill_duration_us=sum(l["duration_us"] for l in fills),
                    alu_peak_inst_per_sm_cycle=doc.get("alu_peak_inst_per_sm_cycle", 2.0),
                    alu_inst_source=fills[0].get("alu_pipe_inst_source") if fills else None)
    except Exception:
        return None


def workload(rank):
    return synth.config2(seed=42 + rank, n_reads=READS_PER_SITE, read_len=READ_LEN, flank=FLANK, d=DEL_LEN)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = []
        for name, col in (("hw_slowdown", 3), ("hw_thermal_slowdown", 4), ("sw_thermal_slowdown", 5), ("sw_power_cap", 6)):
            if any(len(r) > col and r[col].lower().startswith("active") for r in self.rows):
                reasons.append(name)
        return dict(sm_mhz=float(np.median(sm)) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=reasons, samples=len(sm))


# ------------------------------------------------------------------------------------------------------------------
# The reference's own CPU implementation on this box's host cores (oracle/_ref = unmodified gssw.c + GraphAligner.cpp,
# -O3 -msse4.1), on the config-2 batch; the time is what is spent inside the compiled library (reads and graph are
# packed before, results are not post-processed).  Two ways to use the cores, both timed, the best one reported:
#   threads   : one GraphAligner per thread over contiguous chunks of the batch, inside one process -- exactly
#               grm::alignReads (Align.cpp:107-153);
#   processes : the same chunks, one single-threaded process each -- what the reference's README recommends for
#               throughput (one multigrmpy.py per sample, README.md:111-117).  No reference source is touched.
# (Round 1 reported 0.7-3 k reads/s that fell with the sample size: that was this harness, not the reference -- its
# Python post-processing copied the whole CIGAR buffer once per read.)
# ------------------------------------------------------------------------------------------------------------------
_REF = {}


def _ref_chunk(span):
    lo, hi = span
    if _REF["batch"] is not None:
        _REF["batch"].run(lo, hi, 1)
    else:
        from oracle import refbind
        refbind.OracleGraph(_REF["nodes"], _REF["edges"]).align_batch(_REF["reads"][lo:hi])
    return hi - lo


class CpuReference:
    def __init__(self, nodes, edges, reads):
        from oracle import refbind
        self.refbind = refbind
        self.kind = "reference" if refbind.have_ref() else "port"
        self.nodes, self.edges, self.reads = nodes, edges, reads
        self.ncpu = os.cpu_count() or 1
        self.batch = refbind.RefBatch(nodes, edges, reads) if self.kind == "reference" else None  # packed once, outside every timed region
        _REF.update(nodes=nodes, edges=edges, reads=reads, batch=self.batch)
        self.pool = None

    def run_threads(self, n, threads):
        if self.batch is not None:
            return n / self.batch.run(0, n, threads)  # seconds inside the library call only
        t0 = time.perf_counter()
        self.refbind.OracleGraph(self.nodes, self.edges).align_batch(self.reads[:n])
        return n / (time.perf_counter() - t0)

    def run_processes(self, n, procs):
        if self.pool is None or self.pool_size != procs:
            import multiprocessing as mp
            self.close()
            self.pool = mp.get_context("fork").Pool(procs)
            self.pool_size = procs
            self.pool.map(_ref_chunk, [(i, i + 8) for i in range(0, 8 * procs, 8)], chunksize=1)  # warm: first touch in every worker
        step = (n + procs - 1) // procs
        spans = [(lo, min(n, lo + step)) for lo in range(0, n, step)]
        t0 = time.perf_counter()
        self.pool.map(_ref_chunk, spans, chunksize=1)
        return n / (time.perf_counter() - t0)

    def run(self, cfg, n):
        return self.run_processes(n, cfg[1]) if cfg[0] == "processes" else self.run_threads(n, cfg[1])

    def candidates(self):
        n = self.ncpu
        if self.kind != "reference":
            return [("threads", 1)]
        return [("threads", t) for t in sorted({1, max(1, n // 4), max(1, n // 2), n})] + [("processes", n)]

    def table(self, reps=3, budget_s=60.0):
        """-> (rows, best row).  Full batch for every multi-core row, warm, best of `reps`; a row whose warm-up rate says
        the full batch would take more than its share of the time budget runs a bounded sample instead (and says so);
        the single-core row runs a 2 000-read sample."""
        rows = []
        cands = self.candidates()
        share = budget_s / len(cands)
        for cfg in cands:
            n = len(self.reads) if cfg[1] > 1 else min(len(self.reads), 2000)
            probe = self.run(cfg, min(n, 64 * cfg[1]))  # warm-up, and a first estimate of the rate
            n = int(max(min(n, 64 * cfg[1]), min(n, probe * share / reps)))
            rates = [self.run(cfg, n) for _ in range(reps)]
            rows.append(dict(how=cfg[0], cores=cfg[1], reads=n, runs=len(rates), reads_per_s=round(max(rates), 1),
                             spread=round(max(rates) / min(rates), 3)))
        best = max(rows, key=lambda r: r["reads_per_s"])
        return rows, best

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()
            self.pool = None


def cpu_baseline(nodes, edges, reads):
    ref = CpuReference(nodes, edges, reads)
    try:
        rows, best = ref.table()
    finally:
        ref.close()
    return dict(value=best["reads_per_s"], unit="reads/s", cores=best["cores"], kind=ref.kind, how=best["how"],
                sample="%d reads of the %d-read config-2 batch, warm, best of %d runs per row, time inside the library call; "
                       "rows = grm::alignReads-style threads in one process, and one single-threaded process per core "
                       "(host has %d hardware threads)" % (best["reads"], len(reads), best["runs"], ref.ncpu),
                table=rows, same_config=best["reads"] == len(reads))


def bench_reference(args):
    """--impl reference: the reference CPU path on the host cores, full config-2 batch per step, the best way of using
    the cores found by one quick pass over the candidates; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nodes, edges, reads = workload(0)
    ref = CpuReference(nodes, edges, reads)
    try:
        rows, best = ref.table(reps=1, budget_s=30.0)
        cfg = (best["how"], best["cores"])
        n = len(reads)
        if n / best["reads_per_s"] * (args.steps + args.warmup) > 240.0:  # keep the whole run within a few minutes
            n = max(256, int(best["reads_per_s"] * 240.0 / (args.steps + args.warmup)))
        for _ in range(args.warmup):
            ref.run(cfg, n)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            ref.run(cfg, n)
        dt = time.perf_counter() - t0
    finally:
        ref.close()
    value = n * args.steps / dt
    line = dict(metric=METRIC, value=round(value, 1), unit="reads/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=round(dt / args.steps * 1e3, 3), higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="u8", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, step_sample="%d of the batch's %d reads per step" % (n, len(reads)), how=cfg[0],
                            cores=cfg[1], same_config=n == len(reads)),
                cpu_baseline=dict(value=round(value, 1), unit="reads/s", cores=cfg[1], kind=ref.kind, how=cfg[0],
                                  sample="%d reads per step x %d steps (host has %d hardware threads)" % (n, args.steps, ref.ncpu),
                                  selection=rows, same_config=n == len(reads)),
                e2e=dict(value=round(value, 1), unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------------------------
# legs on the other workload shapes
# ------------------------------------------------------------------------------------------------------------------
def config5_sweep(seed=5, n_sites=24, reads_per_site=1000):
    """BASELINE.json configs[4] as a packed sweep: INV / DUP graphs whose variant nodes are 1-10 kb, 1k reads per site."""
    rng = np.random.default_rng(seed)
    graphs, blobs, lens, read_ptr, cost, kinds = [], [], [], [0], [], []
    for i in range(n_sites):
        kind = "INV" if i % 2 else "DUP"
        n = int(rng.integers(1000, 10001))
        nodes, edges = synth.inv_graph(rng, 500, n) if kind == "INV" else synth.dup_graph(rng, n + 500, n)
        b, l = synth.site_reads_packed(rng, nodes, edges, reads_per_site, READ_LEN)
        graphs.append((nodes, edges))
        blobs.append(b)
        lens.append(l)
        read_ptr.append(read_ptr[-1] + reads_per_site)
        cost.append(4 * int(l.sum()) * sum(len(x) for x in nodes))
        kinds.append(kind)
    lens = np.concatenate(lens)
    off = np.zeros(len(lens) + 1, dtype=np.int32)
    off[1:] = np.cumsum(lens)
    return dict(graphs=graphs, blob=np.concatenate(blobs), off=off,
                site=np.repeat(np.arange(n_sites, dtype=np.int32), reads_per_site),
                read_ptr=np.asarray(read_ptr, dtype=np.int32), cost=np.asarray(cost, dtype=np.int64), kinds=kinds)


def parity_sample(sw, rec, ops, want_reads=2000):
    """Compare a sample of whole sites (spread over the sweep, about want_reads reads) with the compiled reference on all
    host cores: graph_pos, score, uniqueness, strand, CIGAR string.  -> (reads checked, mismatches)"""
    from oracle import refbind
    from paragraph_b200 import capi
    n_sites = len(sw["graphs"])
    per = max(1, int(len(sw["site"]) / n_sites))
    pick = sorted(set(np.linspace(0, n_sites - 1, max(1, min(n_sites, want_reads // per))).astype(int).tolist()))
    sub = synth.sweep_subset(sw, pick)
    pk = dict(n_sites=len(pick), blob=None)
    node_ptr, blob, off, edge_ptr, ef, et = capi.Context.pack_graphs(sub["graphs"])
    pk.update(node_ptr=node_ptr, blob=blob, off=off, edge_ptr=edge_ptr, ef=ef, et=et, read_ptr=sub["read_ptr"],
              rblob=sub["blob"].tobytes(), roff=sub["off"], n_reads=len(sub["site"]))
    if refbind.have_ref():
        exp, ecg = refbind.ref_align_sites_packed(pk, threads=os.cpu_count() or 8, cigar_stride=512)
    else:
        return 0, 0
    bad, at = 0, 0
    for s in pick:
        for i in range(sw["read_ptr"][s], sw["read_ptr"][s + 1]):
            r, e = rec[i], exp[at]
            c = capi.format_cigar(r, ops).encode()
            if (r["graph_pos"], r["score"], r["unique"], r["chose_reverse"]) != (e[0], e[1], e[2], e[4]) \
                    or r["status"] != 0 or len(c) != e[5] or c != ecg[at, :len(c)].tobytes():
                bad += 1
            at += 1
    return at, bad


def shape_leg(capi, torch, device, name, sw, reps=5):
    """One BASELINE.json shape at full size on one GPU: graphs resident (registered once), then
    e2e = pg_align_batch with host (page-locked) buffers per pass; kernel-phase times from a PG_SPLIT=1 context."""
    n_reads, cells = len(sw["site"]), int(sw["cost"].sum())
    packed = capi.Context.pack_graphs(sw["graphs"])
    pb, po = capi.PinnedArray(sw["blob"].shape, np.uint8), capi.PinnedArray(sw["off"].shape, np.int32)
    pb.array[:], po.array[:] = sw["blob"], sw["off"]
    out = {}
    ctx = capi.Context(device)
    t0 = time.perf_counter()
    ctx.add_graphs(packed=packed)
    reg_ms = (time.perf_counter() - t0) * 1e3
    for _ in range(2):
        ctx.align_packed(pb.array, po.array, sw["site"])
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        rec, ops = ctx.align_packed(pb.array, po.array, sw["site"])
        ts.append(time.perf_counter() - t0)
    e2e_s = float(np.median(ts))
    checked, bad = parity_sample(sw, rec, ops)
    ctx.close()
    os.environ["PG_SPLIT"] = "1"
    ctx1 = capi.Context(device)
    os.environ.pop("PG_SPLIT")
    ctx1.add_graphs(packed=packed)
    ctx1.upload(pb.array, po.array, sw["site"])
    for _ in range(2):
        ctx1.run()
    f, t = [], []
    for _ in range(3):
        ctx1.run()
        ctx1.download()
        s = ctx1.stats()
        f.append(s["fill_ms"])
        t.append(s["trace_ms"])
    ctx1.close()
    fill_ms, trace_ms = float(np.median(f)), float(np.median(t))
    out.update(sites=len(sw["graphs"]), reads=n_reads, cells=cells,
               e2e=dict(value=round(n_reads / e2e_s, 1), unit="reads/s", ms=round(e2e_s * 1e3, 3),
                        h2d_bytes=int(sw["blob"].nbytes + sw["off"].nbytes + sw["site"].nbytes),
                        d2h_bytes=int(rec.nbytes + ops.nbytes + 8)),
               register_graphs_ms=round(reg_ms, 3), fill_ms=round(fill_ms, 3), trace_ms=round(trace_ms, 3),
               fill_tcell_per_s=round(cells / fill_ms / 1e9, 3), kernels_tcell_per_s=round(cells / (fill_ms + trace_ms) / 1e9, 3),
               parity=dict(reads_checked=checked, mismatches=bad, against="oracle/_ref on all host cores; every read of "
                           "these shapes is compared in tests/test_gpu_configs.py"))
    return out


def sweep_leg(capi, torch, dist, multigpu, rank, world, device, reps=3):
    """The fixed 10 000-site config-4 sweep (vcf2paragraph-shaped DEL / INS / DUP / INV graphs, 30x 150 bp), LPT-sharded
    over the ranks: strong scaling.  A timed pass = register this rank's graphs (pg_add_graphs on the flat arrays of the
    C-ABI: table build for both orientations + upload), upload the reads, align, download records + CIGARs, reduce them
    to per-site summaries and gather those on rank 0.
    No data-path collective; time = max over ranks between two barriers."""
    sw = synth.packed_sweep(seed=44, n_sites=SWEEP_SITES)
    shards = multigpu.partition_sites([int(c) for c in sw["cost"]], world)
    mine = shards[rank]
    sub = synth.sweep_subset(sw, mine)
    pb, po = capi.PinnedArray(sub["blob"].shape, np.uint8), capi.PinnedArray(sub["off"].shape, np.int32)
    pb.array[:], po.array[:] = sub["blob"], sub["off"]
    ctx = capi.Context(device)
    counts = np.diff(sub["read_ptr"])
    cap = max(len(p) for p in shards)  # sites of the largest shard: the gathered tensors have one shape
    # the graphs in the C-ABI's input format (flat arrays, pg_add_graphs); flattening the Python lists of strings is the
    # harness's own cost (23 ms for 10 000 sites) and is reported in rank0_phases, not timed -- registering them is
    packed = capi.Context.pack_graphs(sub["graphs"])

    def one_pass():
        ctx.clear_graphs()
        ctx.add_graphs(packed=packed)
        rec, ops = ctx.align_packed(pb.array, po.array, sub["site"])
        uniq = np.add.reduceat(rec["unique"].astype(np.int64), sub["read_ptr"][:-1]) if len(rec) else np.zeros(0, np.int64)
        score = np.add.reduceat(rec["score"].astype(np.int64), sub["read_ptr"][:-1]) if len(rec) else np.zeros(0, np.int64)
        local = dict(sites=np.asarray(mine, dtype=np.int32), unique=uniq, score_sum=score, reads=counts, n_ops=len(ops))
        # per-site summaries to rank 0 as ONE fixed-size tensor gather (multigpu.gather_site_summaries)
        parts = multigpu.gather_site_summaries(local["sites"], [uniq, score, counts, np.full(len(mine), len(ops))], cap,
                                               dist if world > 1 else None, device="cuda")
        if parts is None:
            return None
        return [dict(sites=s_, unique=c_[0], score_sum=c_[1], reads=c_[2], n_ops=int(c_[3][0]) if len(s_) else 0)
                for s_, c_ in parts]

    one_pass()
    times, phases = [], None
    for _ in range(reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        merged = one_pass()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt[0]))
    # where a pass spends its time on this rank (untimed extra pass, phases measured one by one)
    t0 = time.perf_counter()
    ctx.clear_graphs()
    packed = capi.Context.pack_graphs(sub["graphs"])
    t1 = time.perf_counter()
    ctx.add_graphs(packed=packed)
    t2 = time.perf_counter()
    ctx.align_packed(pb.array, po.array, sub["site"])
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    phases = dict(pack_graphs_ms=round((t1 - t0) * 1e3, 2), register_ms=round((t2 - t1) * 1e3, 2),
                  upload_align_download_ms=round((t3 - t2) * 1e3, 2))
    ctx.close()
    if rank != 0:
        return None
    n_sites = sum(len(m["sites"]) for m in merged)
    n_reads = int(sum(int(m["reads"].sum()) for m in merged))
    best = min(times)
    loads = [sum(int(sw["cost"][i]) for i in p) for p in shards]
    return dict(what="fixed %d-site config-4 sweep (vcf2paragraph-shaped DEL/INS/DUP/INV), LPT shards over %d rank(s); a pass "
                     "registers the rank's graphs (pg_add_graphs on flat arrays), uploads, aligns, downloads and gathers per-site "
                     "summaries; rank0_phases.pack_graphs_ms = flattening the Python graph lists, outside the timed pass" % (SWEEP_SITES, world),
                scaling="strong", n_gpus=world, sites=n_sites, reads=n_reads, seconds=round(best, 5),
                reads_per_s=round(n_reads / best, 1), sites_per_s=round(n_sites / best, 1),
                cells=int(sw["cost"].sum()), tcell_per_s=round(float(sw["cost"].sum()) / best / 1e12, 3),
                load_imbalance=round(max(loads) / (sum(loads) / world), 4), unique_reads=int(sum(int(m["unique"].sum()) for m in merged)),
                rank0_phases=phases, passes=[round(t, 5) for t in times])


def mirror_legs(device, steps):
    """The C++ drop-in surface (tools/cpp/bench_mirror.cpp over paragraph_b200/csrc/host/pg_grm.hh): one process per leg on
    this GPU.  alignReads: config 2 through pgb::grm::alignReads; pipeline: 200 config-4-shaped sites through SitePipeline
    with a per-site read producer (align + filters + counts on the device)."""
    exe = os.path.join(ROOT, "tools", "cpp", "bench_mirror")
    if not os.path.exists(exe):
        return dict(error="tools/cpp/bench_mirror is not built (python -c 'import __graft_entry__ as g; g.build()')")
    out = {}
    ncpu = os.cpu_count() or 1
    threads = max(1, min(16, ncpu))
    with tempfile.TemporaryDirectory() as tmp:
        nodes, edges, reads = workload(0)
        f2 = os.path.join(tmp, "config2.txt")
        synth.write_workload_file(f2, [("DEL", nodes, edges, reads)])
        sw = synth.packed_sweep(seed=4, n_sites=1250)
        f4 = os.path.join(tmp, "config4_share.txt")
        synth.write_workload_file(f4, synth.sweep_as_site_list(sw))
        for key, path, mode, st in (("alignReads_config2", f2, "alignReads", steps), ("pipeline_config4_share", f4, "pipeline", 5)):
            for th in sorted({1, threads}):
                r = subprocess.run([exe, path, mode, str(st), "2", str(th), str(device)], capture_output=True, text=True, timeout=600)
                if r.returncode != 0:
                    out["%s_t%d" % (key, th)] = dict(error=r.stderr.strip()[-300:])
                    continue
                d = json.loads(r.stdout)
                out["%s_t%d" % (key, th)] = dict(value=d["reads_per_s"], unit="reads/s", host_threads=th, reads=d["reads"],
                                                sites=d["sites"], kept=d["kept"], seconds=d["seconds"],
                                                producer_seconds=d["producer_seconds"])
    out["what"] = ("alignReads: pgb::grm::alignReads(graph, paths, vector<unique_ptr<Read>>, NonUniq filter, gssw stage) per "
                   "step = pack + H2D + kernels + D2H + applyRecord (reverse complement, quals, CIGAR string) + filter + "
                   "MAPPED-only swap; pipeline: SitePipeline::addSite per site with the site's Read objects built inside the "
                   "timed region (stand-in for extractReads), alignAndCount on two engines; host threads = the `threads` "
                   "argument of grm::alignReads")
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-legs", action="store_true", help="headline only (profiling runs)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return bench_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    nodes, edges, reads = workload(rank)
    # the CPU baseline first: its process pool forks before this process holds a CUDA context
    base = None
    if not args.no_cpu_baseline and world == 1:
        try:
            base = cpu_baseline(nodes, edges, reads)
        except Exception as e:  # the baseline is reported, never required for the GPU number
            base = dict(value=None, error=str(e))
    elif not args.no_cpu_baseline:
        base = dict(value=None, note="measured at N=1 only")

    import torch
    import torch.distributed as dist
    from paragraph_b200 import capi, multigpu

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; paragraph_b200 has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    ctx = capi.Context(local_rank)
    # a dedicated (non-default) torch stream: the kernels are launched on it through the C-ABI and the CUDA events
    # below are recorded on the same stream
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    ctx.set_stream(stream.cuda_stream)
    ctx.add_graph(nodes, edges)
    blob, off = ctx.pack_reads(reads, pinned=True)  # the step's inputs live in page-locked host memory
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---------------- value: kernels only, inputs resident in HBM
    ctx.upload(blob, off)
    for _ in range(args.warmup):
        ctx.run()
    torch.cuda.synchronize()
    l0 = ctx.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for a, b in ev:
        flush.fill_(1)  # L2 flush, outside the event pair
        a.record(stream)
        ctx.run()
        b.record(stream)
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    launches = ctx.stats()["kernel_launches"] - l0
    rec, ops = ctx.download()
    total_ms = float(sum(step_ms))
    # per-kernel times for the roofline: the default run cuts the batch in two chunk pipelines on two streams whose
    # kernels overlap, so its per-launch event times are stretched; a context with PG_SPLIT=1 runs the same kernels one
    # after the other on the launching stream -- fill phase (forward fill, plan / pair, paired reversed-graph fills) and
    # traceback timed with the library's CUDA events around them, mean of 5 runs after 3 warm-ups
    os.environ["PG_SPLIT"] = "1"
    ctx1 = capi.Context(local_rank)
    os.environ.pop("PG_SPLIT")
    ctx1.set_stream(stream.cuda_stream)
    ctx1.add_graph(nodes, edges)
    ctx1.upload(blob, off)
    for _ in range(3):
        ctx1.run()
    fills, traces = [], []
    for _ in range(5):
        flush.fill_(1)
        ctx1.run()
        ctx1.download()
        s1 = ctx1.stats()
        fills.append(s1["fill_ms"])
        traces.append(s1["trace_ms"])
    st = dict(fill_ms=float(np.mean(fills)), trace_ms=float(np.mean(traces)))
    ctx1.close()
    # ---------------- e2e: host buffers in, host buffers out, through pg_align_batch
    for _ in range(2):
        ctx.align_packed(blob, off)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rec, ops = ctx.align_packed(blob, off)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    h2d = int(blob.nbytes + off.nbytes)
    d2h = int(rec.nbytes + ops.nbytes + 8)
    # ---------------- the same call from two host threads on two contexts (double buffering): while one
    # context's kernels run, the other's H2D / D2H copies and host-side bookkeeping proceed
    ctx2 = capi.Context(local_rank)
    ctx2.add_graph(nodes, edges)
    blob2, off2 = ctx2.pack_reads(reads, pinned=True)

    def worker(cx, b, o, n):
        for _ in range(n):
            cx.align_packed(b, o)

    for cx, b, o in ((ctx, blob, off), (ctx2, blob2, off2)):
        worker(cx, b, o, 2)
    barrier()
    th = [threading.Thread(target=worker, args=(cx, b, o, args.steps)) for cx, b, o in ((ctx, blob, off), (ctx2, blob2, off2))]
    t0 = time.perf_counter()
    for t_ in th:
        t_.start()
    for t_ in th:
        t_.join()
    torch.cuda.synchronize()
    two_s = time.perf_counter() - t0
    ctx2.close()
    # ---------------- reads -> count tables: host reads in, node/edge/path-family fragment counts
    # out; the CIGARs never leave the device.  Reads 2k, 2k+1 form a fragment; labels = the two haplotypes (REF/ALT).
    ctx.set_edge_labels(0, synth.haplotype_labels(nodes, edges))
    pairs = np.arange(READS_PER_SITE, dtype=np.int32) // 2
    for _ in range(2):
        ctx.upload(blob, off)
        ctx.run()
        cnt = ctx.count(fragment=pairs, want_support=False)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.upload(blob, off)
        ctx.run()
        cnt = ctx.count(fragment=pairs, want_support=False)
    torch.cuda.synchronize()
    cnt_s = time.perf_counter() - t0
    cst = ctx.stats()
    # ---------------- the cascade of `paragraph` (path_sequence_matching on, main/paragraph.cpp:60): exact-match stage
    # (grm::PathAligner, k = 32) in front of the DP, same host-buffer call; reads it maps skip the DP kernels
    ctx.set_stages(32, True, True)
    for _ in range(2):
        ctx.align_packed(blob, off)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.align_packed(blob, off)
    torch.cuda.synchronize()
    casc_s = time.perf_counter() - t0
    pst = ctx.path_stats()
    cascade_kernels = ctx.stats()
    ctx.set_stages(0, True, False)
    clocks = sampler.stop()
    cnt_d2h = int(cnt["node_counts"].shape[0] * 16 + cnt["edge_counts"].shape[0] * 16
                  + sum(16 + 16 * v.shape[0] for v in cnt["families"].values()))
    ctx.close()

    # ---------------- the config-4 sweep, sharded over the ranks (every N), and the other shapes / the C++ surface (N = 1)
    sweep = configs = mirror = None
    if not args.no_extra_legs:
        try:
            sweep = sweep_leg(capi, torch, dist, multigpu, rank, world, local_rank)
        except Exception as e:
            sweep = dict(error=repr(e))
        if world == 1:
            configs = {}
            for name, make in (("config3", lambda: synth.packed_sweep(seed=3, n_sites=1000, kinds=("DEL", "INS"), shaped=False)),
                               ("config4_share", lambda: synth.packed_sweep(seed=4, n_sites=1250)),
                               ("config5", config5_sweep)):
                try:
                    configs[name] = shape_leg(capi, torch, local_rank, name, make())
                except Exception as e:
                    configs[name] = dict(error=repr(e))
            try:
                mirror = mirror_legs(local_rank, args.steps)
            except Exception as e:
                mirror = dict(error=repr(e))

    t = torch.tensor([total_ms, e2e_s * 1e3, cnt_s * 1e3, casc_s * 1e3, two_s * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, cnt_ms, casc_ms, two_ms = float(t[0]), float(t[1]), float(t[2]), float(t[3]), float(t[4])
    n_reads_all = READS_PER_SITE * world
    value = n_reads_all * args.steps / (total_ms * 1e-3)
    e2e_value = n_reads_all * args.steps / (e2e_ms * 1e-3)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        # Dominant kernel = pg_fill_kernel.  Roofline = the ALU pipe (packed-int16 DPX and the other integer instructions
        # issue there): EXECUTED ALU-pipe warp instructions of the step's fill launches (counted by ncu on this fixed
        # workload, profiles/) / the fill phase's live duration, against the pipe's issue rate.  Since the speculative
        # dead blocks the kernel no longer executes a fixed number of operations per cell, so the algorithmic cell rate
        # is reported beside it, not as the fraction.
        sm_max = (clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        fill_ms = st["fill_ms"]
        prof = step_profile()
        ach_cells = READS_PER_SITE * CELLS_PER_READ / (fill_ms * 1e-3) if fill_ms > 0 else 0.0
        dpx_peak_cells = 63.2 * 148 * sm_max * 1e6 * 2.0 / DPX_SLOTS_PER_CELL_PAIR
        roof = dict(bound="alu", kernel="pg_fill_kernel<5,32> (fill phase of a step: forward-graph launch + plan / pair + "
                                        "paired reversed-graph launches, run back to back)",
                    note="integer ALU pipe (packed-int16 DPX): the max-plus recurrence is neither HBM-bound (roofline_hbm) nor "
                         "a tensor-core contraction (SURVEY.md 8d, DESIGN.md 4)",
                    algorithmic_gcell_per_s=round(ach_cells / 1e9, 1),
                    full_recurrence_peak_gcell_per_s=round(dpx_peak_cells / 1e9, 1),
                    algorithmic_over_full_recurrence_peak=round(ach_cells / dpx_peak_cells, 4),
                    algorithmic_note="4 L G cells per read / fill-phase time; the denominator is the issue-rate ceiling of the "
                                     "FULL recurrence (5 issue slots per packed cell pair, tools/ubench/dpx_ubench.cu: 63.2 "
                                     "lane-ops/clk/SM); rev_plan skips a quarter of the fills and dead blocks run a one-operation "
                                     "step, so this ratio may pass 1 and is not the roofline fraction")
        if prof and fill_ms > 0:
            peak_inst = prof["alu_peak_inst_per_sm_cycle"] * 148 * sm_max * 1e6
            ach_inst = prof["fill_alu_inst"] / (fill_ms * 1e-3)
            roof.update(achieved=round(ach_inst / 1e9, 1), peak=round(peak_inst / 1e9, 1), unit="G ALU-pipe warp-inst/s",
                        frac=round(ach_inst / peak_inst, 4),
                        traffic=dict(dram_bytes_per_step=prof["fill_dram_bytes"], launches=prof["fill_launches"],
                                     algorithmic_bytes_per_step=READS_PER_SITE * 200, source=prof["source"]),
                        alu_inst_per_step=int(prof["fill_alu_inst"]), inst_per_step=int(prof["fill_inst"]),
                        ncu_fill_duration_us=round(prof["fill_duration_us"], 1), profile=prof["source"],
                        peak_source="ALU pipe issue rate: %.1f warp-inst/clk/SM (B300_MICROARCH.md pipe rates; ncu's own "
                                    "pct_of_peak for this pipe uses the same ceiling) x 148 SMs x clocks.max.sm; instruction "
                                    "count: %s" % (prof["alu_peak_inst_per_sm_cycle"], prof["alu_inst_source"]))
        else:
            roof.update(achieved=None, peak=None, unit="G ALU-pipe warp-inst/s", frac=None, traffic=None,
                        note2="no profiles/r*_step_ncu.json found")
        line = dict(
            metric=METRIC, value=round(value, 1), unit="reads/s",
            n_gpus=world, steps=args.steps, warmup=args.warmup, ms_per_step=round(total_ms / args.steps, 4),
            higher_is_better=True, scaling="weak", vs_baseline=None, dtype="int16 (packed x2, DPX)", data="synthetic",
            config=dict(workload=WORKLOAD, reads_per_gpu=READS_PER_SITE, read_len=READ_LEN, graph_cols=G_COLS,
                        cells_per_read=CELLS_PER_READ, parallelism="sites sharded, %d rank(s), no collective" % world,
                        l2="256 MiB flush write between timed steps; per-step scratch (checkpoints) exceeds L2",
                        library=capi.load().pg_version().decode()),
            e2e=dict(value=round(e2e_value, 1), unit="reads/s", h2d_bytes_per_step=h2d, d2h_bytes_per_step=d2h),
            gpu_launches=int(launches),
            kernels=dict(fill_ms=round(st["fill_ms"], 4), trace_ms=round(st["trace_ms"], 4),
                         count_ms=round(cst["count_ms"], 4),
                         note="fill / trace: run back to back on one stream (PG_SPLIT=1); the timed steps overlap two "
                              "half-batch pipelines on two streams, which is why ms_per_step is below their sum"),
            e2e_counts=dict(value=round(n_reads_all * args.steps / (cnt_ms * 1e-3), 1), unit="reads/s",
                            what="host reads -> filters + disambiguation + node/edge/path-family fragment counts on "
                                 "the device (pg_batch_count); only the count tables are copied back",
                            h2d_bytes_per_step=h2d + int(pairs.nbytes), d2h_bytes_per_step=cnt_d2h,
                            fragments=int(cnt["node_counts"][:, 0].max())),
            e2e_two_contexts=dict(value=round(2 * n_reads_all * args.steps / (two_ms * 1e-3), 1), unit="reads/s",
                                  what="pg_align_batch (host buffers in and out) from two host threads on two contexts "
                                       "per GPU: copies of one overlap the kernels of the other"),
            e2e_cascade=dict(value=round(n_reads_all * args.steps / (casc_ms * 1e-3), 1), unit="reads/s",
                             what="host reads -> exact-match stage (grm::PathAligner, k=32) + DP for the rest "
                                  "(pg_set_stages), records + CIGARs back; rank 0's stage counters and kernel times",
                             path_mapped=pst["mapped"], path_anchored=pst["anchored"], reads=READS_PER_SITE,
                             path_ms=round(pst["path_ms"], 4), fill_ms=round(cascade_kernels["fill_ms"], 4),
                             trace_ms=round(cascade_kernels["trace_ms"], 4)),
            roofline=roof,
            roofline_hbm=dict(bound="hbm", achieved=round(READS_PER_SITE * 200 / (fill_ms * 1e-3) / 1e9, 3) if fill_ms > 0 else None,
                              peak=peaks.get("hbm_gbs"), unit="GB/s",
                              frac=round(READS_PER_SITE * 200 / (fill_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 6)
                              if fill_ms > 0 and peaks.get("hbm_gbs") else None,
                              note="algorithmic traffic ~200 B/read (reads in, records + CIGAR out, SURVEY.md 8d): the path "
                                   "is not HBM-bound; the checkpoint scratch actually written is roofline.traffic"),
            clocks=clocks)
        if sweep is not None:
            line["sweep_config4"] = sweep
        if configs is not None:
            line["configs"] = configs
        if mirror is not None:
            line["e2e_mirror"] = mirror
        if base is not None:
            line["cpu_baseline"] = base
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
