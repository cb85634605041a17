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
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from paragraph_b200 import synth  # noqa: E402

READS_PER_SITE = 10000
READ_LEN = 150
FLANK, DEL_LEN = 500, 300
G_COLS = 2 * FLANK + DEL_LEN
CELLS_PER_READ = 4 * READ_LEN * G_COLS  # 4 fills x L x G = 780 000
DPX_SLOTS_PER_CELL_PAIR = 5.0           # 4 half-rate DPX ops + 2 full-rate VIMNMX per packed pair of cells
WORKLOAD = "config2: 3-node DEL graph (500 bp flanks, D=300), 10k synthetic 150 bp reads per GPU"


def ncu_traffic():
    """dram bytes read + written per launch of the dominant kernel, from the committed ncu --set full capture."""
    path = os.path.join(ROOT, "profiles", "r01j_fill_kernel_ncu.txt")
    try:
        tot, scale = 0.0, {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(path):
            f = line.split()
            if f and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1]) * scale[f[2]]
        return dict(dram_bytes_per_launch=int(tot), source="profiles/r01j_fill_kernel_ncu.txt (ncu --set full of the same workload, tools/profile_run.py)")
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


def cpu_reference(nodes, edges, reads, budget_s=12.0):
    """The reference's own CPU implementation (oracle/_ref = unmodified gssw.c + GraphAligner.cpp, -O3 -msse4.1)
    on this box's host cores, one GraphAligner per thread over contiguous chunks like grm::alignReads
    (Align.cpp:107-153), on a bounded sample.  Thread counts are swept and the best one reported (the reference's
    per-fill allocations make it scale poorly past a few dozen threads)."""
    from oracle import refbind
    kind = "reference" if refbind.have_ref() else "port"
    ncpu = os.cpu_count() or 1
    best = None

    def run(sample, threads):
        t0 = time.perf_counter()
        if kind == "reference":
            refbind.ref_align_batch(nodes, edges, sample, threads=threads)
        else:
            refbind.OracleGraph(nodes, edges).align_batch(sample)
        return len(sample) / (time.perf_counter() - t0)

    run(reads[:64], min(8, ncpu))  # warm
    cands = sorted({ncpu, max(1, ncpu // 2), max(1, ncpu // 4), min(ncpu, 32), min(ncpu, 16), min(ncpu, 8)}) \
        if kind == "reference" else [1]
    t_start = time.perf_counter()
    for t in cands:
        probe = run(reads[:min(len(reads), 64 * t)], t)
        n = min(len(reads), max(64 * t, int(probe * 1.5)))  # ~1.5 s of wall time per candidate
        rate = run(reads[:n], t)
        if best is None or rate > best[0]:
            best = (rate, t, n)
        if time.perf_counter() - t_start > budget_s:
            break
    return dict(value=round(best[0], 1), unit="reads/s", cores=best[1], kind=kind,
                sample="%d reads of the config-2 batch, best of thread counts %s (host has %d hardware threads)"
                       % (best[2], cands, ncpu))


def bench_reference(args):
    """--impl reference: the reference CPU path on the host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    nodes, edges, reads = workload(0)
    from oracle import refbind
    kind = "reference" if refbind.have_ref() else "port"
    ncpu = os.cpu_count() or 1
    base = cpu_reference(nodes, edges, reads, budget_s=8.0)
    threads, per_step = base["cores"], max(512, min(len(reads), int(base["value"] * 1.5)))
    sample = reads[:per_step]

    def step():
        if kind == "reference":
            refbind.ref_align_batch(nodes, edges, sample, threads=threads)
        else:
            refbind.OracleGraph(nodes, edges).align_batch(sample)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    value = per_step * args.steps / dt
    line = dict(metric="reads/sec to graph (150bp, DEL/INS); bit-exact score+CIGAR", value=round(value, 1),
                unit="reads/s", n_gpus=args.gpus, steps=args.steps, warmup=args.warmup,
                ms_per_step=round(dt / args.steps * 1e3, 3), higher_is_better=True, scaling="weak", vs_baseline=None,
                dtype="u8", data="synthetic", impl="reference",
                config=dict(workload=WORKLOAD, step_sample="%d reads per step" % per_step, threads=threads),
                cpu_baseline=dict(value=round(value, 1), unit="reads/s", cores=threads, kind=kind,
                                  sample="%d reads per step x %d steps (host has %d hardware threads)"
                                         % (per_step, args.steps, ncpu)),
                e2e=dict(value=round(value, 1), unit="reads/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return bench_reference(args)

    import torch
    import torch.distributed as dist
    from paragraph_b200 import capi

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; paragraph_b200 has no CPU fallback (use --impl reference "
                         "for the CPU arm)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    nodes, edges, reads = workload(rank)
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
        # dominant kernel = pg_fill_kernel: DPX-issue roofline.  Peak lane-op rate measured with
        # tools/ubench/dpx_ubench.cu on this pool: 63.2 packed-int16 DPX lane-ops / clk / SM.
        sm_max = (clocks.get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        dpx_peak = 63.2 * 148 * sm_max * 1e6                     # lane-ops/s
        peak_cells = dpx_peak * 2.0 / DPX_SLOTS_PER_CELL_PAIR   # cell updates/s
        fill_ms = st["fill_ms"]
        ach_cells = READS_PER_SITE * CELLS_PER_READ / (fill_ms * 1e-3) if fill_ms > 0 else 0.0
        line = dict(
            metric="reads/sec to graph (150bp, DEL/INS); bit-exact score+CIGAR", value=round(value, 1), unit="reads/s",
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
            roofline=dict(bound="alu",
                          note="packed-int16 DPX issue rate of the ALU pipe: the max-plus recurrence is neither HBM-bound "
                               "(roofline_hbm) nor a tensor-core contraction (SURVEY.md 8d, DESIGN.md 4)",
                          kernel="pg_fill_kernel<5,32> (fill phase of a step: forward-graph launch + paired reversed-graph launches)", achieved=round(ach_cells / 1e9, 1), peak=round(peak_cells / 1e9, 1),
                          unit="Gcell/s", frac=round(ach_cells / peak_cells, 4) if peak_cells else None,
                          traffic=ncu_traffic(),
                          peak_source="tools/ubench/dpx_ubench.cu on this pool: 63.2 DPX lane-ops/clk/SM x 148 SM x "
                                      "clocks.max.sm, 2 cells per lane-op, 5 issue slots per cell pair"),
            roofline_hbm=dict(bound="hbm", achieved=round(READS_PER_SITE * 200 / (fill_ms * 1e-3) / 1e9, 3) if fill_ms > 0 else None,
                              peak=peaks.get("hbm_gbs"), unit="GB/s",
                              frac=round(READS_PER_SITE * 200 / (fill_ms * 1e-3) / 1e9 / peaks["hbm_gbs"], 6)
                              if fill_ms > 0 and peaks.get("hbm_gbs") else None,
                              note="algorithmic traffic ~200 B/read (reads in, records + CIGAR out, SURVEY.md 8d): the path "
                                   "is not HBM-bound; the checkpoint scratch actually written is roofline.traffic"),
            clocks=clocks)
        if not args.no_cpu_baseline and world == 1:
            try:
                line["cpu_baseline"] = cpu_reference(nodes, edges, reads)
            except Exception as e:  # the baseline is reported, never required for the GPU number
                line["cpu_baseline"] = dict(value=None, error=str(e))
        elif not args.no_cpu_baseline:
            line["cpu_baseline"] = dict(value=None, note="measured at N=1 only")
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
