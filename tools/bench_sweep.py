"""Site sweep over the GPUs of one box (BASELINE.json configs[3] shape: DEL/INS/DUP/INV sites, 30x 150 bp reads, sites
sharded across ranks with no data-path collective; strong scaling: the sweep is fixed, ranks split it).

    python tools/bench_sweep.py [--sites 2000]                                             (1 GPU)
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_sweep.py --sites 2000

Every rank generates the same synthetic sweep, takes its LPT share (paragraph_b200.multigpu.partition_sites), aligns it in
one multi-site batch through the C-ABI (host buffers in and out, graphs registered inside the timed region) and the
per-site summaries are gathered on rank 0 (gather_object); time = max over ranks between two barriers."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch
import torch.distributed as dist
from paragraph_b200 import capi, multigpu, synth


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sites", type=int, default=2000)
    ap.add_argument("--cascade", type=int, default=1)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    sweep = synth.sites(seed=44, n_sites=a.sites, kinds=("DEL", "INS", "DUP", "INV"))
    costs = [multigpu.site_cost(len(r), 150, sum(len(s) for s in n)) for (_, n, e, r) in sweep]
    mine = multigpu.partition_sites(costs, world)[rank]
    ctx = capi.Context(local)
    if a.cascade:
        ctx.set_stages(32, True, True)
    reads, spans = [], []
    for i in mine:
        spans.append((len(reads), len(reads) + len(sweep[i][3])))
        reads += sweep[i][3]
    blob, off = ctx.pack_reads(reads, pinned=True)

    def one_pass():
        ctx.clear_graphs()
        sid = np.empty(len(reads), dtype=np.int32)
        for (lo, hi), i in zip(spans, mine):
            sid[lo:hi] = ctx.add_graph(sweep[i][1], sweep[i][2])
        rec, ops = ctx.align_packed(blob, off, sid)
        return {i: (int(rec["unique"][lo:hi].sum()), int((rec["mapped_by"][lo:hi] == 1).sum()), hi - lo)
                for (lo, hi), i in zip(spans, mine)}

    one_pass()
    times = []
    for _ in range(a.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        summary = one_pass()
        merged = multigpu.gather_site_results(summary, dist if world > 1 else None)
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        times.append(float(dt[0]))
    if rank == 0:
        n_reads = sum(v[2] for v in merged.values())
        best = min(times)
        print(json.dumps(dict(what="site sweep, sites sharded over ranks (LPT), graphs registered + aligned + gathered per pass",
                              n_gpus=world, sites=len(merged), reads=n_reads, cascade=bool(a.cascade),
                              seconds=round(best, 4), reads_per_s=round(n_reads / best, 1),
                              sites_per_s=round(len(merged) / best, 1),
                              by_exact_match_stage=sum(v[1] for v in merged.values()),
                              load_imbalance=round(max(sum(costs[i] for i in p) for p in multigpu.partition_sites(costs, world))
                                                   / (sum(costs) / world), 4))))
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
