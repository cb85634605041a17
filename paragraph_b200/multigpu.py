"""Site sharding across the GPUs of one box.

SV sites (graph + read batch) are independent (the reference hands (sample, graph) pairs to threads,
src/c++/lib/grmpy/Workflow.cpp:121-143), so the path shards without any data-path collective: one process
per GPU aligns its sites, and only the per-site results are gathered on the host (torch.distributed over whatever
backend the job runs -- NCCL on the GPU box, gloo in the CPU tests: gather_object for per-read results, one
fixed-size tensor gather for per-site integer summaries).
"""
import heapq


def site_cost(n_reads, read_len, graph_len):
    """DP cells of one site: 4 fills x L x G per read."""
    return 4 * int(n_reads) * int(read_len) * int(graph_len)


def partition_sites(costs, world_size):
    """Longest-processing-time-first greedy partition.  Returns world_size lists of site indices, each in
    ascending order; deterministic (ties broken by site index, then by rank)."""
    order = sorted(range(len(costs)), key=lambda i: (-costs[i], i))
    heap = [(0, r) for r in range(world_size)]
    heapq.heapify(heap)
    parts = [[] for _ in range(world_size)]
    for i in order:
        load, r = heapq.heappop(heap)
        parts[r].append(i)
        heapq.heappush(heap, (load + costs[i], r))
    return [sorted(p) for p in parts]


def split_reads(n_reads, world_size, rank):
    """Even contiguous split of one site's reads (single-site jobs replicate the KB-sized graph)."""
    base, rem = divmod(n_reads, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_site_results(local, dist=None, dst=0):
    """local: dict site_index -> result object for the sites this rank aligned.  Returns, on rank dst, the
    merged dict over all ranks (input order is recovered by sorting the keys); None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return dict(local)
    world = dist.get_world_size()
    bucket = [None] * world if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    merged = {}
    for part in bucket:
        for k, v in part.items():
            if k in merged:
                raise RuntimeError("site %r aligned by two ranks" % (k,))
            merged[k] = v
    return merged


def gather_site_summaries(site_ids, columns, cap, dist=None, device="cpu", dst=0):
    """Per-site integer summaries (e.g. unique reads, score sums, read counts) of this rank's sites to rank dst as ONE
    fixed-size tensor gather: row 0 = site ids, rows 1.. = `columns` (equally long int arrays), padded to `cap` sites (the
    largest shard; every rank knows it from partition_sites), last column = number of sites.  Returns, on rank dst, a list
    over ranks of (site_ids, [column arrays]); None elsewhere.  (gather_object pickles and needs two collectives; for a
    0.1 s job on 8 GPUs that was a third of what does not shrink with N.)"""
    import numpy as np
    import torch
    k = len(site_ids)
    if k > cap:
        raise ValueError("shard of %d sites exceeds cap %d" % (k, cap))
    pack = np.zeros((1 + len(columns), cap + 1), dtype=np.int64)
    pack[0, :k] = site_ids
    for r, col in enumerate(columns):
        pack[1 + r, :k] = col
    pack[0, cap] = k

    def unpack(a):
        kk = int(a[0, cap])
        return a[0, :kk].astype(np.int32), [a[1 + r, :kk].copy() for r in range(len(columns))]

    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [unpack(pack)]
    mine = torch.from_numpy(pack).to(device)
    bucket = [torch.empty_like(mine) for _ in range(dist.get_world_size())] if dist.get_rank() == dst else None
    dist.gather(mine, bucket, dst=dst)
    if dist.get_rank() != dst:
        return None
    return [unpack(t.cpu().numpy()) for t in bucket]


def align_sites(ctx, sites, my_sites, flags=0xFFFFFFFF):
    """Align this rank's share of `sites` (list of (nodes, edges, reads)) in ONE multi-site batch on `ctx`
    (a paragraph_b200.capi.Context).  Returns dict site_index -> list of per-read result dicts."""
    ctx.clear_graphs()
    reads, site_ids, spans = [], [], {}
    for i in my_sites:
        nodes, edges, rds = sites[i]
        sid = ctx.add_graph(nodes, edges)
        spans[i] = (len(reads), len(reads) + len(rds))
        reads += rds
        site_ids += [sid] * len(rds)
    if not reads:
        return {}
    res = ctx.align(reads, sites=site_ids, flags=flags)
    return {i: res[a:b] for i, (a, b) in spans.items()}
