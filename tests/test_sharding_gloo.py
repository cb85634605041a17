"""N > 1 host path on CPU: world_size-2 gloo job.  Each rank 'aligns' its LPT share of the sites (with the CPU
lane emulator standing in for the GPU engine) and rank 0 gathers the per-site results; the merged result must
equal the single-process one, site for site.  No collective touches the data path."""
import json
import os
import subprocess
import sys

import numpy as np

from conftest import ROOT
from paragraph_b200 import multigpu, synth

WORKER = r'''
import json, os, sys
sys.path.insert(0, sys.argv[1]); sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
import torch.distributed as dist
from paragraph_b200 import multigpu, synth
import emubind
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
sites = [(k, n, e, r) for (k, n, e, r) in synth.sites(seed=5, n_sites=6, max_reads=6)]
costs = [multigpu.site_cost(len(r), 150, sum(map(len, n))) for (_, n, _, r) in sites]
mine = multigpu.partition_sites(costs, world)[rank]
local = {}
for i in mine:
    _, nodes, edges, reads = sites[i]
    res, _ = emubind.emu_align_batch(nodes, edges, reads)
    local[i] = res
merged = multigpu.gather_site_results(local, dist, dst=0)
# the per-site integer summaries of bench.py's sweep leg: one fixed-size tensor gather
import numpy as np
shards = multigpu.partition_sites(costs, world)
cap = max(len(p) for p in shards)
uniq = [sum(int(x["unique"]) for x in local[i]) for i in mine]
score = [sum(int(x["score"]) for x in local[i]) for i in mine]
parts = multigpu.gather_site_summaries(np.asarray(mine, dtype=np.int32), [uniq, score], cap, dist, device="cpu", dst=0)
dist.barrier()
if rank == 0:
    summ = {}
    for ids, cols in parts:
        for j, i in enumerate(ids):
            summ[str(int(i))] = [int(cols[0][j]), int(cols[1][j])]
    json.dump({"results": {str(k): v for k, v in merged.items()}, "summaries": summ}, open(sys.argv[2], "w"))
else:
    assert parts is None
dist.destroy_process_group()
'''


def test_partition_is_balanced_and_complete():
    rng = np.random.default_rng(0)
    costs = [int(c) for c in rng.integers(1, 1000, size=97)]
    for world in (1, 2, 4, 8):
        parts = multigpu.partition_sites(costs, world)
        assert sorted(sum(parts, [])) == list(range(97))
        loads = [sum(costs[i] for i in p) for p in parts]
        assert max(loads) <= sum(costs) / world + max(costs)
        assert parts == multigpu.partition_sites(costs, world)  # deterministic
    assert [multigpu.split_reads(10, 4, r) for r in range(4)] == [(0, 3), (3, 6), (6, 8), (8, 10)]


def test_two_rank_gloo_gather_matches_single_process(built, tmp_path):
    import emubind
    emubind.build()
    out = tmp_path / "merged.json"
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                    "--master-addr", "127.0.0.1", "--master-port", "29531", str(script), ROOT, str(out)],
                   check=True, env=env, timeout=600)
    doc = json.load(open(out))
    merged, summ = doc["results"], doc["summaries"]
    sites = synth.sites(seed=5, n_sites=6, max_reads=6)
    assert sorted(map(int, merged)) == list(range(6)) and sorted(map(int, summ)) == list(range(6))
    for i, (_, nodes, edges, reads) in enumerate(sites):
        exp, _ = emubind.emu_align_batch(nodes, edges, reads)
        assert merged[str(i)] == exp
        # the tensor gather of the per-site summaries (multigpu.gather_site_summaries) agrees with the per-read results
        assert summ[str(i)] == [sum(int(x["unique"]) for x in exp), sum(int(x["score"]) for x in exp)]


def test_site_summaries_single_process():
    parts = multigpu.gather_site_summaries(np.array([4, 7, 9], dtype=np.int32), [[1, 2, 3], [10, 20, 30]], cap=5)
    assert len(parts) == 1 and parts[0][0].tolist() == [4, 7, 9]
    assert [c.tolist() for c in parts[0][1]] == [[1, 2, 3], [10, 20, 30]]
