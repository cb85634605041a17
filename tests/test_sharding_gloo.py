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
dist.barrier()
if rank == 0:
    json.dump({str(k): v for k, v in merged.items()}, open(sys.argv[2], "w"))
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
    merged = json.load(open(out))
    sites = synth.sites(seed=5, n_sites=6, max_reads=6)
    assert sorted(map(int, merged)) == list(range(6))
    for i, (_, nodes, edges, reads) in enumerate(sites):
        exp, _ = emubind.emu_align_batch(nodes, edges, reads)
        assert merged[str(i)] == exp
