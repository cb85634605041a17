"""A/B of PG_SPLIT (chunks per batch for the fill / traceback overlap) on config 2 through pg_align_batch (host buffers,
wall clock over 30 calls) and through pg_batch_run (CUDA events).  Run under gpurun."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
code = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
blob, off = ctx.pack_reads(reads, pinned=True)
for _ in range(5): ctx.align_packed(blob, off)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): ctx.align_packed(blob, off)
torch.cuda.synchronize(); e2e = (time.perf_counter() - t0) / 30
ctx.upload(blob, off)
for _ in range(3): ctx.run()
ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(20)]
for a, b in ev:
    a.record(stream); ctx.run(); b.record(stream)
torch.cuda.synchronize()
k = sum(a.elapsed_time(b) for a, b in ev) / 20
ctx.set_stages(32, True, True)
for _ in range(3): ctx.align_packed(blob, off)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(30): ctx.align_packed(blob, off)
torch.cuda.synchronize(); casc = (time.perf_counter() - t0) / 30
print("PG_SPLIT=%%s kernels %%.3f ms  e2e %%.3f ms (%%.2f Mreads/s)  cascade e2e %%.3f ms (%%.2f Mreads/s)" %% (os.environ.get("PG_SPLIT", "default"), k, e2e * 1e3, 10000 / e2e / 1e6, casc * 1e3, 10000 / casc / 1e6), flush=True)
''' % ROOT
for s in sys.argv[1:] or ["1", "2", "3", "4", "6"]:
    subprocess.run([sys.executable, "-c", code], env=dict(os.environ, PG_SPLIT=s))
