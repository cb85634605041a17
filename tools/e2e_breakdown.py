"""Where does the end-to-end time of one config-2 batch go?  (run under gpurun)"""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import numpy as np, torch
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0)
ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads)
for _ in range(3):
    ctx.align_packed(blob, off)
def t(f, n=20):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
print("align_packed (one call)      %.3f ms" % t(lambda: ctx.align_packed(blob, off)))
print("upload only                  %.3f ms" % t(lambda: ctx.upload(blob, off)))
ctx.upload(blob, off)
def run_sync():
    ctx.run(); torch.cuda.synchronize()
print("run + device sync            %.3f ms" % t(run_sync))
ctx.run()
print("download only                %.3f ms" % t(lambda: ctx.download()))
print("np.zeros(3.2M u32)           %.3f ms" % t(lambda: np.zeros(10000*320, dtype=np.uint32)))
print(ctx.stats())
