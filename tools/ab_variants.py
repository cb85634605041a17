"""Build compile-time variants of libpgalign.so (tools/ab_variants.py build) and time them on config 2
(tools/ab_variants.py run, under gpurun).  Variants live in ab_build/ (git-ignored, travels with gpurun)."""
import os, subprocess, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
OUT = os.path.join(ROOT, "ab_build")
VARIANTS = {
    "base_u8": ["-DPG_SPEC_DEAD=0"],
    "spec_dead": ["-DPG_SPEC_DEAD=1"],  # speculative "no gap alive" blocks (pg_core.cuh: lane_step_dead)
    "spec_prune": ["-DPG_SPEC_DEAD=1", "-DPG_SPEC_PRUNE=1"],  # plus upper-bound pruning of gaps (gap_relevant)
    "spec_dead_4": ["-DPG_SPEC_DEAD=1", "-DPG_SPEC_STEPS=4"],
    "spec_dead_16": ["-DPG_SPEC_DEAD=1", "-DPG_SPEC_STEPS=16"],
}
def build():
    os.makedirs(OUT, exist_ok=True)
    src = os.path.join(ROOT, "paragraph_b200", "csrc", "pg_kernels.cu")
    procs = []
    for name, flags in VARIANTS.items():  # all at once: one nvcc is single-threaded for minutes on this file
        so = os.path.join(OUT, "libpg_%s.so" % name)
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-shared",
               "-Xcompiler", "-fPIC"] + flags + ["-o", so, src]
        procs.append((so, subprocess.Popen(cmd)))
    for so, p in procs:
        if p.wait() != 0:
            raise SystemExit("nvcc failed for " + so)
        print("built", so)
def run():
    code = r'''
import os, sys
sys.path.insert(0, %r)
from paragraph_b200 import capi, synth
nodes, edges, reads = synth.config2(seed=42, n_reads=10000)
ctx = capi.Context(0); ctx.add_graph(nodes, edges)
blob, off = ctx.pack_reads(reads)
for _ in range(3): ctx.align_packed(blob, off)
ts = []
for _ in range(6):
    ctx.align_packed(blob, off); s = ctx.stats(); ts.append((s["fill_ms"], s["trace_ms"]))
f = min(t[0] for t in ts); t = min(t[1] for t in ts)
import hashlib
rec, ops = ctx.align_packed(blob, off)
h = hashlib.sha1()
for x in rec:
    h.update(repr(tuple(int(x[f]) for f in ("graph_pos", "score", "unique", "chose_reverse", "status", "query_clipped"))).encode())
    h.update(ops[int(x["cigar_off"]):int(x["cigar_off"]) + int(x["cigar_len"])].tobytes())
print("%%-9s W=%%s fill %%.3f trace %%.3f total %%.3f ms  digest %%s" %% (os.environ["PG_VARIANT"], os.environ.get("PG_GEOM_W","32"), f, t, f+t, h.hexdigest()[:12]), flush=True)
''' % ROOT
    digests = set()
    for name in VARIANTS:
        for w in ["32"]:
            env = dict(os.environ, PG_LIB=os.path.join(OUT, "libpg_%s.so" % name), PG_VARIANT=name, PG_GEOM_W=w)
            out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            sys.stdout.write(out.stdout)
            sys.stderr.write(out.stderr[-2000:])
            for line in out.stdout.splitlines():
                if "digest" in line:
                    digests.add(line.split("digest")[1].strip())
    print("RESULT DIGESTS %s (%d distinct over %d variants)" % ("AGREE" if len(digests) == 1 else "DIFFER", len(digests), len(VARIANTS)))
if __name__ == "__main__":
    build() if sys.argv[1] == "build" else run()
