"""Turn the reference's own graph JSONs (share/test-data/paragraph/...) into a committed fixture of graph SHAPES:
tests/golden/shapes_ref_graphs.json.  Run in the build container (needs /root/reference); the GPU box reads only
the fixture.

Loading follows grm::graphFromJson (src/c++/lib/grm/GraphInput.cpp:51-161): node order = JSON order (ids must ascend
along every edge, Graph::addEdge), a first / last node named source / sink becomes the 1-base sequence "X", a node
with a literal "sequence" keeps it.  A node given as "reference": "chr:start-end" needs the genome FASTA (HG19 /
HG38), which is not available here: its bases are synthesised -- one pseudo-random base per (chromosome, position),
so that nodes that overlap or abut on the genome agree with each other the way real reference nodes do.
"""
import glob
import hashlib
import json
import os
import re
import sys

REF = os.environ.get("PG_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "shapes_ref_graphs.json")
FILES = ["long-del/chr4-21369091-21376907.json", "pg-het-ins/pg-het-ins.json", "pg-complex/pg-complex.json",
         "pg-complex/pg-complex-2.json", "pg-complex/pg-complex-3.json", "pg-complex/pg-complex-3-using-symbolic-del.json",
         "haplo-complex/overlapping.json", "simple/del-example-3.json", "simple/del-example-4.json",
         "simple/swap-example-1.json", "simple/swap-example-2.json", "simple/swap-example-2-split.json",
         "simple/swap-example-5.json", "insertions/insertion-test-1.json", "phasing/long-phasing.json",
         "quantification/chr6-53037879-53037949.vcf.json"]


def base_at(chrom, pos):
    return "ACGT"[hashlib.sha1(("%s:%d" % (chrom, pos)).encode()).digest()[0] & 3]


def ref_seq(loc):
    m = re.match(r"^(.+):(\d+)-(\d+)$", loc)
    chrom, a, b = m.group(1), int(m.group(2)), int(m.group(3))
    return "".join(base_at(chrom, p) for p in range(a, b + 1))


def load(path):
    j = json.load(open(path))
    g = j.get("graph", j)
    names, seqs = [], []
    n = len(g["nodes"])
    for i, nd in enumerate(g["nodes"]):
        name = nd.get("name", "node-%d" % (i + 1))
        names.append(name)
        if (i == 0 or i == n - 1) and name.upper() in ("SOURCE", "SINK"):
            seqs.append("X")
        elif "sequence" in nd:
            seqs.append(nd["sequence"])
        else:
            r = nd["reference"]
            seqs.append(ref_seq(r if isinstance(r, str) else r[0]))
    idx = {nm: i for i, nm in enumerate(names)}
    edges = sorted({(idx[e["from"]], idx[e["to"]]) for e in g.get("edges", [])})
    if not all(f < t for f, t in edges):
        raise ValueError("node ids do not ascend along every edge")
    if any(len(s) == 0 for s in seqs):
        raise ValueError("empty node")
    return dict(nodes=seqs, edges=[list(e) for e in edges], node_names=names)


def main():
    out = []
    for f in FILES:
        p = os.path.join(REF, "share", "test-data", "paragraph", f)
        try:
            g = load(p)
        except Exception as e:  # say which files are unusable and why, do not hide them
            print("skipped %s: %s" % (f, e), file=sys.stderr)
            continue
        g["source"] = "share/test-data/paragraph/" + f
        out.append(g)
        print("%-55s %3d nodes %3d edges %6d bp" % (f, len(g["nodes"]), len(g["edges"]), sum(map(len, g["nodes"]))))
    with open(OUT, "w") as fh:
        json.dump(dict(generated_by="tools/make_golden_graphs.py", graphs=out), fh, indent=0)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
