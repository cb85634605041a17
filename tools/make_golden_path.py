"""Generate tests/golden/path_aligner.json from the UNMODIFIED reference PathAligner (oracle/_ref): the unit-test graph
of src/c++/test/test_paragraph_parts.cpp, a DEL site, an INS site and fuzz cases.  Run in the build container
(/root/reference present)."""
import json, os, sys
sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "..", "tests"))
import numpy as np
from oracle import refbind as R
from paragraph_b200 import synth
from test_path_oracle import path_cases

def case(name, nodes, edges, reads, k):
    exp, cnt = R.ref_path_align_batch(nodes, edges, reads, kmer_len=k)
    return dict(name=name, nodes=nodes, edges=[list(e) for e in edges], reads=reads, k=k, expected=exp, counters=list(cnt))

def main():
    rng = np.random.default_rng(2024)
    cases = []
    gnodes = ["AAAAAAAAAAA", "TTTTTTTT", "GGGGGGGG", "AAAAAAAAAAA"]
    gedges = [(0, 1), (0, 2), (0, 3), (1, 3), (2, 3)]
    greads = ["AAAAAAAATTTTCTTTAAAAAAAA", "TTTTTTAAAGAAAATTTTTTT", "AAAAAGCGGGGGGAAAAAA", "AAAAGCGGGGGGAAAAAA",
              "TTTTTTCCCCCCGCTTTTT", "AAAAAAAAAAAAAAAAAAA", "AAAATTTTTTTTAAAA", "AAAAGGGGGGGGAAAA", "TTTTCCCCCCCCTTTT"]
    for k in (4, 8, 12):
        cases.append(case("ref_unit_graph_k%d" % k, gnodes, gedges, greads, k))
    nodes, edges = synth.del_graph(rng, 200, 80)
    reads = synth.simulate_reads(rng, nodes, edges, 60, read_len=100, sub=0.004, indel_frac=0.0)
    cases.append(case("del_site_k32", nodes, edges, reads, 32))
    nodes, edges = synth.ins_graph(rng, 150, 60) if hasattr(synth, "ins_graph") else synth.del_graph(rng, 150, 60)
    reads = synth.simulate_reads(rng, nodes, edges, 40, read_len=90, sub=0.005, indel_frac=0.0)
    cases.append(case("site2_k16", nodes, edges, reads, 16))
    for i, (nodes, edges, reads, k) in enumerate(path_cases(rng, 12, 8)):
        cases.append(case("fuzz_%d" % i, nodes, edges, reads, k))
    out = os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "path_aligner.json")
    with open(out, "w") as f:
        json.dump(dict(source="oracle/_ref (unmodified src/c++/lib/grm/PathAligner.cpp + graph-tools KmerIndex.cpp)", cases=cases), f)
    print("wrote", out, len(cases), "cases", sum(len(c["reads"]) for c in cases), "reads",
          sum(e["mapped"] for c in cases for e in c["expected"]), "mapped")

if __name__ == "__main__":
    main()
