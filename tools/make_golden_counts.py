#!/usr/bin/env python
"""Golden vectors for the disambiguation / read-counting stage.

Run in the build container (needs /root/reference).  Writes tests/golden/counts_*.json:

* counts_phasing.json -- a structural conversion of the reference's own expected output
  share/test-data/paragraph/phasing/expected.json (checked by src/python/test/test_phasing.py): the graph
  (node names + lengths, edges + sequence labels), every read's alignment (graphPos, graphCigar, length, strand,
  uniqueness, fragment id), the reference's verdict for it (MAPPED + graphNodesSupported / graphEdgesSupported /
  graphSequencesSupported, or the filter that removed it) and the site's read_counts_by_{node,edge,sequence}.
  Nothing is recomputed here: all expectations are the reference's committed numbers.
  (The pg-complex/ and quantification/ *.paragraph.json files in the same tree are NOT used: no reference test reads
  their alignments and they predate the current edge filter -- oracle/_ref disagrees with them.)
* counts_unit.json -- the graphs/reads/expectations of ParagraphTest.Aligns (src/c++/test/test_paragraph_parts.cpp:
  46-144) and DisambiguationTest (src/c++/test/test_disambiguation.cpp:45-105), transcribed; both call
  disambiguateReads with null filters.
"""
import json
import os
import re

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def node_len(n):
    # GraphInput.cpp:78-89: first/last node named source/sink becomes the 1-base sequence "X"
    if n["name"].upper() in ("SOURCE", "SINK"):
        return 1
    s = n.get("sequence") or n.get("reference_sequence")
    if s is not None:
        return len(s)
    m = re.match(r".*:(\d+)-(\d+)$", n["reference"])
    return int(m.group(2)) - int(m.group(1)) + 1


def phasing():
    j = json.load(open(os.path.join(REF, "share/test-data/paragraph/phasing/expected.json")))
    reads = []
    for a in j["alignments"]:
        st = a.get("graphMappingStatus")
        err = a.get("error", "")
        reads.append({
            "frag": a["fragmentId"], "len": len(a["bases"]), "pos": a.get("graphPos", 0), "cigar": a["graphCigar"],
            "rev": bool(a.get("isGraphReverseStrand", False)), "unique": bool(a.get("isGraphAlignmentUnique", False)),
            "score": a.get("graphAlignmentScore", 0),
            "verdict": "MAPPED" if st == "MAPPED" else err,
            "nodes": a.get("graphNodesSupported", []), "edges": a.get("graphEdgesSupported", []),
            "seqs": a.get("graphSequencesSupported", []),
        })
    doc = {
        "source": "share/test-data/paragraph/phasing/expected.json",
        "nodes": [{"name": n["name"], "len": node_len(n)} for n in j["nodes"]],
        "edges": [[e["from"], e["to"], e.get("sequences", [])] for e in j["edges"]],
        "reads": reads,
        "read_counts_by_node": j["read_counts_by_node"],
        "read_counts_by_edge": j["read_counts_by_edge"],
        "read_counts_by_sequence": j["read_counts_by_sequence"],
    }
    json.dump(doc, open(os.path.join(OUT, "counts_phasing.json"), "w"), separators=(",", ":"), sort_keys=True)
    print("counts_phasing.json:", len(reads), "reads,", len(doc["nodes"]), "nodes")


def unit():
    doc = {
        "ParagraphTest": {
            "source": "src/c++/test/test_paragraph_parts.cpp:46-144",
            "nodes": [["LF", "AAAAAAAAAAA"], ["P1", "TTTTTTTT"], ["Q1", "GGGGGGGG"], ["RF", "AAAAAAAAAAA"]],
            "edges": [["LF", "P1", ["P"]], ["LF", "Q1", ["Q"]], ["LF", "RF", ["D"]], ["P1", "RF", ["P"]],
                      ["Q1", "RF", ["Q"]]],
            "reads": [
                {"bases": "AAAAAAAATTTTCTTTAAAAAAAA", "pos": 3, "cigar": "0[8M]1[4M1X3M]3[8M]", "len": 24, "rev": False,
                 "nodes": ["LF", "P1", "RF"], "edges": ["LF_P1", "P1_RF"], "seqs": ["P"]},
                {"bases": "TTTTTTAAAGAAAATTTTTTT", "pos": 4, "cigar": "0[7M]1[4M1X3M]3[6M]", "len": 21, "rev": True,
                 "nodes": ["LF", "P1", "RF"], "edges": ["LF_P1", "P1_RF"], "seqs": ["P"]},
                {"bases": "AAAAAGCGGGGGGAAAAAA", "pos": 6, "cigar": "0[5M]2[1M1X6M]3[6M]", "len": 19, "rev": False,
                 "nodes": ["LF", "Q1", "RF"], "edges": ["LF_Q1", "Q1_RF"], "seqs": ["Q"]},
                {"bases": "AAAAGCGGGGGGAAAAAA", "pos": 7, "cigar": "0[4M]2[1M1X6M]3[6M]", "len": 18, "rev": False,
                 "nodes": ["LF", "Q1", "RF"], "edges": ["LF_Q1", "Q1_RF"], "seqs": ["Q"]},
                {"bases": "TTTTTTCCCCCCGCTTTTT", "pos": 6, "cigar": "0[5M]2[1M1X6M]3[6M]", "len": 19, "rev": True,
                 "nodes": ["LF", "Q1", "RF"], "edges": ["LF_Q1", "Q1_RF"], "seqs": ["Q"]},
                {"bases": "AAAAAAAAAAAAAAAAAAA", "pos": 0, "cigar": "0[11M]3[8M]", "len": 19, "rev": False,
                 "nodes": ["LF", "RF"], "edges": ["LF_RF"], "seqs": ["D"]},
            ],
        },
        "DisambiguationTest": {
            "source": "src/c++/test/test_disambiguation.cpp:45-105",
            "nodes": [["LF", "AAAAAAAAAA"], ["R1", "TTTTTTTTTT"], ["R2", "TTTTTTTTTT"], ["A1", "GGGGGGGGGG"],
                      ["RF", "AAAAAAAAAA"]],
            "edges": [["LF", "R1", ["R"]], ["LF", "RF", ["D"]], ["R1", "R2", ["R"]], ["R1", "A1", []],
                      ["R2", "RF", ["R"]], ["A1", "RF", []]],
            # reads are aligned by the test itself (gssw); only graphSequencesSupported is asserted
            "reads": [
                {"bases": "AAAAAAAAAATTTTTTTTTTTTTTTTTTTTAAAAAAAAAA", "seqs": ["R"]},
                {"bases": "AAAAAAAAAATTTTTTTTTTT", "seqs": ["R"]},
                {"bases": "AAAAAAAAAATTTTTTTTTTGGGGGGGGGGAAAAAAAAAA", "seqs": []},
                {"bases": "AAAAAAAAAAAAAAAAAAAA", "seqs": ["D"]},
            ],
        },
    }
    json.dump(doc, open(os.path.join(OUT, "counts_unit.json"), "w"), indent=1, sort_keys=True)
    print("counts_unit.json written")


if __name__ == "__main__":
    phasing()
    unit()
