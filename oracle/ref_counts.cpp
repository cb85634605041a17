// TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
//
// Reference arm for SURVEY.md 8f rank 1 (disambiguation + read counting).  Linked into oracle/_ref/libpgref.so
// together with the UNMODIFIED reference sources
//     src/c++/lib/paragraph/ReadCounting.cpp   (countReads / countNodes / countEdges / countPathFamilies)
//     src/c++/lib/common/Fragment.cpp          (readsToFragments / Fragment::addRead)
//     graph-tools: decodeGraphAlignment, Alignment statistics, PathFamily::containsPath, GraphCoordinates
// What cannot be linked is src/c++/lib/paragraph/Disambiguation.cpp (its translation unit pulls in BamReader,
// GraphInput, KmerAligner -> Boost MPL / htslib, absent here).  The two filter lambdas of alignAndDisambiguate
// (Disambiguation.cpp:212-282) and the loop of disambiguateReads (Disambiguation.cpp:82-142) are therefore
// restated below, operating on the reference's own GraphAlignment / PathFamily objects.  They are pinned by the
// reference's golden vectors: ParagraphTest.Aligns (test_paragraph_parts.cpp:113-144), DisambiguationTest
// (test_disambiguation.cpp:97-105) and the per-read graphNodesSupported/graphEdgesSupported/
// graphSequencesSupported + read_counts_by_* of share/test-data/paragraph/{pg-complex,quantification,phasing}
// (tests/test_counts_oracle.py).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <vector>

#include "common/Fragment.hh"
#include "common/Read.hh"
#include "graphalign/GraphAlignmentOperations.hh"
#include "graphcore/Graph.hh"
#include "graphcore/GraphCoordinates.hh"
#include "graphcore/PathFamily.hh"
#include "paragraph/ReadCounting.hh"

using graphtools::Graph;
using graphtools::GraphAlignment;
using graphtools::NodeId;

namespace
{
// Disambiguation.cpp:212-243
bool nodeSupported(Graph const& graph, common::Read const& read, NodeId node_id)
{
    try
    {
        GraphAlignment aln = graphtools::decodeGraphAlignment(read.graph_pos(), read.graph_cigar(), &graph);
        const size_t half = read.bases().size() / 2;
        const bool short_node = graph.nodeSeq(node_id).size() < half;
        int32_t k = 0;
        for (auto const& na : aln)
        {
            if (aln.getNodeIdByIndex(k) == node_id)
            {
                const size_t nonmatch = na.numMismatched() + na.numClipped();
                const size_t indel = na.numInserted() + na.numDeleted();
                if (short_node && (nonmatch > 0 || indel > 0))
                    return false;
                return nonmatch + indel <= half;
            }
            ++k;
        }
    }
    catch (std::exception const&)
    {
    }
    return false;
}

// Disambiguation.cpp:245-296 (DISABLE_ADDITIONAL_EDGE_FILTER is not defined in the reference build)
bool edgeSupported(Graph const& graph, common::Read const& read, NodeId n1, NodeId n2)
{
    try
    {
        GraphAlignment aln = graphtools::decodeGraphAlignment(read.graph_pos(), read.graph_cigar(), &graph);
        const graphtools::Alignment* prev = nullptr;
        NodeId prev_id = static_cast<NodeId>(-1);
        int32_t k = 0;
        for (auto const& na : aln)
        {
            const NodeId id = aln.getNodeIdByIndex(k);
            if (prev != nullptr && prev_id == n1 && id == n2)
            {
                const int32_t min_overlap = static_cast<int32_t>(read.bases().length() / 10 + 1);
                bool ok = prev->numMatched() >= (unsigned)std::min(prev->referenceLength(), (uint32_t)min_overlap)
                    && na.numMatched() >= (unsigned)std::min(na.referenceLength(), (uint32_t)min_overlap);
                if (ok)
                    ok = (prev->queryLength() < prev->referenceLength() * 2) && (na.queryLength() < na.referenceLength() * 2);
                if (ok)
                {
                    const int32_t l1 = static_cast<int32_t>(graph.nodeSeq(n1).size());
                    const int32_t l2 = static_cast<int32_t>(graph.nodeSeq(n2).size());
                    ok = ((int32_t)prev->numMatched() >= std::min(l1, min_overlap))
                        && ((int32_t)na.numMatched() >= std::min(l2, min_overlap));
                }
                return ok;
            }
            prev = &na;
            prev_id = id;
            ++k;
        }
    }
    catch (std::exception const&)
    {
    }
    return false;
}

// Disambiguation.cpp:82-142; use_filters == 0 reproduces the nullptr-filter call of the reference's unit tests.
void disambiguate(Graph* g, std::vector<common::p_Read>& reads, bool use_filters)
{
    for (auto& read : reads)
    {
        read->clear_graph_sequences_supported();
        read->clear_graph_nodes_supported();
        read->clear_graph_edges_supported();
        if (read->graph_mapping_status() != common::Read::MAPPED)
            continue;
        std::set<std::pair<std::string, std::string>> edges;
        std::set<NodeId> nodes;
        std::set<std::string> families;
        GraphAlignment gm = graphtools::decodeGraphAlignment(read->graph_pos(), read->graph_cigar(), g);
        auto const& path = gm.path();
        bool have_prev = false;
        NodeId pnode = 0;
        for (auto it = path.begin(); it != path.end(); ++it)
        {
            if (have_prev && (!use_filters || edgeSupported(*g, *read, pnode, *it)))
            {
                edges.emplace(g->nodeName(pnode), g->nodeName(*it));
                for (auto const& label : g->edgeLabels(pnode, *it))
                    families.insert(label);
            }
            have_prev = true;
            pnode = *it;
            if (!use_filters || nodeSupported(*g, *read, *it))
                nodes.emplace(*it);
        }
        for (auto n : nodes)
            read->add_graph_nodes_supported(g->nodeName(n));
        for (auto const& e : edges)
            read->add_graph_edges_supported(e.first + "_" + e.second);
        for (auto const& label : families)
        {
            graphtools::PathFamily fam(g, label);
            if (fam.containsPath(path))
                read->add_graph_sequences_supported(label);
        }
    }
}
}

extern "C" {

// One site.  Reads passed in must be the ones alignReads kept (status MAPPED, Align.cpp:81-84).  Node i is named
// "n<i>", label bit k is named "L<k>".  Writes a JSON document
//   {"reads":[{"nodes":[..],"edges":[..],"sequences":[..]},..],
//    "read_counts_by_node":{..},"read_counts_by_edge":{..},"read_counts_by_sequence":{..}}
// into out (NUL-terminated).  Returns the JSON length, -1 if out is too small, -2 if the reference threw
// (a CIGAR decodeGraphAlignment rejects aborts the whole site there: Disambiguation.cpp:102).
int pgref_count_site(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto,
    const uint64_t* edge_labels, int n_reads, const int32_t* read_len, const int32_t* graph_pos, const char* cigars,
    int cigar_stride, const uint8_t* is_graph_reverse, const int32_t* fragment, int use_filters, int detailed,
    char* out, int out_cap)
{
    try
    {
        Graph graph(static_cast<size_t>(n_nodes), false);
        for (int i = 0; i < n_nodes; ++i)
        {
            graph.setNodeName(i, "n" + std::to_string(i));
            graph.setNodeSeq(i, std::string(seq_blob + seq_off[i], seq_blob + seq_off[i + 1]));
        }
        for (int e = 0; e < n_edges; ++e)
        {
            graph.addEdge(efrom[e], eto[e]);
            for (int k = 0; edge_labels && k < 64; ++k)
                if ((edge_labels[e] >> k) & 1)
                    graph.addLabelToEdge(efrom[e], eto[e], "L" + std::to_string(k));
        }
        std::vector<common::p_Read> reads;
        for (int i = 0; i < n_reads; ++i)
        {
            common::p_Read r(new common::Read());
            r->set_fragment_id("f" + std::to_string(fragment ? fragment[i] : i));
            r->set_bases(std::string(static_cast<size_t>(read_len[i]), 'A'));
            r->set_quals(std::string(static_cast<size_t>(read_len[i]), '#'));
            r->set_graph_pos(graph_pos[i]);
            r->set_graph_cigar(std::string(cigars + static_cast<size_t>(i) * cigar_stride));
            r->set_is_graph_reverse_strand(is_graph_reverse && is_graph_reverse[i]);
            r->set_graph_mapping_status(common::Read::MAPPED);
            reads.push_back(std::move(r));
        }
        disambiguate(&graph, reads, use_filters != 0);

        Json::Value doc = Json::objectValue;
        doc["reads"] = Json::arrayValue;
        for (auto const& r : reads)
        {
            Json::Value jr = Json::objectValue;
            jr["nodes"] = Json::arrayValue;
            jr["edges"] = Json::arrayValue;
            jr["sequences"] = Json::arrayValue;
            for (auto const& s : r->graph_nodes_supported())
                jr["nodes"].append(s);
            for (auto const& s : r->graph_edges_supported())
                jr["edges"].append(s);
            for (auto const& s : r->graph_sequences_supported())
                jr["sequences"].append(s);
            doc["reads"].append(jr);
        }
        graphtools::GraphCoordinates coordinates(&graph);
        paragraph::countReads(coordinates, reads, doc, true, true, true, detailed != 0);
        doc.removeMember("fragment_statistics"); // computed by the Boost stand-in, not a parity item

        Json::FastWriter w;
        const std::string s = w.write(doc);
        if (static_cast<int>(s.size()) + 1 > out_cap)
            return -1;
        memcpy(out, s.c_str(), s.size() + 1);
        return static_cast<int>(s.size());
    }
    catch (std::exception const& e)
    {
        if (out_cap > 0)
        {
            strncpy(out, e.what(), static_cast<size_t>(out_cap) - 1);
            out[out_cap - 1] = 0;
        }
        return -2;
    }
}
}
