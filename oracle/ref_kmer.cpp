// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI driver around the UNMODIFIED reference KmerAligner
// (src/c++/lib/grm/KmerAligner.cpp), compiled where it lies by oracle/Makefile into oracle/_ref/libpgref.so.
// KmerAligner = the second stage of grm::CompositeAligner (lib/grm/CompositeAligner.cpp:105-126): gapless alignment of
// the read to the sequence of one of the graph's PATHS, seeded by shared k-mers (16 in the product, 10 in the
// reference's unit test), at most two mismatches, unique unless a second candidate does as well (pickBest).
// The only stand-in is oligo/KmerGenerator.hh (oracle/ref_shim/oligo: the reference's version needs Boost.MPL).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <list>
#include <string>
#include <vector>

#include "grm/KmerAligner.hh"
#include "graphcore/Graph.hh"
#include "graphcore/Path.hh"

using graphtools::Graph;

namespace
{
template <unsigned K>
void run(
    const Graph& g, const std::list<graphtools::Path>& paths, int n_reads, const char* bases_blob, const int32_t* read_off,
    const uint8_t* is_rev, int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride, int32_t* counters2)
{
    grm::KmerAligner<K> al;
    al.setGraph(&g, paths);
    for (int i = 0; i < n_reads; ++i)
    {
        const int len = read_off[i + 1] - read_off[i];
        common::Read r;
        r.set_bases(std::string(bases_blob + read_off[i], bases_blob + read_off[i + 1]));
        r.set_quals(std::string(static_cast<size_t>(len), '#'));
        r.set_is_reverse_strand(is_rev && is_rev[i]);
        al.alignRead(r);
        int32_t* o = out8 + 8 * i;
        // 0 UNMAPPED, 1 MAPPED, 2 BAD_ALIGN (= a second candidate as good as the best one: not unique)
        o[0] = r.graph_mapping_status() == common::Read::MAPPED ? 1 : (r.graph_mapping_status() == common::Read::BAD_ALIGN ? 2 : 0);
        o[1] = r.graph_pos();
        o[2] = r.graph_alignment_score();
        o[3] = r.is_graph_alignment_unique() ? 1 : 0;
        o[4] = r.graph_mapq();
        o[5] = r.is_graph_reverse_strand() ? 1 : 0;
        o[6] = static_cast<int32_t>(r.graph_cigar().size());
        o[7] = 0;
        if (out_bases_blob)
            memcpy(out_bases_blob + read_off[i], r.bases().data(), static_cast<size_t>(len));
        if (cigars && cigar_stride > 0)
        {
            size_t n = std::min(static_cast<size_t>(cigar_stride - 1), r.graph_cigar().size());
            memcpy(cigars + static_cast<size_t>(i) * cigar_stride, r.graph_cigar().data(), n);
            cigars[static_cast<size_t>(i) * cigar_stride + n] = 0;
        }
    }
    counters2[0] = static_cast<int32_t>(al.attempted());
    counters2[1] = static_cast<int32_t>(al.mapped());
}
}

extern "C" {

// paths: path p = the nodes path_nodes[path_ptr[p] .. path_ptr[p+1]) from the first base of its first node to the last
// base of its last node, like grm::pathsFromJson builds them (lib/grm/GraphInput.cpp:168-197).
// kmer_len: 16 (CompositeAligner) or 10 (the reference's unit test) -- the two instantiations the reference has.
// out8 per read = {status, graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, 0}
int pgref_kmer_align_batch(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto,
    int n_paths, const int32_t* path_ptr, const int32_t* path_nodes, int kmer_len, int n_reads, const char* bases_blob,
    const int32_t* read_off, const uint8_t* is_rev, int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride,
    int32_t* counters2)
{
    try
    {
        Graph g(static_cast<size_t>(n_nodes), false);
        for (int i = 0; i < n_nodes; ++i)
        {
            g.setNodeName(i, "n" + std::to_string(i));
            g.setNodeSeq(i, std::string(seq_blob + seq_off[i], seq_blob + seq_off[i + 1]));
        }
        for (int e = 0; e < n_edges; ++e)
            g.addEdge(efrom[e], eto[e]);
        std::list<graphtools::Path> paths;
        for (int p = 0; p < n_paths; ++p)
        {
            std::vector<graphtools::NodeId> nodes(path_nodes + path_ptr[p], path_nodes + path_ptr[p + 1]);
            paths.emplace_back(&g, 0, nodes, g.nodeSeq(nodes.back()).size() - 1);
        }
        if (kmer_len == 16)
            run<16>(g, paths, n_reads, bases_blob, read_off, is_rev, out8, out_bases_blob, cigars, cigar_stride, counters2);
        else if (kmer_len == 10)
            run<10>(g, paths, n_reads, bases_blob, read_off, is_rev, out8, out_bases_blob, cigars, cigar_stride, counters2);
        else
            return -2;
        return 0;
    }
    catch (std::exception const&)
    {
        return -1;
    }
}
}
