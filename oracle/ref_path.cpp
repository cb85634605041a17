// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI driver around the UNMODIFIED reference PathAligner
// (src/c++/lib/grm/PathAligner.cpp) and graph-tools KmerIndex / PathOperations, compiled where they lie by
// oracle/Makefile into oracle/_ref/libpgref.so.  PathAligner = the first stage of grm::CompositeAligner
// (lib/grm/CompositeAligner.cpp:90-107): exact full-length matches anchored by a unique k-mer.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <list>
#include <string>

#include "grm/PathAligner.hh"
#include "graphcore/Graph.hh"

using graphtools::Graph;

extern "C" {

// out8 per read = {mapped, graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, 0};
// counters3 = {attempted, anchored, mapped} (PathAligner.hh:66-68).  out_bases: the read's bases after the call.
int pgref_path_align_batch(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto,
    int kmer_len, int n_reads, const char* bases_blob, const int32_t* read_off, const uint8_t* is_rev, int32_t* out8,
    char* out_bases_blob, char* cigars, int cigar_stride, int32_t* counters3)
{
    try
    {
        Graph g(static_cast<size_t>(n_nodes), false); // paragraph builds Graph{n,false}: lib/grm/GraphInput.cpp:51-161
        for (int i = 0; i < n_nodes; ++i)
        {
            g.setNodeName(i, "n" + std::to_string(i));
            g.setNodeSeq(i, std::string(seq_blob + seq_off[i], seq_blob + seq_off[i + 1]));
        }
        for (int e = 0; e < n_edges; ++e)
            g.addEdge(efrom[e], eto[e]);
        grm::PathAligner al(kmer_len);
        al.setGraph(&g, std::list<graphtools::Path>());
        for (int i = 0; i < n_reads; ++i)
        {
            const int len = read_off[i + 1] - read_off[i];
            common::Read r;
            r.set_bases(std::string(bases_blob + read_off[i], bases_blob + read_off[i + 1]));
            r.set_quals(std::string(static_cast<size_t>(len), '#'));
            r.set_is_reverse_strand(is_rev && is_rev[i]);
            al.alignRead(r);
            int32_t* o = out8 + 8 * i;
            o[0] = r.graph_mapping_status() == common::Read::MAPPED ? 1 : 0;
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
        counters3[0] = static_cast<int32_t>(al.attempted());
        counters3[1] = static_cast<int32_t>(al.anchored());
        counters3[2] = static_cast<int32_t>(al.mapped());
        return 0;
    }
    catch (std::exception const&)
    {
        return -1;
    }
}
}
