// TEST INFRASTRUCTURE ONLY (oracle/): C-ABI driver around the UNMODIFIED
// reference sources, compiled where they lie under /root/reference by
// oracle/Makefile into oracle/_ref/libpgref.so.  Nothing in the product
// (paragraph_b200/, include/) links or loads this library.
//
// Two entry families:
//   pgref_aligner_*  : grm::GraphAligner (src/c++/lib/grm/GraphAligner.cpp:277-404)
//                      = 4 gssw fills + tracebacks + uniqueness + strand choice.
//   pgref_gssw_*     : raw gssw (external/gssw/gssw.c) fill / traceback with the
//                      de-striped mH/mE/mF matrices exposed, to pin oracle/pg_oracle.c
//                      cell by cell.
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstring>
#include <math.h> // ::round for the reference BadAlign.hh (it relies on a transitive include)
#include <memory>
#include <string>
#include <thread>
#include <vector>

#include "grm/GraphAligner.hh"
#include "graphcore/Graph.hh"
#include "graphalign/GraphAlignmentOperations.hh"
// the reference's read filters applied after alignment (src/c++/lib/paragraph/ReadFilter.cpp:73-90); header-only classes
#include "../lib/paragraph/readfilters/BadAlign.hh"
#include "../lib/paragraph/readfilters/NonUniq.hh"

extern "C" {
#include "gssw.h"
}

using graphtools::Graph;

namespace
{
struct RefAligner
{
    Graph graph;
    grm::GraphAligner aligner;
};

Graph makeGraph(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto)
{
    // paragraph builds Graph{n,false}: src/c++/lib/grm/GraphInput.cpp:51-161
    Graph g(static_cast<size_t>(n_nodes), false);
    for (int i = 0; i < n_nodes; ++i)
    {
        g.setNodeName(i, "n" + std::to_string(i));
        g.setNodeSeq(i, std::string(seq_blob + seq_off[i], seq_blob + seq_off[i + 1]));
    }
    for (int e = 0; e < n_edges; ++e)
    {
        g.addEdge(efrom[e], eto[e]);
    }
    return g;
}

void alignOne(
    grm::GraphAligner const& al, const char* bases, int len, int is_rev, unsigned flags, int32_t* out6, char* out_bases,
    char* cigar, int cigar_cap)
{
    common::Read r;
    r.set_bases(std::string(bases, bases + len));
    r.set_quals(std::string(static_cast<size_t>(len), '#'));
    r.set_is_reverse_strand(is_rev != 0);
    al.alignRead(r, flags);
    out6[0] = r.graph_pos();
    out6[1] = r.graph_alignment_score();
    out6[2] = r.is_graph_alignment_unique() ? 1 : 0;
    out6[3] = r.graph_mapq();
    out6[4] = r.is_graph_reverse_strand() ? 1 : 0;
    out6[5] = static_cast<int32_t>(r.graph_cigar().size());
    if (out_bases)
    {
        memcpy(out_bases, r.bases().data(), static_cast<size_t>(len));
    }
    if (cigar && cigar_cap > 0)
    {
        size_t n = std::min(static_cast<size_t>(cigar_cap - 1), r.graph_cigar().size());
        memcpy(cigar, r.graph_cigar().data(), n);
        cigar[n] = 0;
    }
}
}

extern "C" {

void* pgref_aligner_create(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto)
{
    try
    {
        auto* h = new RefAligner{ makeGraph(n_nodes, seq_blob, seq_off, n_edges, efrom, eto), grm::GraphAligner() };
        h->aligner.setGraph(&h->graph);
        return h;
    }
    catch (std::exception const&)
    {
        return nullptr;
    }
}

void pgref_aligner_destroy(void* h) { delete static_cast<RefAligner*>(h); }

// out6 = {graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen}
void pgref_aligner_align(
    void* h, const char* bases, int len, int is_reverse_strand, unsigned flags, int32_t* out6, char* out_bases,
    char* cigar, int cigar_cap)
{
    alignOne(static_cast<RefAligner*>(h)->aligner, bases, len, is_reverse_strand, flags, out6, out_bases, cigar, cigar_cap);
}

// Batch over `threads` host threads, one GraphAligner per thread, reads split into
// contiguous chunks — the scheme of src/c++/lib/grm/Align.cpp:107-153.
// cigars: n_reads * cigar_stride bytes (may be NULL).  Returns 0.
int pgref_align_batch(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto,
    int n_reads, const char* bases_blob, const int32_t* read_off, const uint8_t* is_rev, unsigned flags, int threads,
    int32_t* out6, char* out_bases_blob, char* cigars, int cigar_stride)
{
    if (threads < 1)
        threads = 1;
    Graph graph = makeGraph(n_nodes, seq_blob, seq_off, n_edges, efrom, eto);
    int step = (n_reads + threads - 1) / threads;
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
    {
        int b = t * step, e = std::min(n_reads, b + step);
        if (b >= e)
            break;
        pool.emplace_back([&, b, e]() {
            grm::GraphAligner al;
            al.setGraph(&graph);
            for (int i = b; i < e; ++i)
            {
                alignOne(
                    al, bases_blob + read_off[i], read_off[i + 1] - read_off[i], is_rev ? is_rev[i] : 0, flags,
                    out6 + 6 * i, out_bases_blob ? out_bases_blob + read_off[i] : nullptr,
                    cigars ? cigars + static_cast<size_t>(i) * cigar_stride : nullptr, cigar_stride);
            }
        });
    }
    for (auto& th : pool)
        th.join();
    return 0;
}

// Many sites in one call: `threads` host threads pull whole sites from a shared counter (largest first when the caller
// sorted them so), one GraphAligner per (thread, site) -- the scheme of src/c++/lib/grmpy/Workflow.cpp:108-146, where
// threads pull (sample, graph) pairs.  Graph s = nodes [node_ptr[s], node_ptr[s+1]) of seq_off (offsets into seq_blob,
// n_nodes_total + 1 entries) and edges [edge_ptr[s], edge_ptr[s+1]) (node ids local to the site); its reads are
// [read_ptr[s], read_ptr[s+1]).  Outputs as pgref_align_batch.  Returns the number of sites that threw.
int pgref_align_sites(
    int n_sites, const int32_t* node_ptr, const char* seq_blob, const int32_t* seq_off, const int32_t* edge_ptr,
    const int32_t* efrom, const int32_t* eto, const int32_t* read_ptr, const char* bases_blob, const int32_t* read_off,
    const uint8_t* is_rev, unsigned flags, int threads, int32_t* out6, char* out_bases_blob, char* cigars,
    int cigar_stride)
{
    if (threads < 1)
        threads = 1;
    std::atomic<int> next(0), failed(0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
    {
        pool.emplace_back([&]() {
            for (;;)
            {
                const int s = next.fetch_add(1);
                if (s >= n_sites)
                    return;
                try
                {
                    const int n0 = node_ptr[s], nn = node_ptr[s + 1] - n0;
                    std::vector<int32_t> off(static_cast<size_t>(nn) + 1);
                    for (int i = 0; i <= nn; ++i)
                        off[i] = seq_off[n0 + i] - seq_off[n0];
                    const int e0 = edge_ptr[s], ne = edge_ptr[s + 1] - e0;
                    Graph graph = makeGraph(nn, seq_blob + seq_off[n0], off.data(), ne, efrom + e0, eto + e0);
                    grm::GraphAligner al;
                    al.setGraph(&graph);
                    for (int i = read_ptr[s]; i < read_ptr[s + 1]; ++i)
                    {
                        alignOne(
                            al, bases_blob + read_off[i], read_off[i + 1] - read_off[i], is_rev ? is_rev[i] : 0, flags,
                            out6 + 6 * i, out_bases_blob ? out_bases_blob + read_off[i] : nullptr,
                            cigars ? cigars + static_cast<size_t>(i) * cigar_stride : nullptr, cigar_stride);
                    }
                }
                catch (std::exception const&)
                {
                    failed.fetch_add(1);
                }
            }
        });
    }
    for (auto& th : pool)
        th.join();
    return failed.load();
}

// ---------------------------------------------------------------- read filters (first piece)
// For each read: decode its graph CIGAR with the reference's decodeGraphAlignment (which also validates it against
// the graph and throws on any inconsistency) and run readfilters::NonUniq and readfilters::BadAlign.
// out4[i] = {decode_ok, query_clipped, nonuniq_filtered, badalign_filtered}
int pgref_filter_batch(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto,
    int n_reads, const int32_t* read_len, const int32_t* graph_pos, const uint8_t* unique, const char* cigars,
    int cigar_stride, double bad_align_frac, int32_t* out4)
{
    Graph graph = makeGraph(n_nodes, seq_blob, seq_off, n_edges, efrom, eto);
    paragraph::readfilters::BadAlign bad(&graph, bad_align_frac);
    paragraph::readfilters::NonUniq nonuniq;
    int failures = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        common::Read r;
        r.set_bases(std::string(static_cast<size_t>(read_len[i]), 'A'));
        r.set_graph_pos(graph_pos[i]);
        r.set_graph_cigar(std::string(cigars + static_cast<size_t>(i) * cigar_stride));
        r.set_is_graph_alignment_unique(unique[i] != 0);
        int32_t* o = out4 + 4 * i;
        o[0] = o[1] = o[2] = o[3] = 0;
        o[2] = nonuniq.filterRead(r).first ? 1 : 0;
        if (r.graph_cigar().empty()) // score-0 read / AF_CIGAR off: nothing to decode (the reference's NonUniq removes it first)
            continue;
        try
        {
            const graphtools::GraphAlignment m = graphtools::decodeGraphAlignment(r.graph_pos(), r.graph_cigar(), &graph);
            size_t clipped = 0;
            for (auto const& aln : m)
                clipped += aln.numClipped();
            o[0] = (m.queryLength() == static_cast<uint32_t>(read_len[i])) ? 1 : 0;
            o[1] = static_cast<int32_t>(clipped);
            o[3] = bad.filterRead(r).first ? 1 : 0;
        }
        catch (std::exception const&)
        {
            ++failures;
        }
    }
    return failures;
}

// ---------------------------------------------------------------- raw gssw
struct RefGssw
{
    gssw_graph* g;
    std::vector<gssw_node*> nodes;
    int8_t* nt;
    int8_t* mat;
};

void* pgref_gssw_create(
    int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom, const int32_t* eto)
{
    auto* h = new RefGssw;
    h->nt = gssw_create_nt_table();
    h->mat = gssw_create_score_matrix(1, 4);
    for (int i = 0; i < n_nodes; ++i)
    {
        std::string s(seq_blob + seq_off[i], seq_blob + seq_off[i + 1]);
        h->nodes.push_back(gssw_node_create(nullptr, static_cast<uint32_t>(i), s.c_str(), h->nt, h->mat));
    }
    // predecessors in ascending id per node, as GraphAligner.cpp:147-157 (std::set order)
    for (int to = 0; to < n_nodes; ++to)
    {
        std::vector<int> preds;
        for (int e = 0; e < n_edges; ++e)
            if (eto[e] == to)
                preds.push_back(efrom[e]);
        std::sort(preds.begin(), preds.end());
        for (int p : preds)
            gssw_nodes_add_edge(h->nodes[p], h->nodes[to]);
    }
    h->g = gssw_graph_create(static_cast<uint32_t>(n_nodes));
    for (auto* n : h->nodes)
        gssw_graph_add_node(h->g, n);
    return h;
}

void pgref_gssw_destroy(void* hv)
{
    auto* h = static_cast<RefGssw*>(hv);
    gssw_graph_destroy(h->g);
    free(h->nt);
    free(h->mat);
    delete h;
}

// Fill + traceback one (already upper-cased) read.  Outputs:
//   node_stats[4*n] = {score1, ref_end1, read_end1, is_byte} per node
//   mats8 : if non-NULL, concatenated per node: mH[len*L], mE[len*L], mF[len*L]  (bytes; byte mode only)
//   mats16: if non-NULL, the same cells widened to 16 bits (byte mode) or copied (word mode, after gssw's fallback)
//   res3 = {max_node_id, position, score}; cigar string "id[..]id[..]"
static int fillTrace(
    void* hv, const char* read, int32_t* node_stats, uint8_t* mats8, uint16_t* mats16, int32_t* res3, char* cigar,
    int cigar_cap)
{
    auto* h = static_cast<RefGssw*>(hv);
    int L = static_cast<int>(strlen(read));
    gssw_graph_fill(h->g, read, h->nt, h->mat, 6, 1, 15, 2);
    size_t off = 0;
    for (size_t i = 0; i < h->nodes.size(); ++i)
    {
        gssw_align* a = h->nodes[i]->alignment;
        node_stats[4 * i + 0] = a->score1;
        node_stats[4 * i + 1] = a->ref_end1;
        node_stats[4 * i + 2] = a->read_end1;
        node_stats[4 * i + 3] = a->is_byte;
        size_t sz = static_cast<size_t>(h->nodes[i]->len) * L;
        const void* m3[3] = { a->mH, a->mE, a->mF };
        if (mats8 && a->is_byte)
        {
            for (int k = 0; k < 3; ++k)
                memcpy(mats8 + off + k * sz, m3[k], sz);
        }
        if (mats16)
        {
            for (int k = 0; k < 3; ++k)
                for (size_t x = 0; x < sz; ++x)
                    mats16[off + k * sz + x] = a->is_byte ? static_cast<const uint8_t*>(m3[k])[x]
                                                          : static_cast<const uint16_t*>(m3[k])[x];
        }
        off += 3 * sz;
    }
    gssw_graph_mapping* gm = gssw_graph_trace_back(h->g, read, L, h->nt, h->mat, 6, 1);
    res3[0] = static_cast<int32_t>(h->g->max_node->id);
    res3[1] = gm->position;
    res3[2] = gm->score;
    std::string s;
    for (uint32_t i = 0; i < gm->cigar.length; ++i)
    {
        gssw_node_cigar* nc = gm->cigar.elements + i;
        s += std::to_string(nc->node->id) + "[";
        for (int32_t j = 0; j < nc->cigar->length; ++j)
            s += std::to_string(nc->cigar->elements[j].length) + nc->cigar->elements[j].type;
        s += "]";
    }
    gssw_graph_mapping_destroy(gm);
    if (cigar && cigar_cap > 0)
    {
        size_t n = std::min(static_cast<size_t>(cigar_cap - 1), s.size());
        memcpy(cigar, s.data(), n);
        cigar[n] = 0;
    }
    return static_cast<int>(s.size());
}

int pgref_gssw_fill_trace(
    void* hv, const char* read, int32_t* node_stats, uint8_t* mats, int32_t* res3, char* cigar, int cigar_cap)
{
    return fillTrace(hv, read, node_stats, mats, nullptr, res3, cigar, cigar_cap);
}

int pgref_gssw_fill_trace16(
    void* hv, const char* read, int32_t* node_stats, uint16_t* mats, int32_t* res3, char* cigar, int cigar_cap)
{
    return fillTrace(hv, read, node_stats, nullptr, mats, res3, cigar, cigar_cap);
}
}
