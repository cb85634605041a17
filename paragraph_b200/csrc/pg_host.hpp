// paragraph_b200 -- host-side construction of the device graph tables (pure C++, no CUDA).
// Replaces GraphAlignerImpl::initializeGraph for the graph and reverseGraph(graph)
// (src/c++/lib/grm/GraphAligner.cpp:110-167, 277-285): upper-case the node sequences (:126,138), map
// characters to gssw codes (gssw_create_nt_table, gssw.c:4206-4220), predecessor lists in ascending id
// order (:147-157), plus the reversed graph (graph-tools src/graphcore/GraphOperations.cpp:38-60).
// Used by the C-ABI (pg_capi.cu) and by the CPU lane emulator in tests/emu.
#pragma once
#include <algorithm>
#include <cstdint>
#include <string>
#include <vector>

#include "pg_core.cuh"

namespace pg
{
namespace host
{

struct GraphStore
{
    std::vector<SiteDev> sites;
    std::vector<uint8_t> bytes;
    std::vector<int32_t> ints;
    int max_nodes = 0;
    int max_G = 0;

    void clear()
    {
        sites.clear();
        bytes.clear();
        ints.clear();
        max_nodes = max_G = 0;
    }

    // returns site id >= 0, or -1 with err set
    int add(int n_nodes, const char* blob, const int32_t* off, int n_edges, const int32_t* ef, const int32_t* et,
            std::string& err)
    {
        if (n_nodes <= 0 || !blob || !off || n_edges < 0 || (n_edges > 0 && (!ef || !et)))
        {
            err = "pg_add_graph: bad arguments";
            return -1;
        }
        if (n_nodes > 65535)
        {
            err = "pg_add_graph: more than 65535 nodes";
            return -1;
        }
        int64_t G = 0;
        for (int i = 0; i < n_nodes; ++i)
        {
            if (off[i + 1] <= off[i])
            {
                err = "pg_add_graph: empty node sequence (node " + std::to_string(i) + ")";
                return -1;
            }
            G += off[i + 1] - off[i];
        }
        if (G > (1 << 24))
        {
            err = "pg_add_graph: graph longer than 16 Mbases";
            return -1;
        }
        for (int e = 0; e < n_edges; ++e)
        {
            // graphtools::Graph::addEdge throws when source > sink (Graph.cpp:113-116); a self loop has no
            // meaning for the DP either (GraphAligner.cpp:149 asserts pred < node)
            if (ef[e] < 0 || et[e] >= n_nodes || ef[e] >= et[e])
            {
                err = "pg_add_graph: edge " + std::to_string(ef[e]) + "->" + std::to_string(et[e])
                    + " breaks topological order";
                return -1;
            }
        }
        SiteDev sd;
        sd.n_nodes = n_nodes;
        sd.G = (int32_t)G;
        sd.n_edges = n_edges;
        const int32_t base0 = off[0];
        // forward characters (upper-cased), concatenated in node order
        std::vector<uint8_t> chars((size_t)G);
        for (int64_t x = 0; x < G; ++x)
            chars[(size_t)x] = to_upper((uint8_t)blob[base0 + x]);
        for (int o = 0; o < 2; ++o)
        {
            while (bytes.size() % 16) // the staged span [codes - SENT, ...) must be 16-byte aligned for the bulk copy (TMA)
                bytes.push_back(0);
            bytes.insert(bytes.end(), SENT, (uint8_t)5);
            sd.codes_off[o] = (int32_t)bytes.size();
            for (int64_t x = 0; x < G; ++x)
                bytes.push_back((uint8_t)nt_code(chars[(size_t)(o == 0 ? x : G - 1 - x)]));
            bytes.insert(bytes.end(), SENT + CK + 4 + 16, (uint8_t)5); // + slack so the span can be rounded up to 16 bytes
        }
        sd.chars_off = (int32_t)bytes.size();
        bytes.insert(bytes.end(), chars.begin(), chars.end());
        for (int o = 0; o < 2; ++o)
        {
            sd.tab_off[o] = (int32_t)ints.size();
            std::vector<int32_t> len((size_t)n_nodes), start((size_t)n_nodes);
            for (int i = 0; i < n_nodes; ++i)
            {
                const int src = o == 0 ? i : n_nodes - 1 - i;
                len[(size_t)i] = off[src + 1] - off[src];
            }
            int32_t acc = 0;
            for (int i = 0; i < n_nodes; ++i)
            {
                start[(size_t)i] = acc;
                acc += len[(size_t)i];
            }
            std::vector<std::vector<int32_t>> preds((size_t)n_nodes);
            for (int e = 0; e < n_edges; ++e)
            {
                const int f = o == 0 ? ef[e] : n_nodes - 1 - et[e];
                const int t = o == 0 ? et[e] : n_nodes - 1 - ef[e];
                preds[(size_t)t].push_back(f);
            }
            ints.insert(ints.end(), start.begin(), start.end());
            ints.insert(ints.end(), len.begin(), len.end());
            int32_t ptr = 0;
            for (int i = 0; i < n_nodes; ++i)
            {
                ints.push_back(ptr);
                auto& p = preds[(size_t)i];
                std::sort(p.begin(), p.end());
                p.erase(std::unique(p.begin(), p.end()), p.end()); // Graph::addEdge rejects duplicates anyway
                ptr += (int32_t)p.size();
            }
            ints.push_back(ptr);
            for (int i = 0; i < n_nodes; ++i)
                ints.insert(ints.end(), preds[(size_t)i].begin(), preds[(size_t)i].end());
            // pad pred_idx to n_edges entries so both orientations occupy the same space
            for (int32_t x = ptr; x < n_edges; ++x)
                ints.push_back(0);
        }
        sites.push_back(sd);
        max_nodes = std::max(max_nodes, n_nodes);
        max_G = std::max(max_G, (int)G);
        return (int)sites.size() - 1;
    }
};

// per-read scratch sizes in 32-bit words (see pg_core.cuh "per-task scratch layout")
inline size_t info_words(int max_nodes, int W) { return (size_t)max_nodes * 3 * W; }
inline size_t last_words(int max_nodes, int R, int W) { return (size_t)max_nodes * 2 * R * W; }
inline size_t ckpt_words(int max_G, int R, int W) { return (size_t)num_ckpt(max_G, W) * (R + 1) * W; }

// "<node>[<len><op>...]..." -- GraphAlignerImpl::extractCigar (GraphAligner.cpp:88-108)
inline std::string format_cigar(const Record& rec, const uint32_t* ops)
{
    static const char OPC[] = "MXNIDS??";
    std::string s;
    int cur = -1;
    for (uint32_t i = 0; i < rec.cigar_len; ++i)
    {
        const uint32_t w = ops[rec.cigar_off + i];
        const int node = (int)(w >> 16);
        if (node != cur)
        {
            if (cur >= 0)
                s += "]";
            s += std::to_string(node) + "[";
            cur = node;
        }
        if ((w & 7u) == 7u) // OP_NONE: node present, nothing to print
            continue;
        s += std::to_string((w >> 3) & 0x1FFFu);
        s += OPC[w & 7u];
    }
    if (cur >= 0)
        s += "]";
    return s;
}

} // namespace host
} // namespace pg
