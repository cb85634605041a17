// paragraph_b200 -- read filters, disambiguation and fragment counting on the device CIGAR arena.
//
// What the reference does on the host after alignReads, per site:
//   createReadFilter chain NonUniq -> BadAlign            src/c++/lib/paragraph/ReadFilter.cpp:73-90
//   disambiguateReads + the node/edge support filters     src/c++/lib/paragraph/Disambiguation.cpp:82-142, 212-296
//   readsToFragments, countNodes/countEdges/countPathFamilies
//                                                         src/c++/lib/common/Fragment.cpp:141-182,
//                                                         src/c++/lib/paragraph/ReadCounting.cpp:52-127
// There every step re-parses the "id[..]" CIGAR *string* (decodeGraphAlignment is called once by BadAlign, once per
// path node by the node filter, once per path edge by the edge filter, once by disambiguateReads and once by
// Fragment::addRead).  Here the op words the trace kernel left in HBM are walked once per read.
//
// Shared between the CUDA build (pg_kernels.cu) and the CPU lane emulator (tests/emu), like pg_core.cuh.
#pragma once
#include "pg_core.cuh"

namespace pg
{

enum Verdict : uint8_t
{
    V_MAPPED = 0,    // passed the filter chain; takes part in the counts
    V_NONUNIQ = 1,   // readfilters::NonUniq
    V_BAD_ALIGN = 2, // readfilters::BadAlign
    V_INVALID = 3    // decodeGraphAlignment would throw (the reference aborts the site): reported, never counted
};

constexpr uint32_t SUP_NODE_MASK = 0xFFFFu;
constexpr uint32_t SUP_NODE = 0x40000000u; // path word: read supports this node
constexpr uint32_t SUP_EDGE = 0x80000000u; // path word: read supports the edge (previous path node -> this one)

struct ReadSupport // == pg_read_support (include/pg_align.h)
{
    uint64_t sequences; // bit k: edge label (path family) k is in graph_sequences_supported
    uint32_t path_off;  // first path word of this read ( == its record's cigar_off)
    uint16_t path_len;  // number of path nodes; 0 unless MAPPED
    uint8_t verdict;
    uint8_t graph_reverse; // is_graph_reverse_strand = read.is_reverse_strand() != chose_reverse (GraphAligner.cpp:358)
};

struct Count4 // == pg_count4: "<name>", ":READS", ":FWD", ":REV" of ReadCounting.cpp:52-69
{
    uint32_t fragments, reads, fwd, rev;
};

struct CountParams
{
    int32_t remove_nonuniq;
    int32_t use_support_filters;
    double bad_align_frac;
    int32_t family_slots;
};

// Per-site tables for counting, all indexed from the site's bases:
//   node_base / edge_base : first row of the site in node_counts / edge_counts (and in lab_out/lab_in / lab_edge,
//                           csr_input, which are laid out the same way);
//   fam_base / key_base   : the site's family table: `slots` keys at fam_keys[key_base ...] and
//                           slots x (1 + n_nodes + n_edges) Count4 rows at fam_counts[fam_base ...].
//                           slots = min(family_slots, 2^(labels used by the site)): a DEL/INS graph with the two
//                           labels REF and ALT gets 4, only graphs with many haplotype labels get the full table.
struct CountSite
{
    int32_t node_base, edge_base;
    int64_t fam_base;
    int32_t slots, key_base;
};

struct CountTables
{
    const SiteDev* sites;
    const int32_t* gints;
    const CountSite* csite;
    const int32_t* csr_input; // [edge_base + csr position] -> index of the edge in pg_add_graph's input order
    const uint64_t* lab_edge; // [edge_base + csr position] label mask of the edge
    const uint64_t* lab_out;  // [node_base + v] OR of the labels of v's outgoing edges (PathFamily outNodes)
    const uint64_t* lab_in;   // [node_base + v] OR of the labels of v's incoming edges (PathFamily inNodes)
};

// CSR position of edge a->b in the forward predecessor lists, -1 if there is no such edge
PG_HD int csr_edge(const GraphView& g, int a, int b)
{
    for (int p = g.pred_ptr[b]; p < g.pred_ptr[b + 1]; ++p)
        if (g.pred_idx[p] == a)
            return p;
    return -1;
}

struct NodeAln // Alignment::updateCounts of one node's ops (LinearAlignment.cpp:53-84)
{
    int node;
    int matched, mismatched, missing, clipped, inserted, deleted;
    PG_HD int ref_len() const { return matched + mismatched + missing + deleted; }
    PG_HD int query_len() const { return matched + mismatched + missing + inserted + clipped; }
    PG_HD void clear(int n)
    {
        node = n;
        matched = mismatched = missing = clipped = inserted = deleted = 0;
    }
    PG_HD void add(uint32_t w)
    {
        const int len = (int)((w >> 3) & 0x1FFFu);
        switch (w & 7u)
        {
        case OP_M: matched += len; break;
        case OP_X: mismatched += len; break;
        case OP_N: missing += len; break;
        case OP_I: inserted += len; break;
        case OP_D: deleted += len; break;
        case OP_S: clipped += len; break;
        default: break; // OP_NONE: the node is on the path ("id[]") but carries no op
        }
    }
};

PG_HD int imin(int a, int b) { return a < b ? a : b; }

// Disambiguation.cpp:212-243
PG_HD bool node_supported(const NodeAln& a, int node_len, int read_len)
{
    const int half = read_len / 2;
    const int nonmatch = a.mismatched + a.clipped, indel = a.inserted + a.deleted;
    if (node_len < half && (nonmatch > 0 || indel > 0))
        return false;
    return nonmatch + indel <= half;
}

// Disambiguation.cpp:245-296
PG_HD bool edge_supported(const NodeAln& p, const NodeAln& c, int plen, int clen, int read_len)
{
    const int mno = read_len / 10 + 1;
    return p.matched >= imin(p.ref_len(), mno) && c.matched >= imin(c.ref_len(), mno)
        && p.query_len() < 2 * p.ref_len() && c.query_len() < 2 * c.ref_len() && p.matched >= imin(plen, mno)
        && c.matched >= imin(clen, mno);
}

// One read: filter chain, then disambiguateReads.  Writes sup and (for a MAPPED read) one path word per path node at
// path[rec.cigar_off ...].
PG_HD void support_read(const Record& rec, const uint32_t* ops, int read_len, int site, bool is_reverse_strand,
                        const CountTables& t, const CountParams& prm, ReadSupport& sup, uint32_t* path)
{
    sup.sequences = 0;
    sup.path_off = rec.cigar_off;
    sup.path_len = 0;
    // gssw: is_graph_reverse_strand = read.is_reverse_strand() != chose_reverse (GraphAligner.cpp:358-359);
    // PathAligner sets it to the matched strand itself (PathAligner.cpp:124-135)
    sup.graph_reverse = rec.mapped_by == STAGE_PATH ? (uint8_t)(rec.chose_reverse ? 1 : 0)
                                                    : (uint8_t)((is_reverse_strand ? 1 : 0) ^ (rec.chose_reverse ? 1 : 0));
    if (prm.remove_nonuniq && !rec.unique)
    {
        sup.verdict = V_NONUNIQ;
        return;
    }
    const SiteDev& sd = t.sites[site];
    const GraphView g = make_view(sd, nullptr, t.gints, 0);
    const CountSite cs = t.csite[site];
    const uint32_t* op = ops + rec.cigar_off;
    const int n_ops = (int)rec.cigar_len;

    // ---- pass 1: what decodeGraphAlignment / Path::isValid check, and BadAlign's sums
    bool valid = n_ops > 0;
    int qlen = 0, clipped = 0;
    {
        int prev_node = -1, last_ref = 0, n_path = 0;
        for (int x = 0; x < n_ops && valid; ++x)
        {
            const int node = (int)(op[x] >> 16);
            if (node != prev_node)
            {
                if (node >= sd.n_nodes) // cannot come from the trace kernel; alignments imported by the caller may
                    valid = false;
                else if (prev_node >= 0 && (prev_node > node || csr_edge(g, prev_node, node) < 0))
                    valid = false;
                prev_node = node;
                last_ref = 0;
                ++n_path;
            }
            NodeAln one;
            one.clear(node);
            one.add(op[x]);
            last_ref += one.ref_len();
            qlen += one.query_len();
            clipped += one.clipped;
        }
        if (valid)
        {
            const int first = (int)(op[0] >> 16);
            const int end = (n_path == 1 ? rec.graph_pos : 0) + last_ref - 1;
            valid = rec.graph_pos >= 0 && rec.graph_pos < g.node_len[first] && end >= 0 && end < g.node_len[prev_node]
                && !(n_path == 1 && rec.graph_pos > end);
        }
    }
    if (!valid)
    {
        sup.verdict = V_INVALID;
        return;
    }
    // BadAlign.hh:62-73: aligned < round(frac * queryLength); round() of a non-negative double
    const double thr = (double)(long long)(prm.bad_align_frac * (double)qlen + 0.5);
    if ((double)(qlen - clipped) < thr)
    {
        sup.verdict = V_BAD_ALIGN;
        return;
    }
    sup.verdict = V_MAPPED;

    // ---- pass 2: node by node
    NodeAln prev, cur;
    prev.clear(-1);
    int k = 0;
    uint64_t overlapped = 0, fail = 0;
    int x = 0;
    while (x < n_ops)
    {
        cur.clear((int)(op[x] >> 16));
        while (x < n_ops && (int)(op[x] >> 16) == cur.node)
            cur.add(op[x++]);
        uint32_t w = (uint32_t)cur.node;
        if (k > 0)
        {
            const int e = csr_edge(g, prev.node, cur.node);
            const uint64_t lab = t.lab_edge[cs.edge_base + e];
            if (!prm.use_support_filters
                || edge_supported(prev, cur, g.node_len[prev.node], g.node_len[cur.node], read_len))
            {
                w |= SUP_EDGE;
                overlapped |= lab;
            }
            // PathFamily::containsPath (PathFamily.cpp:88-106): a step that is not in the family but leaves one of
            // its out-nodes or enters one of its in-nodes breaks the family
            fail |= ~lab & (t.lab_out[cs.node_base + prev.node] | t.lab_in[cs.node_base + cur.node]);
        }
        if (!prm.use_support_filters || node_supported(cur, g.node_len[cur.node], read_len))
            w |= SUP_NODE;
        path[rec.cigar_off + k] = w;
        ++k;
        prev = cur;
    }
    sup.path_len = (uint16_t)k;
    sup.sequences = overlapped & ~fail;
}

// ---------------------------------------------------------------------------------------------
// fragments: reads with the same fragment id are chained (next[]) in input order by the host; the chain head
// accumulates the whole fragment (Fragment::addRead, Fragment.cpp:33-67 counters and :141-156 set unions) and adds
// it to the site tables.
// ---------------------------------------------------------------------------------------------
#ifdef __CUDA_ARCH__
#define PG_ATOMIC_ADD(p, v) atomicAdd((p), (v))
#else
#define PG_ATOMIC_ADD(p, v) (*(p) += (v))
#endif

PG_HD void add4(Count4* c, uint32_t reads, uint32_t fwd, uint32_t rev)
{
    PG_ATOMIC_ADD(&c->fragments, 1u);
    PG_ATOMIC_ADD(&c->reads, reads);
    PG_ATOMIC_ADD(&c->fwd, fwd);
    PG_ATOMIC_ADD(&c->rev, rev);
}

// slot of `key` in the site's family table (open addressing over family_slots keys), -1 when the table is full
PG_HD int family_slot(unsigned long long* keys, int slots, unsigned long long key)
{
    int h = (int)((key * 0x9E3779B97F4A7C15ull) >> 40) % slots;
    for (int probe = 0; probe < slots; ++probe)
    {
#ifdef __CUDA_ARCH__
        const unsigned long long old = atomicCAS(&keys[h], 0ull, key);
#else
        const unsigned long long old = keys[h];
        if (old == 0)
            keys[h] = key;
#endif
        if (old == 0 || old == key)
            return h;
        h = (h + 1) % slots;
    }
    return -1;
}

// has an earlier MAPPED member of the chain (before `upto`) already contributed this node / edge?
PG_HD bool seen_before(int head, int upto, const int32_t* next, const ReadSupport* sup, const uint32_t* path, int node,
                       int from /* -1: node query */)
{
    for (int i = head; i != upto; i = next[i])
    {
        if (sup[i].verdict != V_MAPPED)
            continue;
        const uint32_t* w = path + sup[i].path_off;
        for (int k = 0; k < sup[i].path_len; ++k)
        {
            if ((int)(w[k] & SUP_NODE_MASK) != node)
                continue;
            if (from < 0 ? (w[k] & SUP_NODE) != 0
                         : (k > 0 && (w[k] & SUP_EDGE) && (int)(w[k - 1] & SUP_NODE_MASK) == from))
                return true;
        }
    }
    return false;
}

// returns false when the site's family table overflowed
PG_HD bool count_fragment(int head, int site, const int32_t* next, const ReadSupport* sup, const uint32_t* path,
                          const CountTables& t, const CountParams& prm, Count4* node_counts, Count4* edge_counts,
                          unsigned long long* fam_keys, Count4* fam_counts)
{
    uint32_t reads = 0, fwd = 0, rev = 0;
    unsigned long long seqs = 0;
    for (int i = head; i >= 0; i = next[i])
    {
        if (sup[i].verdict != V_MAPPED)
            continue;
        ++reads;
        if (sup[i].graph_reverse)
            ++rev;
        else
            ++fwd;
        seqs |= sup[i].sequences;
    }
    if (reads == 0)
        return true;
    const SiteDev& sd = t.sites[site];
    const GraphView g = make_view(sd, nullptr, t.gints, 0);
    const CountSite cs = t.csite[site];
    Count4* fam = nullptr;
    if (seqs)
    {
        const int slot = family_slot(fam_keys + cs.key_base, cs.slots, seqs);
        if (slot < 0)
            return false;
        fam = fam_counts + cs.fam_base + (int64_t)slot * (1 + sd.n_nodes + sd.n_edges);
        add4(&fam[0], reads, fwd, rev);
    }
    for (int i = head; i >= 0; i = next[i])
    {
        if (sup[i].verdict != V_MAPPED)
            continue;
        const uint32_t* w = path + sup[i].path_off;
        for (int k = 0; k < sup[i].path_len; ++k)
        {
            const int node = (int)(w[k] & SUP_NODE_MASK);
            if ((w[k] & SUP_NODE) && !seen_before(head, i, next, sup, path, node, -1))
            {
                bool dup = false; // the same node twice in one path cannot happen (ids ascend), kept for safety
                for (int q = 0; q < k; ++q)
                    dup |= (int)(w[q] & SUP_NODE_MASK) == node && (w[q] & SUP_NODE);
                if (!dup)
                {
                    add4(&node_counts[cs.node_base + node], reads, fwd, rev);
                    if (fam)
                        add4(&fam[1 + node], reads, fwd, rev);
                }
            }
            if (k > 0 && (w[k] & SUP_EDGE))
            {
                const int from = (int)(w[k - 1] & SUP_NODE_MASK);
                if (!seen_before(head, i, next, sup, path, node, from))
                {
                    const int e = t.csr_input[cs.edge_base + csr_edge(g, from, node)];
                    add4(&edge_counts[cs.edge_base + e], reads, fwd, rev);
                    if (fam)
                        add4(&fam[1 + sd.n_nodes + e], reads, fwd, rev);
                }
            }
        }
    }
    return true;
}

} // namespace pg
