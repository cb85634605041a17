// paragraph_b200 -- host-side construction of the device graph tables (pure C++, no CUDA).
// Replaces GraphAlignerImpl::initializeGraph for the graph and reverseGraph(graph)
// (src/c++/lib/grm/GraphAligner.cpp:110-167, 277-285): upper-case the node sequences (:126,138), map
// characters to gssw codes (gssw_create_nt_table, gssw.c:4206-4220), predecessor lists in ascending id
// order (:147-157), plus the reversed graph (graph-tools src/graphcore/GraphOperations.cpp:38-60).
// Used by the C-ABI (pg_capi.cu) and by the CPU lane emulator in tests/emu.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "pg_core.cuh"
#include "pg_count.cuh"
#include "pg_path.cuh"

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
    int max_tab_ints = 0; // largest per-orientation int table (SiteDev::tab_ints)
    // for the counting stage (pg_count.cuh): the edges as given (input order is the order of the edge count rows),
    // their path-family label masks (pg_set_edge_labels; 0 = unlabelled) and each site's first row
    std::vector<int32_t> in_from, in_to;
    std::vector<uint64_t> in_label;
    std::vector<int64_t> edge_base, node_base; // [n_sites + 1]

    void clear()
    {
        sites.clear();
        bytes.clear();
        ints.clear();
        max_nodes = max_G = max_tab_ints = 0;
        in_from.clear();
        in_to.clear();
        in_label.clear();
        edge_base.clear();
        node_base.clear();
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
        sd.raw_off = (int32_t)bytes.size(); // as given: the exact-match stage compares characters case-sensitively
        bytes.insert(bytes.end(), (const uint8_t*)blob + base0, (const uint8_t*)blob + base0 + G);
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
        sd.tab_ints = 3 * n_nodes + 1 + n_edges;
        max_tab_ints = std::max(max_tab_ints, sd.tab_ints);
        sites.push_back(sd);
        if (edge_base.empty())
        {
            edge_base.push_back(0);
            node_base.push_back(0);
        }
        in_from.insert(in_from.end(), ef, ef + n_edges);
        in_to.insert(in_to.end(), et, et + n_edges);
        in_label.insert(in_label.end(), (size_t)n_edges, 0ull);
        edge_base.push_back(edge_base.back() + n_edges);
        node_base.push_back(node_base.back() + n_nodes);
        max_nodes = std::max(max_nodes, n_nodes);
        max_G = std::max(max_G, (int)G);
        return (int)sites.size() - 1;
    }
};

// ---------------------------------------------------------------------------------------------
// counting stage, host side (tables for pg_count.cuh)
// ---------------------------------------------------------------------------------------------
struct CountHostTables
{
    std::vector<CountSite> csite;
    std::vector<int32_t> csr_input;
    std::vector<uint64_t> lab_edge, lab_out, lab_in;
    int64_t fam_rows = 0, fam_keys = 0;
};

// per site: CSR edge -> input edge index, label masks per CSR edge and per node, bases of the count rows
inline void build_count_tables(const GraphStore& gs, int slots, CountHostTables& t)
{
    const size_t ns = gs.sites.size();
    t.csr_input.assign((size_t)gs.edge_base[ns], 0);
    t.lab_edge.assign((size_t)gs.edge_base[ns], 0);
    t.lab_out.assign((size_t)gs.node_base[ns], 0);
    t.lab_in.assign((size_t)gs.node_base[ns], 0);
    t.csite.resize(ns);
    int64_t fam = 0, keys = 0;
    for (size_t s = 0; s < ns; ++s)
    {
        const SiteDev& sd = gs.sites[s];
        const GraphView g = make_view(sd, gs.bytes.data(), gs.ints.data(), 0);
        const int64_t eb = gs.edge_base[s], nb = gs.node_base[s];
        uint64_t used = 0;
        for (int e = 0; e < sd.n_edges; ++e)
            used |= gs.in_label[(size_t)(eb + e)];
        int n_labels = 0;
        for (; used; used &= used - 1)
            ++n_labels;
        const int site_slots = n_labels >= 30 ? slots : (int)std::min<int64_t>(slots, (int64_t)1 << n_labels);
        t.csite[s] = CountSite{ (int32_t)nb, (int32_t)eb, fam, site_slots, (int32_t)keys };
        fam += (int64_t)site_slots * (1 + sd.n_nodes + sd.n_edges);
        keys += site_slots;
        for (int e = sd.n_edges - 1; e >= 0; --e) // descending: a duplicated input edge maps to its first occurrence
            t.csr_input[(size_t)(eb + csr_edge(g, gs.in_from[(size_t)(eb + e)], gs.in_to[(size_t)(eb + e)]))] = e;
        for (int e = 0; e < sd.n_edges; ++e)
        {
            const int a = gs.in_from[(size_t)(eb + e)], b = gs.in_to[(size_t)(eb + e)];
            const uint64_t lab = gs.in_label[(size_t)(eb + e)];
            t.lab_edge[(size_t)(eb + csr_edge(g, a, b))] |= lab;
            t.lab_out[(size_t)(nb + a)] |= lab;
            t.lab_in[(size_t)(nb + b)] |= lab;
        }
    }
    t.fam_rows = fam;
    t.fam_keys = keys;
}

// Chain the reads of each fragment in input order (readsToFragments keeps one Fragment per fragment_id,
// Fragment.cpp:165-181); head[i] = 1 for the first read of a fragment.  site may be nullptr (single site).
inline bool build_fragment_chains(const int32_t* fragment, const int32_t* site, int n, std::vector<int32_t>& next,
                                  std::vector<uint8_t>& head, std::string& err)
{
    next.assign((size_t)n, -1);
    head.assign((size_t)n, 1);
    if (!fragment)
        return true;
    std::unordered_map<int32_t, int32_t> last;
    last.reserve((size_t)n);
    for (int i = 0; i < n; ++i)
    {
        if (fragment[i] < 0)
        {
            err = "negative fragment id at read " + std::to_string(i);
            return false;
        }
        auto it = last.find(fragment[i]);
        if (it == last.end())
        {
            last.emplace(fragment[i], i);
            continue;
        }
        if (site && site[it->second] != site[i])
        {
            err = "fragment " + std::to_string(fragment[i]) + " spans two sites";
            return false;
        }
        next[(size_t)it->second] = i;
        head[(size_t)i] = 0;
        it->second = i;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// exact-match stage (pg_path.cuh): per-site index of the UNIQUE k-mer paths
// ---------------------------------------------------------------------------------------------
// Replaces graphtools::KmerIndex(graph, k) (graph-tools src/graphalign/KmerIndex.cpp:75-125): every path of k
// characters through the graph (extendPathEnd, src/graphcore/PathOperations.cpp:73-103) is enumerated; PathAligner only
// ever uses k-mers that spell exactly one path (numPaths(kmer) == 1, PathAligner.cpp:99), so only those are kept.
struct PathIndexHost
{
    std::vector<PathSite> sites;   // one per GraphStore site
    std::vector<PathEntry> table;  // open addressing, per site a power-of-two block
    std::vector<int32_t> lists;    // node lists of the entries
    std::vector<int32_t> succ;     // per site: succ_ptr[n+1], succ_idx[]
    int k = 0;
};

inline void build_path_index(const GraphStore& gs, int k, PathIndexHost& out)
{
    out = PathIndexHost();
    out.k = k;
    struct KP
    {
        uint64_t h;
        int32_t start_pos, end_pos, list_off, n_nodes;
    };
    for (size_t si = 0; si < gs.sites.size(); ++si)
    {
        const SiteDev& sd = gs.sites[si];
        const int n = sd.n_nodes;
        const int32_t* t = gs.ints.data() + sd.tab_off[0];
        const int32_t *node_start = t, *node_len = t + n;
        const uint8_t* raw = gs.bytes.data() + sd.raw_off;
        // successors, ascending, without duplicates
        std::vector<std::vector<int32_t>> succ((size_t)n);
        for (int64_t e = gs.edge_base[si]; e < gs.edge_base[si + 1]; ++e)
            succ[(size_t)gs.in_from[(size_t)e]].push_back(gs.in_to[(size_t)e]);
        PathSite ps;
        ps.k = k;
        ps.raw_off = sd.raw_off;
        ps.succ_ptr_off = (int32_t)out.succ.size();
        {
            int32_t ptr = 0;
            for (int i = 0; i < n; ++i)
            {
                auto& v = succ[(size_t)i];
                std::sort(v.begin(), v.end());
                v.erase(std::unique(v.begin(), v.end()), v.end());
                out.succ.push_back(ptr);
                ptr += (int32_t)v.size();
            }
            out.succ.push_back(ptr);
            for (int i = 0; i < n; ++i)
                out.succ.insert(out.succ.end(), succ[(size_t)i].begin(), succ[(size_t)i].end());
        }
        // enumerate the k-mer paths (depth first over the successors, like extendPathEnd)
        std::vector<KP> kps;
        std::vector<int32_t> lists; // site-local node lists of ALL k-mer paths
        std::vector<int32_t> stack_nodes((size_t)k + 2);
        struct Frame
        {
            int depth, end, ext, next_succ;
        };
        std::vector<Frame> st;
        for (int v0 = 0; v0 < n; ++v0)
            for (int pos = 0; pos < node_len[v0]; ++pos)
            {
                stack_nodes[0] = v0;
                st.clear();
                st.push_back({ 1, pos, k - 1, 0 });
                while (!st.empty())
                {
                    Frame& f = st.back();
                    const int last = stack_nodes[(size_t)f.depth - 1];
                    const int room = node_len[last] - f.end - 1;
                    if (f.ext <= room)
                    {
                        KP kp;
                        kp.start_pos = pos;
                        kp.end_pos = f.end + f.ext;
                        kp.list_off = (int32_t)lists.size();
                        kp.n_nodes = f.depth;
                        lists.insert(lists.end(), stack_nodes.begin(), stack_nodes.begin() + f.depth);
                        uint64_t h = 0;
                        for (int x = 0; x < f.depth; ++x)
                        {
                            const int nd = stack_nodes[(size_t)x];
                            const int a = x == 0 ? pos : 0, b = x == f.depth - 1 ? kp.end_pos : node_len[nd] - 1;
                            for (int p = a; p <= b; ++p)
                                h = path_hash_step(h, raw[node_start[nd] + p]);
                        }
                        kp.h = h ? h : 1;
                        kps.push_back(kp);
                        st.pop_back();
                        continue;
                    }
                    const auto& sv = succ[(size_t)last];
                    if (f.next_succ >= (int)sv.size())
                    {
                        st.pop_back();
                        continue;
                    }
                    const int c = sv[(size_t)f.next_succ++];
                    const Frame nf = { f.depth + 1, 0, f.ext - room - 1, 0 };
                    stack_nodes[(size_t)f.depth] = c;
                    st.push_back(nf);
                }
            }
        // group equal k-mers: sort by hash, compare the characters inside a run of equal hashes
        auto kmer_char = [&](const KP& kp, int j) -> uint8_t {
            for (int x = 0; x < kp.n_nodes; ++x)
            {
                const int nd = lists[(size_t)kp.list_off + (size_t)x];
                const int a = x == 0 ? kp.start_pos : 0, b = x == kp.n_nodes - 1 ? kp.end_pos : node_len[nd] - 1;
                if (j <= b - a)
                    return raw[node_start[nd] + a + j];
                j -= b - a + 1;
            }
            return 0;
        };
        auto same_kmer = [&](const KP& a, const KP& b) {
            for (int j = 0; j < k; ++j)
                if (kmer_char(a, j) != kmer_char(b, j))
                    return false;
            return true;
        };
        std::vector<uint32_t> order(kps.size());
        for (size_t i = 0; i < order.size(); ++i)
            order[i] = (uint32_t)i;
        std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return kps[a].h < kps[b].h; });
        std::vector<uint32_t> uniq;
        for (size_t i = 0; i < order.size();)
        {
            size_t j = i;
            while (j < order.size() && kps[order[j]].h == kps[order[i]].h)
                ++j;
            // run [i, j): usually one string; count occurrences of each distinct string
            std::vector<char> done(j - i, 0);
            for (size_t a = i; a < j; ++a)
            {
                if (done[a - i])
                    continue;
                int cnt = 1;
                for (size_t b = a + 1; b < j; ++b)
                    if (!done[b - i] && same_kmer(kps[order[a]], kps[order[b]]))
                    {
                        done[b - i] = 1;
                        ++cnt;
                    }
                if (cnt == 1)
                    uniq.push_back(order[a]);
            }
            i = j;
        }
        size_t cap = 8;
        while (cap < 2 * uniq.size() + 1)
            cap <<= 1;
        ps.table_off = (int32_t)out.table.size();
        ps.table_mask = (int32_t)(cap - 1);
        ps.lists_off = (int32_t)out.lists.size();
        PathEntry empty;
        memset(&empty, 0, sizeof empty);
        out.table.resize(out.table.size() + cap, empty);
        PathEntry* tab = out.table.data() + ps.table_off;
        for (uint32_t id : uniq)
        {
            const KP& kp = kps[id];
            PathEntry e;
            e.key_lo = (uint32_t)kp.h;
            e.key_hi = (uint32_t)(kp.h >> 32);
            e.start_pos = kp.start_pos;
            e.end_pos = kp.end_pos;
            e.n_nodes = kp.n_nodes;
            e.nodes_off = (int32_t)(out.lists.size() - (size_t)ps.lists_off);
            out.lists.insert(out.lists.end(), lists.begin() + kp.list_off, lists.begin() + kp.list_off + kp.n_nodes);
            uint32_t slot = (uint32_t)(kp.h ^ (kp.h >> 29)) & (uint32_t)ps.table_mask;
            while (tab[slot].n_nodes != 0)
                slot = (slot + 1) & (uint32_t)ps.table_mask;
            tab[slot] = e;
        }
        out.sites.push_back(ps);
    }
}

// per-read scratch sizes in 32-bit words (see pg_core.cuh "per-task scratch layout")
inline size_t last_words(int max_nodes, int R, int W) { return (size_t)max_nodes * 2 * R * W; }
inline size_t ckpt_words(int max_G, int R, int W)
{
    return (size_t)num_ckpt(max_G, W) * (is_wide(R, W) ? 2 * R + 2 : R + 1) * W;
}

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
