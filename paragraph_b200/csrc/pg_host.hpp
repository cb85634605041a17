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
#include <thread>
#include <unordered_map>
#include <vector>

#include "pg_core.cuh"
#include "pg_kmer.cuh"
#include "pg_count.cuh"
#include "pg_path.cuh"

namespace pg
{
namespace host
{

struct CodeTable // per input character: gssw column code and the upper-cased character (pg_core.cuh: nt_code, to_upper)
{
    uint8_t code[256], upper[256];
    CodeTable()
    {
        for (int ch = 0; ch < 256; ++ch)
        {
            upper[ch] = to_upper((uint8_t)ch);
            code[ch] = (uint8_t)nt_code(upper[ch]);
        }
    }
};

struct GraphStore
{
    std::vector<SiteDev> sites;
    std::vector<uint8_t> bytes;
    std::vector<int32_t> ints;
    int max_nodes = 0;
    int max_G = 0;
    int max_tab_ints = 0; // largest per-orientation int table (SiteDev::tab_ints)
    std::vector<int32_t> scratch_deg, scratch_idx, scratch_fill; // add(): reused between sites
    // for the counting stage (pg_count.cuh): the edges as given (input order is the order of the edge count rows),
    // their path-family label masks (pg_set_edge_labels; 0 = unlabelled) and each site's first row
    std::vector<int32_t> in_from, in_to;
    std::vector<uint64_t> in_label;
    std::vector<int64_t> edge_base, node_base; // [n_sites + 1]
    // the paths of each site's graph JSON (GraphInput.cpp:168-197; pg_set_paths): what grm::KmerAligner aligns to
    std::vector<std::vector<std::vector<int32_t>>> paths; // [site][path] = node ids

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
        paths.clear();
    }

    // sizes of everything add() appends to: pg_add_graphs registers all of its sites or none
    struct Mark
    {
        size_t sites, bytes, ints, in_edges, bases;
        int max_nodes, max_G, max_tab_ints;
    };
    Mark mark() const { return Mark{ sites.size(), bytes.size(), ints.size(), in_from.size(), edge_base.size(), max_nodes, max_G, max_tab_ints }; }
    void rollback(const Mark& m)
    {
        sites.resize(m.sites);
        bytes.resize(m.bytes);
        ints.resize(m.ints);
        in_from.resize(m.in_edges);
        in_to.resize(m.in_edges);
        in_label.resize(m.in_edges);
        edge_base.resize(m.bases);
        node_base.resize(m.bases);
        paths.resize(m.sites);
        max_nodes = m.max_nodes;
        max_G = m.max_G;
        max_tab_ints = m.max_tab_ints;
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
        // forward characters (upper-cased), concatenated in node order.  (Everything below is written through pointers
        // into space reserved once per site: a 10 000-site sweep registers its graphs inside the timed pass.)
        static const CodeTable lut;
        const uint8_t* raw = (const uint8_t*)blob + base0;
        const size_t tail = (size_t)SENT + CK + 4 + 16; // + slack so the span can be rounded up to 16 bytes
        size_t pos = (bytes.size() + 15) & ~(size_t)15; // the staged span [codes - SENT, ...) must be 16-byte aligned (TMA)
        size_t need = pos;
        for (int o = 0; o < 2; ++o)
            need = ((need + 15) & ~(size_t)15) + SENT + (size_t)G + tail;
        need += 2 * (size_t)G;
        bytes.resize(need, 0);
        uint8_t* B = bytes.data();
        for (int o = 0; o < 2; ++o)
        {
            pos = (pos + 15) & ~(size_t)15;
            memset(B + pos, 5, SENT);
            pos += SENT;
            sd.codes_off[o] = (int32_t)pos;
            if (o == 0)
                for (int64_t x = 0; x < G; ++x)
                    B[pos + (size_t)x] = lut.code[raw[x]];
            else
                for (int64_t x = 0; x < G; ++x)
                    B[pos + (size_t)x] = lut.code[raw[G - 1 - x]];
            pos += (size_t)G;
            memset(B + pos, 5, tail);
            pos += tail;
        }
        sd.chars_off = (int32_t)pos;
        for (int64_t x = 0; x < G; ++x)
            B[pos + (size_t)x] = lut.upper[raw[x]];
        pos += (size_t)G;
        sd.raw_off = (int32_t)pos; // as given: the exact-match stage compares characters case-sensitively
        memcpy(B + pos, raw, (size_t)G);
        const int tab_ints = 3 * n_nodes + 1 + n_edges;
        const size_t ints0 = ints.size();
        ints.resize(ints0 + 2 * (size_t)tab_ints, 0);
        scratch_deg.assign((size_t)n_nodes + 1, 0);
        for (int o = 0; o < 2; ++o)
        {
            sd.tab_off[o] = (int32_t)(ints0 + (size_t)o * tab_ints);
            int32_t* start = ints.data() + sd.tab_off[o];
            int32_t* len = start + n_nodes;
            int32_t* ptr = len + n_nodes;
            int32_t* idx = ptr + n_nodes + 1;
            int32_t acc = 0;
            for (int i = 0; i < n_nodes; ++i)
            {
                const int src = o == 0 ? i : n_nodes - 1 - i;
                len[i] = off[src + 1] - off[src];
                start[i] = acc;
                acc += len[i];
            }
            // predecessor lists: counting sort of the edges by target, each list ascending and without duplicates
            // (Graph::addEdge rejects duplicates anyway)
            std::fill(scratch_deg.begin(), scratch_deg.end(), 0);
            for (int e = 0; e < n_edges; ++e)
                ++scratch_deg[(size_t)(o == 0 ? et[e] : n_nodes - 1 - ef[e]) + 1];
            for (int i = 0; i < n_nodes; ++i)
                scratch_deg[(size_t)i + 1] += scratch_deg[(size_t)i];
            scratch_idx.assign((size_t)n_edges, 0);
            scratch_fill.assign(scratch_deg.begin(), scratch_deg.end() - 1);
            for (int e = 0; e < n_edges; ++e)
            {
                const int f = o == 0 ? ef[e] : n_nodes - 1 - et[e];
                const int t = o == 0 ? et[e] : n_nodes - 1 - ef[e];
                scratch_idx[(size_t)scratch_fill[(size_t)t]++] = f;
            }
            int32_t out = 0;
            for (int i = 0; i < n_nodes; ++i)
            {
                ptr[i] = out;
                int32_t* b = scratch_idx.data() + scratch_deg[(size_t)i];
                int32_t* e = scratch_idx.data() + scratch_deg[(size_t)i + 1];
                std::sort(b, e);
                e = std::unique(b, e);
                for (; b < e; ++b)
                    idx[out++] = *b;
            }
            ptr[n_nodes] = out; // (pred_idx is padded with zeros to n_edges entries: both orientations occupy the same space)
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
        paths.emplace_back();
        return (int)sites.size() - 1;
    }

    // graphtools::Path(graph, 0, nodes, last node length - 1) for every path (GraphInput.cpp:168-197): node ids must
    // exist and consecutive nodes must be joined by an edge (Path::assertValidity)
    bool set_paths(int site, int n_paths, const int32_t* path_ptr, const int32_t* path_nodes, std::string& err)
    {
        if (site < 0 || (size_t)site >= sites.size() || n_paths < 0 || (n_paths > 0 && (!path_ptr || !path_nodes)))
        {
            err = "pg_set_paths: bad arguments";
            return false;
        }
        const SiteDev& sd = sites[(size_t)site];
        const int64_t eb = edge_base[(size_t)site];
        std::vector<std::vector<int32_t>> out;
        for (int p = 0; p < n_paths; ++p)
        {
            std::vector<int32_t> nodes(path_nodes + path_ptr[p], path_nodes + path_ptr[p + 1]);
            if (nodes.empty())
            {
                err = "pg_set_paths: empty path";
                return false;
            }
            for (size_t i = 0; i < nodes.size(); ++i)
            {
                bool ok = nodes[i] >= 0 && nodes[i] < sd.n_nodes;
                if (ok && i > 0)
                {
                    ok = false;
                    for (int e = 0; e < sd.n_edges && !ok; ++e)
                        ok = in_from[(size_t)(eb + e)] == nodes[i - 1] && in_to[(size_t)(eb + e)] == nodes[i];
                }
                if (!ok)
                {
                    err = "pg_set_paths: path " + std::to_string(p) + " is not a path of the graph";
                    return false;
                }
            }
            out.push_back(std::move(nodes));
        }
        paths[(size_t)site] = std::move(out);
        return true;
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

// one site's share of the index, with site-local offsets
struct PathSiteBuild
{
    std::vector<int32_t> succ;      // succ_ptr[n+1], succ_idx[]
    std::vector<PathEntry> entries; // the unique k-mers, not yet placed in the table
    std::vector<int32_t> lists;     // node lists of the kept entries
};

// successors of every node of a site, ascending, without duplicates; csr = succ_ptr[n+1], succ_idx[]
inline void build_successors(const GraphStore& gs, size_t si, std::vector<std::vector<int32_t>>& succ,
                             std::vector<int32_t>& csr)
{
    const int n = gs.sites[si].n_nodes;
    succ.assign((size_t)n, std::vector<int32_t>());
    for (int64_t e = gs.edge_base[si]; e < gs.edge_base[si + 1]; ++e)
        succ[(size_t)gs.in_from[(size_t)e]].push_back(gs.in_to[(size_t)e]);
    int32_t ptr = 0;
    for (int i = 0; i < n; ++i)
    {
        auto& v = succ[(size_t)i];
        std::sort(v.begin(), v.end());
        v.erase(std::unique(v.begin(), v.end()), v.end());
        csr.push_back(ptr);
        ptr += (int32_t)v.size();
    }
    csr.push_back(ptr);
    for (int i = 0; i < n; ++i)
        csr.insert(csr.end(), succ[(size_t)i].begin(), succ[(size_t)i].end());
}

inline void build_path_site(const GraphStore& gs, size_t si, int k, PathSiteBuild& out)
{
    struct KP
    {
        uint64_t h;
        int32_t v0, start_pos, end_pos, list_off, n_nodes; // list_off < 0: the path stays inside node v0
    };
    const SiteDev& sd = gs.sites[si];
    const int n = sd.n_nodes;
    const int32_t* t = gs.ints.data() + sd.tab_off[0];
    const int32_t *node_start = t, *node_len = t + n;
    const uint8_t* raw = gs.bytes.data() + sd.raw_off;
    std::vector<std::vector<int32_t>> succ;
    build_successors(gs, si, succ, out.succ);
    // enumerate the k-mer paths: inside a node by rolling the hash along it, across node ends depth first over the
    // successors like extendPathEnd
    std::vector<KP> kps;
    kps.reserve((size_t)sd.G + 16);
    std::vector<int32_t> lists; // node lists of the multi-node k-mer paths
    std::vector<int32_t> stack_nodes((size_t)k + 2);
    struct Frame
    {
        int depth, end, ext, next_succ;
    };
    std::vector<Frame> st;
    const uint64_t top = path_hash_pow(k);
    for (int v0 = 0; v0 < n; ++v0)
    {
        const uint8_t* sq = raw + node_start[v0];
        const int len = node_len[v0];
        uint64_t h = 0;
        for (int pos = 0; pos + k <= len; ++pos) // k-mers inside the node
        {
            if (pos == 0)
                for (int j = 0; j < k; ++j)
                    h = path_hash_step(h, sq[j]);
            else
                h = path_hash_step(h - ((uint64_t)sq[pos - 1] + 1u) * top, sq[pos + k - 1]);
            KP kp = { h ? h : 1, v0, pos, pos + k - 1, -1, 1 };
            kps.push_back(kp);
        }
        for (int pos = std::max(0, len - k + 1); pos < len; ++pos) // k-mers that leave the node
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
                    kp.v0 = v0;
                    kp.start_pos = pos;
                    kp.end_pos = f.end + f.ext;
                    kp.list_off = (int32_t)lists.size();
                    kp.n_nodes = f.depth;
                    lists.insert(lists.end(), stack_nodes.begin(), stack_nodes.begin() + f.depth);
                    uint64_t hh = 0;
                    for (int x = 0; x < f.depth; ++x)
                    {
                        const int nd = stack_nodes[(size_t)x];
                        const int a = x == 0 ? pos : 0, b = x == f.depth - 1 ? kp.end_pos : node_len[nd] - 1;
                        for (int p = a; p <= b; ++p)
                            hh = path_hash_step(hh, raw[node_start[nd] + p]);
                    }
                    kp.h = hh ? hh : 1;
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
    }
    // group equal k-mers: sort by hash, compare the characters inside a run of equal hashes
    auto kmer_char = [&](const KP& kp, int j) -> uint8_t {
        if (kp.list_off < 0)
            return raw[node_start[kp.v0] + kp.start_pos + j];
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
    // count the occurrences of every distinct k-mer with a scratch open-addressing table over ALL k-mer paths (key =
    // hash, verified by comparing the characters with the first path that took the slot); the unique ones are kept
    std::vector<uint32_t> uniq;
    {
        size_t cap2 = 16;
        while (cap2 < 2 * kps.size() + 2)
            cap2 <<= 1;
        std::vector<uint32_t> first(cap2, 0xFFFFFFFFu); // index of the first path with this k-mer
        std::vector<uint32_t> count(cap2, 0);
        for (uint32_t id = 0; id < (uint32_t)kps.size(); ++id)
        {
            const uint64_t h = kps[id].h;
            for (size_t slot = (size_t)(h ^ (h >> 29)) & (cap2 - 1);; slot = (slot + 1) & (cap2 - 1))
            {
                if (first[slot] == 0xFFFFFFFFu)
                {
                    first[slot] = id;
                    count[slot] = 1;
                    break;
                }
                if (kps[first[slot]].h == h && same_kmer(kps[first[slot]], kps[id]))
                {
                    ++count[slot];
                    break;
                }
            }
        }
        uniq.reserve(kps.size());
        for (size_t slot = 0; slot < cap2; ++slot)
            if (count[slot] == 1)
                uniq.push_back(first[slot]);
        std::sort(uniq.begin(), uniq.end()); // enumeration order: deterministic tables
    }
    out.lists.reserve(uniq.size() + lists.size());
    out.entries.reserve(uniq.size());
    for (uint32_t id : uniq)
    {
        const KP& kp = kps[id];
        PathEntry e;
        e.key_lo = (uint32_t)kp.h;
        e.key_hi = (uint32_t)(kp.h >> 32);
        e.start_pos = kp.start_pos;
        e.end_pos = kp.end_pos;
        e.n_nodes = kp.n_nodes;
        e.nodes_off = (int32_t)out.lists.size();
        if (kp.list_off < 0)
            out.lists.push_back(kp.v0);
        else
            out.lists.insert(out.lists.end(), lists.begin() + kp.list_off, lists.begin() + kp.list_off + kp.n_nodes);
        out.entries.push_back(e);
    }
}

// open addressing with linear probing into a zeroed power-of-two block, load factor <= 0.5: most lookups are misses
// (every k-mer of the strand that does not match), and a miss walks to the next empty slot -- 1.6 probes on average at
// load 0.3 but 4 at 0.6, which is what the stage's kernel time follows (0.047 vs 0.080 ms per 10k reads on the B200)
inline size_t path_table_cap(size_t n_entries)
{
    size_t cap = 8;
    while (cap < 2 * n_entries + 1)
        cap <<= 1;
    return cap;
}
inline void place_path_entries(const std::vector<PathEntry>& entries, PathEntry* tab, size_t cap)
{
    for (const PathEntry& e : entries)
    {
        const uint64_t h = ((uint64_t)e.key_hi << 32) | e.key_lo;
        uint32_t slot = (uint32_t)(h ^ (h >> 29)) & (uint32_t)(cap - 1);
        while ((tab[slot].key_lo | tab[slot].key_hi) != 0u)
            slot = (slot + 1) & (uint32_t)(cap - 1);
        tab[slot] = e;
    }
}

// ---- sizing for the DEVICE-side build (pg_kernels.cu: pg_path_index_*): how many k-mer paths a site has, and an upper
// bound of the node-list entries they need.  P[v][r] = number of ways to read r more characters starting at the first
// character of node v (extendPathEnd, PathOperations.cpp:73-103): 1 if they fit the node, else the sum over successors.
inline void count_kmer_paths(const GraphStore& gs, size_t si, int k, const int32_t* succ_ptr, const int32_t* succ_idx,
                             int64_t& n_paths, int64_t& list_ints)
{
    const SiteDev& sd = gs.sites[si];
    const int n = sd.n_nodes;
    const int32_t* node_len = gs.ints.data() + sd.tab_off[0] + n;
    std::vector<double> P((size_t)n * (size_t)k, 0.0); // [v][r], r = 1 .. k-1 (doubles: dense graphs can overflow int64)
    for (int v = n - 1; v >= 0; --v) // successors have larger ids
        for (int r = 1; r < k; ++r)
        {
            double c = 0;
            if (r <= node_len[v])
                c = 1;
            else
                for (int x = succ_ptr[v]; x < succ_ptr[v + 1]; ++x)
                    c += P[(size_t)succ_idx[x] * (size_t)k + (size_t)(r - node_len[v])];
            P[(size_t)v * (size_t)k + (size_t)r] = c;
        }
    double paths = 0, multi = 0;
    for (int v = 0; v < n; ++v)
    {
        const int len = node_len[v];
        paths += std::max(0, len - k + 1);
        for (int pos = std::max(0, len - k + 1); pos < len; ++pos)
        {
            const int r = k - (len - pos);
            for (int x = succ_ptr[v]; x < succ_ptr[v + 1]; ++x)
                multi += P[(size_t)succ_idx[x] * (size_t)k + (size_t)r];
        }
    }
    n_paths = (int64_t)std::min(paths + multi, 4e18);
    list_ints = (int64_t)std::min(paths + multi * (double)std::min(k, n), 4e18);
}

// all sites; the per-site builds are independent and run on a few host threads
inline void build_path_index(const GraphStore& gs, int k, PathIndexHost& out, int threads = 0)
{
    out = PathIndexHost();
    out.k = k;
    const size_t ns = gs.sites.size();
    std::vector<PathSiteBuild> parts(ns);
    if (threads <= 0)
        threads = (int)std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency()));
    threads = (int)std::min<size_t>((size_t)threads, std::max<size_t>(1, ns / 8));
    if (threads <= 1)
        for (size_t si = 0; si < ns; ++si)
            build_path_site(gs, si, k, parts[si]);
    else
    {
        std::vector<std::thread> pool;
        for (int w = 0; w < threads; ++w)
            pool.emplace_back([&, w]() {
                for (size_t si = (size_t)w; si < ns; si += (size_t)threads)
                    build_path_site(gs, si, k, parts[si]);
            });
        for (auto& th : pool)
            th.join();
    }
    size_t n_table = 0, n_lists = 0, n_succ = 0;
    out.sites.resize(ns);
    for (size_t si = 0; si < ns; ++si)
    {
        PathSite& ps = out.sites[si];
        ps.k = k;
        ps.raw_off = gs.sites[si].raw_off;
        ps.succ_ptr_off = (int32_t)n_succ;
        ps.table_off = (int32_t)n_table;
        ps.table_mask = (int32_t)(path_table_cap(parts[si].entries.size()) - 1);
        ps.lists_off = (int32_t)n_lists;
        n_table += (size_t)ps.table_mask + 1;
        n_lists += parts[si].lists.size();
        n_succ += parts[si].succ.size();
    }
    out.table.resize(n_table); // value-initialised: n_nodes = 0 marks an empty slot
    out.lists.resize(n_lists);
    out.succ.resize(n_succ);
    auto fill = [&](size_t si) {
        const PathSite& ps = out.sites[si];
        place_path_entries(parts[si].entries, out.table.data() + ps.table_off, (size_t)ps.table_mask + 1);
        std::copy(parts[si].lists.begin(), parts[si].lists.end(), out.lists.begin() + ps.lists_off);
        std::copy(parts[si].succ.begin(), parts[si].succ.end(), out.succ.begin() + ps.succ_ptr_off);
    };
    if (threads <= 1)
        for (size_t si = 0; si < ns; ++si)
            fill(si);
    else
    {
        std::vector<std::thread> pool;
        for (int w = 0; w < threads; ++w)
            pool.emplace_back([&, w]() {
                for (size_t si = (size_t)w; si < ns; si += (size_t)threads)
                    fill(si);
            });
        for (auto& th : pool)
            th.join();
    }
}

// per-read scratch sizes in 32-bit words (see pg_core.cuh "per-task scratch layout")
// ---------------------------------------------------------------------------------------------
// k-mer stage (pg_kmer.cuh), host side: per path its sequence, node offsets and sorted k-mers
// (KmerAligner.cpp:135-160 BasicPath, :120-133 makeKmers)
// ---------------------------------------------------------------------------------------------
struct KmerIndexHost
{
    std::vector<KmerSiteDev> sites;
    std::vector<KmerPathDev> paths;
    std::vector<uint8_t> seqs;
    std::vector<int32_t> nodes;
    std::vector<KmerPos> kmers;
    int max_path_nodes = 0, max_paths = 0;
};
inline void build_kmer_index(const GraphStore& gs, int k, KmerIndexHost& ix)
{
    const size_t ns = gs.sites.size();
    ix.sites.resize(ns);
    for (size_t si = 0; si < ns; ++si)
    {
        const SiteDev& sd = gs.sites[si];
        const int32_t* node_start = gs.ints.data() + sd.tab_off[0];
        const int32_t* node_len = node_start + sd.n_nodes;
        const uint8_t* raw = gs.bytes.data() + sd.raw_off;
        ix.sites[si].n_paths = (int32_t)gs.paths[si].size();
        ix.sites[si].path0 = (int32_t)ix.paths.size();
        ix.max_paths = std::max(ix.max_paths, ix.sites[si].n_paths);
        for (auto const& nodes : gs.paths[si])
        {
            KmerPathDev kp;
            kp.seq_off = (int32_t)ix.seqs.size();
            kp.n_nodes = (int32_t)nodes.size();
            kp.nodes_off = (int32_t)ix.nodes.size();
            ix.max_path_nodes = std::max(ix.max_path_nodes, kp.n_nodes);
            ix.nodes.insert(ix.nodes.end(), nodes.begin(), nodes.end());
            int32_t len = 0;
            for (int32_t v : nodes)
            {
                ix.nodes.push_back(len);
                ix.seqs.insert(ix.seqs.end(), raw + node_start[v], raw + node_start[v] + node_len[v]);
                len += node_len[v];
            }
            kp.len = len;
            kp.kmers_off = (int32_t)ix.kmers.size();
            const uint8_t* s = ix.seqs.data() + kp.seq_off;
            uint32_t val = 0;
            int have = 0;
            const uint32_t mask = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
            for (int32_t i = 0; i < len; ++i)
            {
                const int b = kmer_base_value(s[i]);
                if (b > 3)
                {
                    have = 0;
                    continue;
                }
                val = (val << 2) | (uint32_t)b;
                if (++have >= k)
                    ix.kmers.push_back(KmerPos{ val & mask, i - k + 1 });
            }
            kp.n_kmers = (int32_t)ix.kmers.size() - kp.kmers_off;
            std::sort(ix.kmers.begin() + kp.kmers_off, ix.kmers.end(), [](const KmerPos& a, const KmerPos& b) {
                return a.kmer < b.kmer || (a.kmer == b.kmer && a.pos < b.pos);
            });
            ix.paths.push_back(kp);
        }
    }
}

inline size_t last_words(int max_nodes, int R, int W) // node table of one read: a row per (node, lane), pg_core.cuh Sizes
{
    const int v = is_wide(R, W) ? 2 * R : R, iw = is_wide(R, W) ? 4 : 3;
    return (size_t)max_nodes * W * (size_t)((v + iw + 3) & ~3);
}
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
