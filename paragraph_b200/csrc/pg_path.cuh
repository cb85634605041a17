// paragraph_b200 -- the exact-match stage in front of the DP: grm::PathAligner
// (src/c++/lib/grm/PathAligner.cpp:75-164; first stage of grm::CompositeAligner, lib/grm/CompositeAligner.cpp:90-107).
// Shared verbatim between the sm_100a kernel (pg_kernels.cu: pg_path_kernel) and the CPU emulator (tests/emu).
//
// Reference algorithm: for both strands of the read, slide over its k-mers (k = 32); a k-mer that occurs on exactly
// one path of the graph (graph-tools KmerIndex, src/graphalign/KmerIndex.cpp:85-125) anchors an exact match that is
// extended greedily in both directions (extendPath{End,Start}Matching, src/graphcore/PathOperations.cpp:117-271: walk
// along the node while characters agree; at a node end take the neighbour with the strictly longest common
// prefix/suffix, compared over min(neighbour lengths) characters; stop on a tie or on no match); the scan resumes
// one past the end of the match.  The read is MAPPED if some match spans the whole read: the first one (forward strand
// first) gives graph_pos, score = read length, CIGAR "<overlap>M" per node, reverse strand flag; unique / mapq 60
// iff it is the only full-length match.
//
// Here: the host (pg_host.hpp: build_path_index) enumerates the graph's k-mer paths once per site and keeps the
// UNIQUE ones in an open-addressing table keyed by a 64-bit polynomial hash, each with its start position and its node
// list; a lookup verifies the k characters along that node list, so hash collisions cannot produce a false anchor.
// One thread per read scans both strands (rolling hash), extends, counts matches; the first full-length match is
// walked a second time to write its op words (node << 16 | overlap << 3 | OP_M) at their final place in the arena.
#pragma once
#include "pg_core.cuh"

namespace pg
{

constexpr uint64_t PATH_HASH_B = 0x9E3779B97F4A7C15ull; // odd multiplier of the polynomial hash (mod 2^64)

struct PathEntry // one unique k-mer of a site
{
    uint32_t key_lo, key_hi; // hash of the k characters; key_hi | key_lo == 0 never occurs (host adds 1 on zero)
    int32_t start_pos;       // position in the first node
    int32_t nodes_off;       // first entry of the node list in PathSite::node_lists
    int32_t n_nodes;         // > 0: nodes on the path; <= 0 with a key: a k-mer that is NOT unique (device-built tables
                             // keep such slots occupied so that probing continues past them); key 0 = empty slot
    int32_t end_pos;         // position in the last node (inclusive)
};

// Device-resident index of one site (built by pg_host.hpp); all arrays live in device (or emulator) memory.
struct PathSite
{
    int32_t k;               // k-mer length the index was built for (0 = site has no index)
    int32_t table_off;       // first PathEntry of this site's table
    int32_t table_mask;      // capacity - 1 (capacity is a power of two)
    int32_t lists_off;       // offset of this site's node lists
    int32_t raw_off;         // byte offset of the RAW node characters, concatenated in node order
    int32_t succ_ptr_off;    // int offset: succ_ptr[n+1], succ_idx[] (ascending ids); predecessors come from GraphView
};

struct PathView
{
    const PathEntry* table;
    int32_t mask;
    const int32_t* lists;
    const uint8_t* raw;        // raw characters of column c (forward orientation)
    const int32_t* node_start; // forward GraphView tables
    const int32_t* node_len;
    const int32_t* pred_ptr;
    const int32_t* pred_idx;
    const int32_t* succ_ptr;
    const int32_t* succ_idx;
    int32_t k;
};

PG_HD uint64_t path_hash_step(uint64_t h, uint8_t c) { return h * PATH_HASH_B + (uint64_t)c + 1u; }
PG_HD uint64_t path_hash_pow(int k) // B^(k-1)
{
    uint64_t p = 1;
    for (int i = 1; i < k; ++i)
        p *= PATH_HASH_B;
    return p;
}

// character j of the read as the stage sees it: the bases as given, or graphtools::reverseComplement of them
// (case-sensitive, anything but ACGT -> 'N': SequenceOperations.cpp:66-89)
PG_HD uint8_t path_read_char(const uint8_t* bases, int L, int strand, int j)
{
    return strand ? complement_base(bases[L - 1 - j]) : bases[j];
}

struct PathMatch
{
    int qpos, plen;            // start in the query, length of the extended match
    int start_node, start_pos; // Path::nodeIds().front(), Path::startPosition()
    int n_before, n_nodes;     // nodes prepended by the start extension, nodes in total
};

// Does the k-mer at q[pos..pos+k) sit on the entry's path?  (collision check + nothing else: the entry is the only
// path of the graph spelling its k-mer)
PG_HD bool path_entry_matches(const PathView& v, const PathEntry& e, const uint8_t* bases, int L, int strand, int pos)
{
    int j = pos;
    for (int x = 0; x < e.n_nodes; ++x)
    {
        const int node = v.lists[e.nodes_off + x];
        const int a = x == 0 ? e.start_pos : 0, b = x == e.n_nodes - 1 ? e.end_pos : v.node_len[node] - 1;
        const uint8_t* s = v.raw + v.node_start[node];
        for (int p = a; p <= b; ++p, ++j)
            if (s[p] != path_read_char(bases, L, strand, j))
                return false;
    }
    return j == pos + v.k;
}

PG_HD const PathEntry* path_lookup(const PathView& v, uint64_t h, const uint8_t* bases, int L, int strand, int pos)
{
    if (h == 0)
        h = 1;
    const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    for (uint32_t slot = (uint32_t)(h ^ (h >> 29)) & (uint32_t)v.mask;; slot = (slot + 1) & (uint32_t)v.mask)
    {
        const PathEntry& e = v.table[slot];
        if ((e.key_lo | e.key_hi) == 0u)
            return nullptr;
        if (e.n_nodes > 0 && e.key_lo == lo && e.key_hi == hi && path_entry_matches(v, e, bases, L, strand, pos))
            return &e;
    }
}

// extendPathEndMatching + extendPathStartMatching (PathOperations.cpp:117-271) from the anchor `e` found at query
// position pos.  ops: when non-null, the op words of the extended path are written to ops[0 .. m.n_nodes) -- the caller
// passes the PathMatch of a first (counting) pass so that the prepended nodes land at their final index.
PG_HD void path_extend(const PathView& v, const PathEntry& e, const uint8_t* bases, int L, int strand, int pos,
                       PathMatch& m, uint32_t* ops, const PathMatch* first_pass)
{
    // ---- end extension ----
    int pos_in_query = pos + v.k;
    int node = v.lists[e.nodes_off + e.n_nodes - 1];
    int pos_in_node = e.end_pos + 1;
    int n_after = 0;
    const int base = first_pass ? first_pass->n_before : 0; // index of the anchor's first node in ops
    if (ops)
        for (int x = 0; x < e.n_nodes; ++x)
            ops[base + x] = (uint32_t)v.lists[e.nodes_off + x] << 16; // lengths are filled in below
    bool moved = true;
    while (moved)
    {
        moved = false;
        const uint8_t* s = v.raw + v.node_start[node];
        const int nl = v.node_len[node];
        while (pos_in_query < L && pos_in_node < nl && path_read_char(bases, L, strand, pos_in_query) == s[pos_in_node])
        {
            moved = true;
            ++pos_in_node;
            ++pos_in_query;
        }
        if (pos_in_node >= nl)
        {
            int num_longest = 0, longest = 0, best = 0, min_size = 0x7fffffff;
            for (int x = v.succ_ptr[node]; x < v.succ_ptr[node + 1]; ++x)
                min_size = v.node_len[v.succ_idx[x]] < min_size ? v.node_len[v.succ_idx[x]] : min_size;
            for (int x = v.succ_ptr[node]; x < v.succ_ptr[node + 1]; ++x)
            {
                const int c = v.succ_idx[x];
                const uint8_t* cs = v.raw + v.node_start[c];
                int p = 0;
                while (p < min_size && pos_in_query + p < L && cs[p] == path_read_char(bases, L, strand, pos_in_query + p))
                    ++p;
                if (p > longest)
                {
                    longest = p;
                    best = c;
                    num_longest = 1;
                }
                else if (p == longest)
                    ++num_longest;
            }
            if (longest == 0 || num_longest != 1)
                break;
            if (ops)
                ops[base + e.n_nodes + n_after] = (uint32_t)best << 16;
            ++n_after;
            pos_in_query += longest;
            pos_in_node = longest;
            node = best;
            moved = true;
        }
    }
    const int end_pos = pos_in_node - 1;
    const int q_end = pos_in_query;
    // ---- start extension ----
    pos_in_query = pos;
    node = v.lists[e.nodes_off];
    pos_in_node = e.start_pos;
    int n_before = 0;
    moved = true;
    while (moved)
    {
        moved = false;
        const uint8_t* s = v.raw + v.node_start[node];
        while (pos_in_query > 0 && pos_in_node > 0
               && path_read_char(bases, L, strand, pos_in_query - 1) == s[pos_in_node - 1])
        {
            moved = true;
            --pos_in_node;
            --pos_in_query;
        }
        if (pos_in_node == 0)
        {
            int num_longest = 0, longest = 0, best = 0, min_size = 0x7fffffff;
            for (int x = v.pred_ptr[node]; x < v.pred_ptr[node + 1]; ++x)
                min_size = v.node_len[v.pred_idx[x]] < min_size ? v.node_len[v.pred_idx[x]] : min_size;
            for (int x = v.pred_ptr[node]; x < v.pred_ptr[node + 1]; ++x)
            {
                const int c = v.pred_idx[x];
                const int cl = v.node_len[c];
                const uint8_t* cs = v.raw + v.node_start[c];
                int pp = cl, mlen = 0;
                while (pp > cl - min_size && pos_in_query - mlen > 0
                       && cs[pp - 1] == path_read_char(bases, L, strand, pos_in_query - mlen - 1))
                {
                    --pp;
                    ++mlen;
                }
                if (mlen > longest)
                {
                    longest = mlen;
                    best = c;
                    num_longest = 1;
                }
                else if (mlen == longest)
                    ++num_longest;
            }
            if (longest == 0 || num_longest != 1)
                break;
            ++n_before;
            if (ops)
                ops[base - n_before] = (uint32_t)best << 16;
            pos_in_query -= longest;
            node = best;
            pos_in_node = v.node_len[node] - longest;
            moved = true;
        }
    }
    m.qpos = pos_in_query;
    m.plen = q_end - pos_in_query;
    m.start_node = node;
    m.start_pos = pos_in_node;
    m.n_before = n_before;
    m.n_nodes = n_before + e.n_nodes + n_after;
    if (ops) // overlap of the path with each of its nodes (Path::getOverlapLength, Path.cpp:238-262) as "<len>M"
        for (int x = 0; x < m.n_nodes; ++x)
        {
            const int nd = (int)(ops[x] >> 16);
            int ov = v.node_len[nd];
            if (m.n_nodes == 1)
                ov = end_pos - m.start_pos + 1;
            else if (x == 0)
                ov = v.node_len[nd] - m.start_pos;
            else if (x == m.n_nodes - 1)
                ov = end_pos + 1;
            ops[x] = cigar_word(nd, OP_M, ov);
        }
}

struct PathResult
{
    int n_matches; // anchors found (PathAligner's anchored_ counts reads with >= 1)
    int n_full;    // matches spanning the whole read
    int strand, seed_pos; // of the first full-length match
    uint64_t seed_hash;
    PathMatch first;
};

// One strand of PathAligner::alignRead's scan (PathAligner.cpp:93-108).  q = the L characters of that strand (the
// bases, or their reverse complement: the caller prepares them so that the scan reads plain characters), nothing is
// written.  The k-mer hashes of PATH_BATCH consecutive positions are rolled and their table slots fetched together
// (independent loads in flight instead of one dependent probe after the other); the slots are then examined in order,
// because a match makes the scan jump past its end.
constexpr int PATH_BATCH = 8;
PG_HD void path_scan_strand(const PathView& v, const uint8_t* q, int L, int strand, PathResult& r)
{
    r.n_matches = r.n_full = 0;
    r.strand = strand;
    r.seed_pos = 0;
    r.seed_hash = 0;
    const int k = v.k;
    if (k <= 0 || L < k)
        return;
    const uint64_t top = path_hash_pow(k);
    uint64_t hprev = 0;
    int have = -1; // position whose k-mer hash hprev holds
    int pos = 0;
    while (pos + k <= L)
    {
        const int nb = (L - k - pos + 1) < PATH_BATCH ? (L - k - pos + 1) : PATH_BATCH;
        uint64_t h[PATH_BATCH];
        int32_t nn[PATH_BATCH];
        if (have == pos - 1 && pos > 0) // roll on from the previous batch
            h[0] = path_hash_step(hprev - ((uint64_t)q[pos - 1] + 1u) * top, q[pos + k - 1]);
        else
        {
            uint64_t x = 0;
            for (int j = 0; j < k; ++j)
                x = path_hash_step(x, q[pos + j]);
            h[0] = x;
        }
PG_UNROLL
        for (int j = 1; j < PATH_BATCH; ++j)
            h[j] = j < nb ? path_hash_step(h[j - 1] - ((uint64_t)q[pos + j - 1] + 1u) * top, q[pos + j + k - 1]) : 0;
PG_UNROLL
        for (int j = 0; j < PATH_BATCH; ++j) // first probe slot of every position of the batch
        {
            const uint64_t hj = h[j] ? h[j] : 1;
            const PathEntry& e0 = v.table[(uint32_t)(hj ^ (hj >> 29)) & (uint32_t)v.mask];
            nn[j] = (j < nb && (e0.key_lo | e0.key_hi) != 0u) ? 1 : 0;
        }
        bool jumped = false;
        for (int j = 0; j < nb; ++j)
        {
            if (nn[j] == 0) // empty first slot: this k-mer is on no path at all
                continue;
            const uint64_t hj = h[j] ? h[j] : 1;
            // occupied: the full lookup (key compare, collision check along the entry's nodes, further probing)
            const PathEntry* e = path_lookup(v, hj, q, L, 0, pos + j);
            if (!e)
                continue;
            PathMatch m;
            path_extend(v, *e, q, L, 0, pos + j, m, nullptr, nullptr);
            ++r.n_matches;
            if (m.plen == L)
            {
                if (r.n_full == 0)
                {
                    r.seed_pos = pos + j;
                    r.seed_hash = hj;
                    r.first = m;
                }
                ++r.n_full;
            }
            pos = m.qpos + m.plen + 1; // PathAligner.cpp:106 and the loop's ++pos
            have = -1;
            jumped = true;
            break;
        }
        if (!jumped)
        {
            hprev = h[nb - 1];
            have = pos + nb - 1;
            pos += nb;
        }
    }
}

// both strands in the reference's order: forward first, so a forward full-length match is "the first one"
PG_HD void path_combine(const PathResult& fwd, const PathResult& rev, PathResult& r)
{
    r = fwd.n_full > 0 ? fwd : rev;
    r.n_matches = fwd.n_matches + rev.n_matches;
    r.n_full = fwd.n_full + rev.n_full;
}

// the characters of one strand as the stage sees them (see path_read_char)
PG_HD void path_strand_chars(const uint8_t* bases, int L, int strand, uint8_t* q)
{
    for (int j = 0; j < L; ++j)
        q[j] = path_read_char(bases, L, strand, j);
}

// Second walk of the first full-length match: writes its r.first.n_nodes op words to ops[0 .. n_nodes).
// q = the characters of strand r.strand (path_strand_chars)
PG_HD void path_emit(const PathView& v, const uint8_t* q, int L, const PathResult& r, uint32_t* ops)
{
    const PathEntry* e = path_lookup(v, r.seed_hash, q, L, 0, r.seed_pos);
    PathMatch m;
    path_extend(v, *e, q, L, 0, r.seed_pos, m, ops, &r.first);
}

// The record of a read the stage mapped (PathAligner.cpp:124-161)
PG_HD void path_record(const PathResult& r, int L, Record& rec)
{
    rec.graph_pos = r.first.start_pos;         // path.startPosition()
    rec.score = (int16_t)L;                    // path.length()
    rec.query_clipped = 0;
    rec.unique = (uint8_t)(r.n_full == 1);     // a second full-length match -> not unique, mapq 0
    rec.chose_reverse = (uint8_t)r.strand;     // bases := reverseComplement(bases), is_graph_reverse_strand := true
    rec.status = 0;
    rec.mapped_by = (uint8_t)STAGE_PATH;
    rec.cigar_off = 0;
    rec.cigar_len = (uint32_t)r.first.n_nodes;
}

// view of one site's index over the uploaded (or host) arrays
PG_HD PathView make_path_view(const PathSite& ps, const SiteDev& sd, const PathEntry* table, const int32_t* lists,
                              const int32_t* succ, const uint8_t* bytes, const int32_t* ints)
{
    PathView v;
    v.table = table + ps.table_off;
    v.mask = ps.table_mask;
    v.lists = lists + ps.lists_off;
    v.raw = bytes + ps.raw_off;
    const int32_t* t = ints + sd.tab_off[0];
    v.node_start = t;
    v.node_len = t + sd.n_nodes;
    v.pred_ptr = t + 2 * sd.n_nodes;
    v.pred_idx = t + 3 * sd.n_nodes + 1;
    v.succ_ptr = succ + ps.succ_ptr_off;
    v.succ_idx = v.succ_ptr + sd.n_nodes + 1;
    v.k = ps.k;
    return v;
}

} // namespace pg
