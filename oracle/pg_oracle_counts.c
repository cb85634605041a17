/* TEST INFRASTRUCTURE ONLY -- never linked into, imported by, or called from the product path.
 *
 * Plain-C restatement of the reference's read filtering + disambiguation + read counting for ONE site
 * (SURVEY.md 8f rank 1):
 *   createReadFilter chain NonUniq -> BadAlign          src/c++/lib/paragraph/ReadFilter.cpp:73-90
 *   decodeGraphAlignment / Path validity                graph-tools GraphAlignmentOperations.cpp:84-106, Path.cpp:86-145
 *   node / edge support filters (lambdas)               src/c++/lib/paragraph/Disambiguation.cpp:212-296
 *   disambiguateReads                                   src/c++/lib/paragraph/Disambiguation.cpp:82-142
 *   PathFamily::containsPath                            graph-tools PathFamily.cpp:88-106
 *   readsToFragments / Fragment::addRead (counters)     src/c++/lib/common/Fragment.cpp:33-67,141-182
 *   countNodes / countEdges / countPathFamilies         src/c++/lib/paragraph/ReadCounting.cpp:52-127
 * Pinned by tests/test_counts_oracle.py against oracle/_ref (the unmodified ReadCounting.cpp, Fragment.cpp and
 * graph-tools, see oracle/ref_counts.cpp) on fuzzed alignments, and against the reference's own golden vectors
 * (tests/golden/counts_unit.json, counts_phasing.json).
 */
#include "pg_oracle.h"

#include <stdlib.h>
#include <string.h>

#define MAX_PATH_NODES 1024

typedef struct
{
    int node;
    int matched, mismatched, missing, clipped, inserted, deleted;
} node_aln;

static int ref_len(const node_aln* a) { return a->matched + a->mismatched + a->missing + a->deleted; }
static int query_len(const node_aln* a) { return a->matched + a->mismatched + a->missing + a->inserted + a->clipped; }

/* splitGraphCigar + splitNodeCigar + Alignment::decodeCigar.  Returns the number of node alignments, -1 if the
 * string is malformed / too long. */
static int parse_graph_cigar(const char* s, node_aln* out)
{
    int n = 0;
    while (*s)
    {
        long id = 0;
        int digits = 0;
        while (*s >= '0' && *s <= '9')
        {
            id = id * 10 + (*s - '0');
            ++s;
            ++digits;
        }
        if (*s != '[' || digits == 0 || n >= MAX_PATH_NODES)
            return -1;
        ++s;
        node_aln a;
        memset(&a, 0, sizeof a);
        a.node = (int)id;
        while (*s && *s != ']')
        {
            long len = 0;
            digits = 0;
            while (*s >= '0' && *s <= '9')
            {
                len = len * 10 + (*s - '0');
                ++s;
                ++digits;
            }
            if (digits == 0)
                return -1;
            switch (*s)
            {
            case 'M': a.matched += (int)len; break;
            case 'X': a.mismatched += (int)len; break;
            case 'N': a.missing += (int)len; break;
            case 'S': a.clipped += (int)len; break;
            case 'I': a.inserted += (int)len; break;
            case 'D': a.deleted += (int)len; break;
            default: return -1;
            }
            ++s;
        }
        if (*s != ']')
            return -1;
        ++s;
        out[n++] = a;
    }
    return n;
}

typedef struct
{
    int n_nodes, n_edges;
    const int32_t* node_len;
    const int32_t *efrom, *eto;
    const uint64_t* labels;
    uint64_t *out_labels, *in_labels; /* per node: labels of outgoing / incoming edges (PathFamily in/outNodes) */
} site_graph;

static int edge_index(const site_graph* g, int a, int b)
{
    for (int e = 0; e < g->n_edges; ++e)
        if (g->efrom[e] == a && g->eto[e] == b)
            return e;
    return -1;
}

/* Path::Impl::isValid on what decodeGraphAlignment builds.  1 = the reference accepts it. */
static int path_valid(const site_graph* g, int pos, const node_aln* p, int n)
{
    if (n <= 0)
        return 0; /* empty CIGAR: alignments.back() on an empty vector in the reference (undefined) */
    for (int k = 0; k < n; ++k)
        if (p[k].node < 0 || p[k].node >= g->n_nodes)
            return 0;
    const int last_start = (n == 1) ? pos : 0;
    const int end = last_start + ref_len(&p[n - 1]) - 1;
    if (pos < 0 || pos >= g->node_len[p[0].node])
        return 0;
    if (end < 0 || end >= g->node_len[p[n - 1].node])
        return 0;
    for (int k = 0; k + 1 < n; ++k)
    {
        if (p[k].node > p[k + 1].node)
            return 0;
        if (edge_index(g, p[k].node, p[k + 1].node) < 0)
            return 0;
    }
    if (n == 1 && pos > end)
        return 0;
    return 1;
}

static int imin(int a, int b) { return a < b ? a : b; }

/* Disambiguation.cpp:212-243 */
static int node_supported(const site_graph* g, const node_aln* a, int read_len)
{
    const int half = read_len / 2;
    const int short_node = g->node_len[a->node] < half;
    const int nonmatch = a->mismatched + a->clipped;
    const int indel = a->inserted + a->deleted;
    if (short_node && (nonmatch > 0 || indel > 0))
        return 0;
    return nonmatch + indel <= half;
}

/* Disambiguation.cpp:245-296 */
static int edge_supported(const site_graph* g, const node_aln* prev, const node_aln* cur, int read_len)
{
    const int mno = read_len / 10 + 1;
    int ok = prev->matched >= imin(ref_len(prev), mno) && cur->matched >= imin(ref_len(cur), mno);
    if (ok)
        ok = query_len(prev) < ref_len(prev) * 2 && query_len(cur) < ref_len(cur) * 2;
    if (ok)
        ok = prev->matched >= imin(g->node_len[prev->node], mno) && cur->matched >= imin(g->node_len[cur->node], mno);
    return ok;
}

static void add4(pgo_count4* c, int reads, int fwd, int rev)
{
    c->fragments += 1;
    c->reads += (uint32_t)reads;
    c->fwd += (uint32_t)fwd;
    c->rev += (uint32_t)rev;
}

int pgo_count_site(
    int n_nodes, const int32_t* node_len, int n_edges, const int32_t* efrom, const int32_t* eto,
    const uint64_t* edge_labels, int n_reads, const int32_t* read_len, const int32_t* graph_pos, const uint8_t* unique,
    const char* cigars, int cigar_stride, const uint8_t* is_graph_reverse, const int32_t* fragment, int remove_nonuniq,
    double bad_align_frac, int use_support_filters, pgo_read_support* support, uint32_t* path_words, int path_cap,
    int* path_used, pgo_count4* node_counts, pgo_count4* edge_counts, uint32_t* family_words, int family_cap,
    int* family_used)
{
    site_graph g = { n_nodes, n_edges, node_len, efrom, eto, edge_labels, NULL, NULL };
    g.out_labels = calloc((size_t)n_nodes + 1, sizeof(uint64_t));
    g.in_labels = calloc((size_t)n_nodes + 1, sizeof(uint64_t));
    for (int e = 0; e < n_edges; ++e)
    {
        const uint64_t lab = edge_labels ? edge_labels[e] : 0;
        g.out_labels[efrom[e]] |= lab;
        g.in_labels[eto[e]] |= lab;
    }
    memset(node_counts, 0, sizeof(pgo_count4) * (size_t)n_nodes);
    memset(edge_counts, 0, sizeof(pgo_count4) * (size_t)n_edges);
    const int stride = 1 + n_nodes + n_edges; /* family entry: total, nodes, edges */
    int fam_n = 0, pw = 0, rc = 0;
    node_aln* path = malloc(sizeof(node_aln) * MAX_PATH_NODES);

    /* ---- per read: filter chain, then disambiguateReads */
    for (int i = 0; i < n_reads; ++i)
    {
        pgo_read_support* s = &support[i];
        memset(s, 0, sizeof *s);
        s->path_off = (uint32_t)pw;
        s->graph_reverse = is_graph_reverse ? is_graph_reverse[i] : 0;
        if (remove_nonuniq && !unique[i])
        {
            s->verdict = PGO_V_NONUNIQ;
            continue;
        }
        const char* cigar = cigars + (size_t)i * cigar_stride;
        const int n = parse_graph_cigar(cigar, path);
        if (!path_valid(&g, graph_pos[i], path, n))
        {
            s->verdict = PGO_V_INVALID; /* BadAlign's decodeGraphAlignment throws: the reference aborts the site */
            continue;
        }
        int clipped = 0;
        if (pgo_bad_align(cigar, bad_align_frac, &clipped))
        {
            s->verdict = PGO_V_BAD_ALIGN;
            continue;
        }
        s->verdict = PGO_V_MAPPED;
        s->path_len = (uint16_t)n;
        uint64_t overlapped = 0, fail = 0;
        for (int k = 0; k < n; ++k)
        {
            uint32_t w = (uint32_t)path[k].node;
            if (k > 0)
            {
                const int e = edge_index(&g, path[k - 1].node, path[k].node);
                const uint64_t lab = edge_labels ? edge_labels[e] : 0;
                if (!use_support_filters || edge_supported(&g, &path[k - 1], &path[k], read_len[i]))
                {
                    w |= PGO_SUP_EDGE;
                    overlapped |= lab;
                }
                /* PathFamily::containsPath: a path step outside the family that leaves one of its out-nodes or
                 * enters one of its in-nodes breaks it */
                fail |= ~lab & (g.out_labels[path[k - 1].node] | g.in_labels[path[k].node]);
            }
            if (!use_support_filters || node_supported(&g, &path[k], read_len[i]))
                w |= PGO_SUP_NODE;
            if (pw < path_cap)
                path_words[pw] = w;
            ++pw;
        }
        s->sequences = overlapped & ~fail;
    }
    if (pw > path_cap)
        rc = -1;

    /* ---- fragments (first-appearance order), MAPPED reads only: Align.cpp:81-84 dropped the others */
    int max_frag = -1;
    for (int i = 0; i < n_reads; ++i)
    {
        const int f = fragment ? fragment[i] : i;
        if (f < 0)
        {
            rc = -2;
            goto done;
        }
        if (f > max_frag)
            max_frag = f;
    }
    {
        int* head = malloc(sizeof(int) * (size_t)(max_frag + 2));
        int* next = malloc(sizeof(int) * (size_t)n_reads + sizeof(int));
        int* tail = malloc(sizeof(int) * (size_t)(max_frag + 2));
        uint8_t* nmark = malloc((size_t)n_nodes + 1);
        uint8_t* emark = malloc((size_t)n_edges + 1);
        for (int f = 0; f <= max_frag; ++f)
            head[f] = tail[f] = -1;
        for (int i = 0; i < n_reads; ++i)
        {
            next[i] = -1;
            if (support[i].verdict != PGO_V_MAPPED || rc == -1)
                continue;
            const int f = fragment ? fragment[i] : i;
            if (head[f] < 0)
                head[f] = i;
            else
                next[tail[f]] = i;
            tail[f] = i;
        }
        for (int i0 = 0; i0 < n_reads && rc == 0; ++i0)
        {
            const int f = fragment ? fragment[i0] : i0;
            if (support[i0].verdict != PGO_V_MAPPED || head[f] != i0)
                continue;
            int reads = 0, fwd = 0, rev = 0;
            uint64_t seqs = 0;
            memset(nmark, 0, (size_t)n_nodes + 1);
            memset(emark, 0, (size_t)n_edges + 1);
            for (int i = i0; i >= 0; i = next[i])
            {
                ++reads;
                if (support[i].graph_reverse)
                    ++rev;
                else
                    ++fwd;
                seqs |= support[i].sequences;
                const uint32_t* w = path_words + support[i].path_off;
                for (int k = 0; k < support[i].path_len; ++k)
                {
                    const int node = (int)(w[k] & PGO_SUP_NODE_MASK);
                    if (w[k] & PGO_SUP_NODE)
                        nmark[node] = 1;
                    if (k > 0 && (w[k] & PGO_SUP_EDGE))
                        emark[edge_index(&g, (int)(w[k - 1] & PGO_SUP_NODE_MASK), node)] = 1;
                }
            }
            for (int v = 0; v < n_nodes; ++v)
                if (nmark[v])
                    add4(&node_counts[v], reads, fwd, rev);
            for (int e = 0; e < n_edges; ++e)
                if (emark[e])
                    add4(&edge_counts[e], reads, fwd, rev);
            if (seqs)
            {
                int slot = -1;
                for (int q = 0; q < fam_n; ++q)
                {
                    const uint32_t* h = family_words + (size_t)q * (2 + 4 * (size_t)stride);
                    if ((((uint64_t)h[1] << 32) | h[0]) == seqs)
                        slot = q;
                }
                if (slot < 0)
                {
                    if ((fam_n + 1) * (2 + 4 * stride) > family_cap)
                    {
                        rc = -3;
                        break;
                    }
                    slot = fam_n++;
                    uint32_t* h = family_words + (size_t)slot * (2 + 4 * (size_t)stride);
                    memset(h, 0, sizeof(uint32_t) * (2 + 4 * (size_t)stride));
                    h[0] = (uint32_t)seqs;
                    h[1] = (uint32_t)(seqs >> 32);
                }
                pgo_count4* c = (pgo_count4*)(family_words + (size_t)slot * (2 + 4 * (size_t)stride) + 2);
                add4(&c[0], reads, fwd, rev);
                for (int v = 0; v < n_nodes; ++v)
                    if (nmark[v])
                        add4(&c[1 + v], reads, fwd, rev);
                for (int e = 0; e < n_edges; ++e)
                    if (emark[e])
                        add4(&c[1 + n_nodes + e], reads, fwd, rev);
            }
        }
        free(head);
        free(next);
        free(tail);
        free(nmark);
        free(emark);
    }
done:
    if (path_used)
        *path_used = pw;
    if (family_used)
        *family_used = fam_n * (2 + 4 * stride);
    free(path);
    free(g.out_labels);
    free(g.in_labels);
    return rc;
}
