/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of paragraph's exact-match stage, grm::PathAligner
 * (src/c++/lib/grm/PathAligner.cpp:75-164), with the graph-tools pieces it stands on:
 *   KmerIndex           graph-tools src/graphalign/KmerIndex.cpp:85-125     (every k-mer path of the graph)
 *   extendPath{Start,End}            src/graphcore/PathOperations.cpp:43-115
 *   extendPath{End,Start}Matching    src/graphcore/PathOperations.cpp:117-271
 *   projectAlignmentOntoGraph / generateCigar   src/graphalign/GraphAlignmentOperations.cpp:130-164,
 *                                               src/graphalign/GraphAlignment.cpp:106-117
 * Parity status: PINNED against oracle/_ref/libpgref.so (the unmodified PathAligner.cpp + KmerIndex.cpp compiled by
 * oracle/Makefile, driver oracle/ref_path.cpp) on seeded fuzz inputs and the committed fixtures
 * (tests/test_path_oracle.py, tests/golden/path_*.json).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pg_oracle.h"

typedef struct
{
    int n;
    char** seq; /* RAW node sequences: the k-mer index and the extension compare characters as stored
                   (graphtools::Graph keeps them as given, Graph.cpp:88-100) */
    int* len;
    int *nsucc, **succ, *npred, **pred; /* ascending ids (std::set iteration) */
} pgraph;

typedef struct
{
    int start, end; /* positions in the first / last node, both inclusive (Path.cpp:238-262) */
    int nn;
    int* nodes;
} ppath;

typedef struct kent
{
    char* kmer;
    int count;
    ppath first; /* kmer_to_paths_map[kmer].front() */
    struct kent* next;
} kent;

struct pgo_path_index
{
    pgraph g;
    int k;
    int nb;
    kent** bucket;
};

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

static void add_unique(int** arr, int* n, int v)
{
    for (int i = 0; i < *n; ++i)
        if ((*arr)[i] == v)
            return;
    *arr = (int*)realloc(*arr, sizeof(int) * (size_t)(*n + 1));
    (*arr)[(*n)++] = v;
}

static uint64_t hash_bytes(const char* s, int n)
{
    uint64_t h = 1469598103934665603ull;
    for (int i = 0; i < n; ++i)
    {
        h ^= (unsigned char)s[i];
        h *= 1099511628211ull;
    }
    return h;
}

static int path_len(const pgraph* g, const ppath* p) /* Path::length, Path.cpp:293-302 */
{
    if (p->nn == 1)
        return p->end - p->start + 1;
    int L = g->len[p->nodes[0]] - p->start + p->end + 1;
    for (int i = 1; i + 1 < p->nn; ++i)
        L += g->len[p->nodes[i]];
    return L;
}

/* insert one k-mer path (KmerIndex::Impl::addKmerPaths, KmerIndex.cpp:97-114; no sequence expansion: Graph{n,false}) */
static void index_add(struct pgo_path_index* ix, const int* nodes, int nn, int start, int end)
{
    const pgraph* g = &ix->g;
    char* km = (char*)malloc((size_t)ix->k + 1);
    int w = 0;
    for (int i = 0; i < nn; ++i)
    {
        const int a = i == 0 ? start : 0, b = i == nn - 1 ? end : g->len[nodes[i]] - 1;
        for (int x = a; x <= b; ++x)
            km[w++] = g->seq[nodes[i]][x];
    }
    km[w] = 0;
    kent** slot = &ix->bucket[hash_bytes(km, w) % (uint64_t)ix->nb];
    for (kent* e = *slot; e; e = e->next)
        if (memcmp(e->kmer, km, (size_t)w) == 0)
        {
            ++e->count;
            free(km);
            return;
        }
    kent* e = (kent*)calloc(1, sizeof(kent));
    e->kmer = km;
    e->count = 1;
    e->first.start = start;
    e->first.end = end;
    e->first.nn = nn;
    e->first.nodes = (int*)malloc(sizeof(int) * (size_t)nn);
    memcpy(e->first.nodes, nodes, sizeof(int) * (size_t)nn);
    e->next = *slot;
    *slot = e;
}

/* extendPathEnd (PathOperations.cpp:73-103) from a path that ends at (nodes[nn-1], end): depth-first over successors */
static void extend_end(struct pgo_path_index* ix, int* nodes, int nn, int start, int end, int ext)
{
    const pgraph* g = &ix->g;
    const int last = nodes[nn - 1];
    const int room = g->len[last] - end - 1;
    if (ext <= room)
    {
        index_add(ix, nodes, nn, start, end + ext);
        return;
    }
    for (int s = 0; s < g->nsucc[last]; ++s)
    {
        nodes[nn] = g->succ[last][s];
        extend_end(ix, nodes, nn + 1, start, 0, ext - room - 1);
    }
}

struct pgo_path_index* pgo_path_index_create(int n_nodes, const char* blob, const int32_t* off, int n_edges,
                                             const int32_t* ef, const int32_t* et, int kmer_len)
{
    if (n_nodes <= 0 || kmer_len <= 0)
        return NULL;
    struct pgo_path_index* ix = (struct pgo_path_index*)calloc(1, sizeof(*ix));
    pgraph* g = &ix->g;
    g->n = n_nodes;
    g->seq = (char**)calloc((size_t)n_nodes, sizeof(char*));
    g->len = (int*)calloc((size_t)n_nodes, sizeof(int));
    g->nsucc = (int*)calloc((size_t)n_nodes, sizeof(int));
    g->npred = (int*)calloc((size_t)n_nodes, sizeof(int));
    g->succ = (int**)calloc((size_t)n_nodes, sizeof(int*));
    g->pred = (int**)calloc((size_t)n_nodes, sizeof(int*));
    long total = 0;
    for (int i = 0; i < n_nodes; ++i)
    {
        g->len[i] = off[i + 1] - off[i];
        g->seq[i] = (char*)malloc((size_t)g->len[i] + 1);
        memcpy(g->seq[i], blob + off[i], (size_t)g->len[i]);
        g->seq[i][g->len[i]] = 0;
        total += g->len[i];
    }
    for (int e = 0; e < n_edges; ++e)
    {
        add_unique(&g->succ[ef[e]], &g->nsucc[ef[e]], et[e]);
        add_unique(&g->pred[et[e]], &g->npred[et[e]], ef[e]);
    }
    for (int i = 0; i < n_nodes; ++i)
    {
        qsort(g->succ[i], (size_t)g->nsucc[i], sizeof(int), cmp_int);
        qsort(g->pred[i], (size_t)g->npred[i], sizeof(int), cmp_int);
    }
    ix->k = kmer_len;
    ix->nb = (int)(4 * total + 64);
    ix->bucket = (kent**)calloc((size_t)ix->nb, sizeof(kent*));
    int* nodes = (int*)malloc(sizeof(int) * (size_t)(kmer_len + 2));
    /* KmerIndex::Impl::Impl / addKmerPathsStartingAtNode, KmerIndex.cpp:75-95 */
    for (int v = 0; v < n_nodes; ++v)
        for (int pos = 0; pos < g->len[v]; ++pos)
        {
            nodes[0] = v;
            extend_end(ix, nodes, 1, pos, pos, kmer_len - 1);
        }
    free(nodes);
    return ix;
}

void pgo_path_index_destroy(struct pgo_path_index* ix)
{
    if (!ix)
        return;
    for (int b = 0; b < ix->nb; ++b)
        for (kent* e = ix->bucket[b]; e;)
        {
            kent* nx = e->next;
            free(e->kmer);
            free(e->first.nodes);
            free(e);
            e = nx;
        }
    free(ix->bucket);
    pgraph* g = &ix->g;
    for (int i = 0; i < g->n; ++i)
    {
        free(g->seq[i]);
        free(g->succ[i]);
        free(g->pred[i]);
    }
    free(g->seq);
    free(g->len);
    free(g->nsucc);
    free(g->npred);
    free(g->succ);
    free(g->pred);
    free(ix);
}

static const kent* index_find(const struct pgo_path_index* ix, const char* kmer)
{
    for (const kent* e = ix->bucket[hash_bytes(kmer, ix->k) % (uint64_t)ix->nb]; e; e = e->next)
        if (memcmp(e->kmer, kmer, (size_t)ix->k) == 0)
            return e;
    return NULL;
}

typedef struct
{
    int* v;
    int n, cap, head; /* nodes live in v[head .. head+n) so that the start can grow to the left */
} nodelist;

static void nl_init(nodelist* l, const int* nodes, int nn, int room)
{
    l->cap = nn + 2 * room + 4;
    l->v = (int*)malloc(sizeof(int) * (size_t)l->cap);
    l->head = room + 2;
    l->n = nn;
    memcpy(l->v + l->head, nodes, sizeof(int) * (size_t)nn);
}

/* extendPathEndMatching, PathOperations.cpp:117-189.  (*end) in/out = end position in the last node */
static void extend_end_matching(const pgraph* g, nodelist* nl, int* end, int plen, const char* q, int qlen, int qpos)
{
    int pos_in_query = qpos + plen;
    int node = nl->v[nl->head + nl->n - 1];
    int pos_in_node = *end + 1;
    int moved = 1;
    while (moved)
    {
        moved = 0;
        while (pos_in_query < qlen && pos_in_node < g->len[node] && q[pos_in_query] == g->seq[node][pos_in_node])
        {
            moved = 1;
            ++pos_in_node;
            ++pos_in_query;
        }
        if (pos_in_node >= g->len[node])
        {
            int num_longest = 0, longest = 0, best = 0;
            int min_size = 0x7fffffff;
            for (int s = 0; s < g->nsucc[node]; ++s)
                if (g->len[g->succ[node][s]] < min_size)
                    min_size = g->len[g->succ[node][s]];
            for (int s = 0; s < g->nsucc[node]; ++s)
            {
                const int c = g->succ[node][s];
                int p = 0;
                while (p < min_size && pos_in_query + p < qlen && g->seq[c][p] == q[pos_in_query + p])
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
            nl->v[nl->head + nl->n++] = best;
            pos_in_query += longest;
            pos_in_node = longest;
            node = best;
            moved = 1;
        }
    }
    *end = pos_in_node - 1;
}

/* extendPathStartMatching, PathOperations.cpp:191-266.  (*start) in/out; *qpos in/out */
static void extend_start_matching(const pgraph* g, nodelist* nl, int* start, const char* q, int* qpos)
{
    int pos_in_query = *qpos;
    int node = nl->v[nl->head];
    int pos_in_node = *start;
    int moved = 1;
    while (moved)
    {
        moved = 0;
        while (pos_in_query > 0 && pos_in_node > 0 && q[pos_in_query - 1] == g->seq[node][pos_in_node - 1])
        {
            moved = 1;
            --pos_in_node;
            --pos_in_query;
        }
        if (pos_in_node == 0)
        {
            int num_longest = 0, longest = 0, best = 0;
            int min_size = 0x7fffffff;
            for (int s = 0; s < g->npred[node]; ++s)
                if (g->len[g->pred[node][s]] < min_size)
                    min_size = g->len[g->pred[node][s]];
            for (int s = 0; s < g->npred[node]; ++s)
            {
                const int c = g->pred[node][s];
                int pp = g->len[c], m = 0;
                while (pp > g->len[c] - min_size && pos_in_query - m > 0 && g->seq[c][pp - 1] == q[pos_in_query - m - 1])
                {
                    --pp;
                    ++m;
                }
                if (m > longest)
                {
                    longest = m;
                    best = c;
                    num_longest = 1;
                }
                else if (m == longest)
                    ++num_longest;
            }
            if (longest == 0 || num_longest != 1)
                break;
            nl->v[--nl->head] = best;
            ++nl->n;
            pos_in_query -= longest;
            node = best;
            pos_in_node = g->len[node] - longest;
            moved = 1;
        }
    }
    *start = pos_in_node;
    *qpos = pos_in_query;
}

static char comp_base(char b) /* graph-tools SequenceOperations.cpp:66-81 */
{
    switch (b)
    {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
    }
}

/* PathAligner::alignRead, PathAligner.cpp:75-164.
 * out8 = {mapped, graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, anchored} */
int pgo_path_align_read(const struct pgo_path_index* ix, const char* bases, int L, int32_t* out8, char* out_bases,
                        char* cigar, int cigar_cap)
{
    const pgraph* g = &ix->g;
    const int k = ix->k;
    memset(out8, 0, 8 * sizeof(int32_t));
    if (cigar && cigar_cap > 0)
        cigar[0] = 0;
    if (out_bases)
        memcpy(out_bases, bases, (size_t)L);
    if (L < k) /* :83-87 */
        return 0;
    char* rc = (char*)malloc((size_t)L + 1);
    for (int i = 0; i < L; ++i)
        rc[i] = comp_base(bases[L - 1 - i]);
    rc[L] = 0;
    int n_matches = 0, n_full = 0;
    int full_rev = 0, full_start = 0, full_end = 0;
    nodelist full;
    memset(&full, 0, sizeof full);
    for (int strand = 0; strand < 2; ++strand) /* :90-109 */
    {
        const char* q = strand ? rc : bases;
        for (int pos = 0; pos + k <= L; ++pos)
        {
            const kent* e = index_find(ix, q + pos);
            if (!e || e->count != 1)
                continue;
            int qpos = pos;
            nodelist nl;
            nl_init(&nl, e->first.nodes, e->first.nn, L + 2);
            int start = e->first.start, end = e->first.end;
            extend_end_matching(g, &nl, &end, path_len(g, &e->first), q, L, qpos);
            extend_start_matching(g, &nl, &start, q, &qpos);
            ppath ext = { start, end, nl.n, nl.v + nl.head };
            const int plen = path_len(g, &ext);
            ++n_matches;
            if (plen == L) /* the filter of :117-118 */
            {
                if (n_full == 0)
                {
                    full = nl;
                    full_rev = strand;
                    full_start = start;
                    full_end = end;
                    nl.v = NULL;
                }
                ++n_full;
            }
            free(nl.v);
            pos = qpos + plen; /* :106, then ++pos */
        }
    }
    out8[7] = n_matches > 0; /* anchored_, :111-114 */
    if (n_full > 0)
    {
        out8[0] = 1;
        out8[1] = full_start;          /* path.startPosition(), :149 */
        out8[2] = L;                   /* path.length(), :147 */
        out8[3] = n_full == 1;         /* :152-161 */
        out8[4] = n_full == 1 ? 60 : 0;
        out8[5] = full_rev;            /* set_is_graph_reverse_strand(isReverse), :127-135 -- not xor-ed with the read's strand */
        if (full_rev && out_bases)
            memcpy(out_bases, rc, (size_t)L);
        /* "<L>M" projected onto the path: one "<overlap>M" per node */
        int len = 0;
        char tmp[48];
        for (int i = 0; i < full.n; ++i)
        {
            const int node = full.v[full.head + i];
            int ov = g->len[node];
            if (full.n == 1)
                ov = full_end - full_start + 1;
            else if (i == 0)
                ov = g->len[node] - full_start;
            else if (i == full.n - 1)
                ov = full_end + 1;
            snprintf(tmp, sizeof tmp, "%d[%dM]", node, ov);
            for (const char* p = tmp; *p; ++p)
            {
                if (cigar && len < cigar_cap - 1)
                    cigar[len] = *p;
                ++len;
            }
        }
        if (cigar && cigar_cap > 0)
            cigar[len < cigar_cap - 1 ? len : cigar_cap - 1] = 0;
        out8[6] = len;
        free(full.v);
    }
    free(rc);
    return 0;
}

int pgo_path_align_batch(const struct pgo_path_index* ix, int n_reads, const char* blob, const int32_t* off,
                         int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride, int32_t* counters3)
{
    counters3[0] = counters3[1] = counters3[2] = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        pgo_path_align_read(ix, blob + off[i], off[i + 1] - off[i], out8 + 8 * i,
                            out_bases_blob ? out_bases_blob + off[i] : NULL,
                            cigars ? cigars + (size_t)i * cigar_stride : NULL, cigar_stride);
        counters3[0] += 1;
        counters3[1] += out8[8 * i + 7];
        counters3[2] += out8[8 * i + 0];
    }
    return 0;
}
