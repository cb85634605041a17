/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of grm::KmerAligner<K>, the second stage of the reference's
 * CompositeAligner cascade (src/c++/lib/grm/CompositeAligner.cpp:105-126; off by default in both CLIs,
 * src/c++/main/grmpy.cpp:72).  Reference: src/c++/lib/grm/KmerAligner.cpp.
 *
 * What it computes: gapless alignment of a read (both strands) to the sequence of one of the graph's PATHS (the
 * "paths" of the graph JSON, GraphInput.cpp:168-197): every k-mer the read shares with a path proposes an offset
 * (KmerAligner.cpp:246-294), the offsets are scored by their number of mismatching characters, the n_paths + 1 best
 * are kept in a bounded heap, and pickBest (:479-517) maps the read to the best one if it has at most two mismatches --
 * uniquely unless an equally good candidate that FOLLOWS it in the heap array gives a different (position, CIGAR).
 *
 * The result depends on the C++ library: the candidates live in a std::vector managed with std::push_heap /
 * std::pop_heap under a comparator that only looks at the mismatch count, and pickBest scans that vector in ARRAY
 * order.  Which of several equally bad candidates is evicted, and which equally good one comes first, is decided by
 * the heap algorithm.  The reference is built with GNU libstdc++ (here: 13.3); its <bits/stl_heap.h> algorithms
 * (__push_heap: sift the value up while the parent compares less; __adjust_heap: move the larger child up to a leaf,
 * then push the value up from there; pop_heap: swap first and last, __adjust_heap from the root) are restated below
 * operation by operation.
 *
 * Parity status: PINNED (tests/test_kmer_oracle.py) against the reference's own unit-test vectors
 * (src/c++/test/test_kmeraligner.cpp:149-191, K = 10) and against oracle/_ref/libpgref.so = the unmodified
 * KmerAligner.cpp compiled here (oracle/ref_kmer.cpp; only oligo/KmerGenerator.hh is a stand-in) on seeded fuzz inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "pg_oracle.h"

typedef struct
{
    uint32_t kmer;
    int32_t pos;
} kpos;

typedef struct
{
    int path, pos, rev;
    unsigned mm;
} cand;

typedef struct
{
    int n_nodes;     /* nodes of the path */
    int* node_id;    /* [n_nodes] */
    int* node_start; /* [n_nodes] offset of each node in the path sequence (BasicPath::starts, :140-160) */
    char* seq;       /* path sequence = node sequences as given, concatenated (Path::seq) */
    int len;
    kpos* kmers; /* sorted by (kmer, position) (makeKmers, :120-133) */
    int n_kmers;
} kpath;

struct pgo_kmer_index
{
    int k;
    int n_paths;
    kpath* paths;
};

/* oligo::Translator<> (Nucleotides.hh:59-341): A/a 0, C/c 1, G/g 2, T/t 3, everything else invalid */
static int base_value(char c)
{
    switch (c)
    {
    case 'A': case 'a': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

static int kpos_less(const void* a, const void* b)
{
    const kpos* x = (const kpos*)a;
    const kpos* y = (const kpos*)b;
    if (x->kmer != y->kmer)
        return x->kmer < y->kmer ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos ? 1 : 0);
}

/* makeKmers (KmerAligner.cpp:120-133) over oligo::KmerGenerator (KmerGenerator.hh:41-157): every run of k valid
 * characters gives a k-mer (2 bits per base, first base most significant) at the position of its first base; k-mers
 * containing any other character are skipped; then sorted by (k-mer, position).  (An empty sequence gives none.) */
static int make_kmers(const char* s, int len, int k, kpos* out)
{
    int n = 0, have = 0;
    uint32_t v = 0;
    const uint32_t mask = k >= 16 ? 0xFFFFFFFFu : ((1u << (2 * k)) - 1u);
    for (int i = 0; i < len; ++i)
    {
        const int b = base_value(s[i]);
        if (b > 3)
        {
            have = 0;
            continue;
        }
        v = (v << 2) | (uint32_t)b;
        if (++have >= k)
        {
            out[n].kmer = v & mask;
            out[n].pos = i - k + 1;
            ++n;
        }
    }
    qsort(out, (size_t)n, sizeof(kpos), kpos_less);
    return n;
}

struct pgo_kmer_index* pgo_kmer_index_create(int n_nodes, const char* blob, const int32_t* off, int n_paths,
                                             const int32_t* path_ptr, const int32_t* path_nodes, int k)
{
    if (n_nodes <= 0 || !blob || !off || n_paths < 0 || k < 2 || k > 16)
        return NULL;
    struct pgo_kmer_index* ix = (struct pgo_kmer_index*)calloc(1, sizeof(*ix));
    ix->k = k;
    ix->n_paths = n_paths;
    ix->paths = (kpath*)calloc((size_t)(n_paths > 0 ? n_paths : 1), sizeof(kpath));
    for (int p = 0; p < n_paths; ++p)
    {
        kpath* kp = &ix->paths[p];
        kp->n_nodes = path_ptr[p + 1] - path_ptr[p];
        kp->node_id = (int*)malloc(sizeof(int) * (size_t)(kp->n_nodes > 0 ? kp->n_nodes : 1));
        kp->node_start = (int*)malloc(sizeof(int) * (size_t)(kp->n_nodes > 0 ? kp->n_nodes : 1));
        int len = 0;
        for (int i = 0; i < kp->n_nodes; ++i)
        {
            const int v = path_nodes[path_ptr[p] + i];
            kp->node_id[i] = v;
            kp->node_start[i] = len;
            len += off[v + 1] - off[v];
        }
        kp->len = len;
        kp->seq = (char*)malloc((size_t)len + 1);
        for (int i = 0; i < kp->n_nodes; ++i)
        {
            const int v = kp->node_id[i];
            memcpy(kp->seq + kp->node_start[i], blob + off[v], (size_t)(off[v + 1] - off[v]));
        }
        kp->seq[len] = 0;
        kp->kmers = (kpos*)malloc(sizeof(kpos) * (size_t)(len > 0 ? len : 1));
        kp->n_kmers = make_kmers(kp->seq, len, k, kp->kmers);
    }
    return ix;
}

void pgo_kmer_index_destroy(struct pgo_kmer_index* ix)
{
    if (!ix)
        return;
    for (int p = 0; p < ix->n_paths; ++p)
    {
        free(ix->paths[p].node_id);
        free(ix->paths[p].node_start);
        free(ix->paths[p].seq);
        free(ix->paths[p].kmers);
    }
    free(ix->paths);
    free(ix);
}

/* ---- GNU libstdc++ <bits/stl_heap.h>, comparator = Candidate::lessMismatches (KmerAligner.cpp:78-83) ---- */
static int less_mm(const cand* a, const cand* b) { return a->mm < b->mm; }

static void push_heap_(cand* first, long hole, long top, cand value) /* std::__push_heap */
{
    long parent = (hole - 1) / 2;
    while (hole > top && less_mm(&first[parent], &value))
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}

static void adjust_heap_(cand* first, long hole, long len, cand value) /* std::__adjust_heap */
{
    const long top = hole;
    long child = hole;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (less_mm(&first[child], &first[child - 1]))
            --child;
        first[hole] = first[child];
        hole = child;
    }
    if ((len & 1) == 0 && child == (len - 2) / 2)
    {
        child = 2 * (child + 1);
        first[hole] = first[child - 1];
        hole = child - 1;
    }
    push_heap_(first, hole, top, value);
}

/* std::push_heap(first, first + n): the new element is first[n - 1] */
static void std_push_heap(cand* first, long n) { push_heap_(first, n - 1, 0, first[n - 1]); }
/* std::pop_heap(first, first + n): afterwards the largest element is first[n - 1] */
static void std_pop_heap(cand* first, long n)
{
    if (n > 1)
    {
        const cand value = first[n - 1];
        first[n - 1] = first[0];
        adjust_heap_(first, 0, n - 1, value);
    }
}

static int int_less(const void* a, const void* b)
{
    const int x = *(const int*)a, y = *(const int*)b;
    return x < y ? -1 : (x > y ? 1 : 0);
}

/* align<reverse> (KmerAligner.cpp:246-294) */
static void seed_path(const struct pgo_kmer_index* ix, int p, int rev, const char* seq, int L, const kpos* sk, int nsk,
                      cand* cands, int* n_cands, int capacity, int* seeds)
{
    const kpath* kp = &ix->paths[p];
    int ns = 0, pp = 0;
    for (int i = 0; i < nsk; ++i)
    {
        /* one forward walk over the path's sorted k-mers: a k-mer that occurs twice in the READ finds the path's
         * occurrences already consumed by its first occurrence (:253-272) */
        while (pp < kp->n_kmers && kp->kmers[pp].kmer < sk[i].kmer)
            ++pp;
        while (pp < kp->n_kmers && kp->kmers[pp].kmer == sk[i].kmer)
        {
            const int offset = kp->kmers[pp].pos - sk[i].pos;
            if (0 <= offset && kp->len >= offset + L) /* candidates that overhang the path are ignored */
                seeds[ns++] = offset;
            ++pp;
        }
    }
    qsort(seeds, (size_t)ns, sizeof(int), int_less); /* :274-281: sorted by position, duplicates removed */
    int nu = 0;
    for (int i = 0; i < ns; ++i)
        if (nu == 0 || seeds[nu - 1] != seeds[i])
            seeds[nu++] = seeds[i];
    for (int i = 0; i < nu; ++i)
    {
        unsigned mm = 0; /* countMismatches (:232-240): plain character comparison */
        for (int x = 0; x < L; ++x)
            mm += seq[x] != kp->seq[seeds[i] + x];
        cand c;
        c.path = p;
        c.pos = seeds[i];
        c.rev = rev;
        c.mm = mm;
        cands[(*n_cands)++] = c;
        std_push_heap(cands, *n_cands);
        if (capacity == *n_cands) /* full: the candidate with most mismatches goes (:286-292) */
        {
            std_pop_heap(cands, *n_cands);
            --*n_cands;
        }
    }
}

/* graphtools::reverseComplement (SequenceOperations.cpp:66-89): case-sensitive, anything but ACGT -> 'N' */
static char comp_base(char c) { return c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N'; }

static int put_num(char* out, int cap, int at, long v)
{
    char tmp[24];
    int n = 0;
    if (v == 0)
        tmp[n++] = '0';
    while (v > 0)
    {
        tmp[n++] = (char)('0' + v % 10);
        v /= 10;
    }
    while (n > 0)
    {
        if (at < cap - 1)
            out[at] = tmp[n - 1];
        ++at;
        --n;
    }
    return at;
}
static int put_chr(char* out, int cap, int at, char c)
{
    if (at < cap - 1)
        out[at] = c;
    return at + 1;
}

/* updateAlignment + buildCigar + makeCigarBit (KmerAligner.cpp:318-478).  Returns the CIGAR length; *pos_out =
 * graph_pos, *score_out = number of matching bases. */
static int update_alignment(const struct pgo_kmer_index* ix, const cand* c, const char* seq, int L, int* pos_out,
                            int* score_out, char* cigar, int cap)
{
    const kpath* kp = &ix->paths[c->path];
    const char* ref = kp->seq + c->pos;
    int left = 0; /* calculateSoftClip (:318-325): leading / trailing 'N' of the PATH (N-filled source / sink) */
    while (left < L && ref[left] == 'N')
        ++left;
    int right = 0;
    while (right < L - left && ref[L - 1 - right] == 'N')
        ++right;
    const int pos = c->pos + left;
    /* findStartNode (:162-178): the path node that contains offset pos (the last one when pos is past the end) */
    int sn = kp->n_nodes - 1;
    for (int i = 0; i < kp->n_nodes; ++i)
        if (kp->node_start[i] >= pos)
        {
            sn = kp->node_start[i] > pos ? i - 1 : i;
            break;
        }
    if (sn < 0)
        sn = 0;
    int at = 0, score = 0;
    long this_start = pos - kp->node_start[sn];
    *pos_out = (int)this_start;
    long left_len = L - left - right;
    int lclip = left;
    const char* q = seq + left;
    for (int i = sn; i < kp->n_nodes && left_len > 0; ++i) /* buildCigar (:374-421) */
    {
        long this_len = left_len;
        if (i + 1 < kp->n_nodes)
        {
            const long room = (long)kp->node_start[i + 1] - kp->node_start[i] - this_start;
            if (room < this_len)
                this_len = room;
        }
        if (this_len > 0)
        {
            const char* r = kp->seq + this_start + kp->node_start[i];
            at = put_num(cigar, cap, at, kp->node_id[i]);
            at = put_chr(cigar, cap, at, '[');
            if (lclip)
            {
                at = put_num(cigar, cap, at, lclip);
                at = put_chr(cigar, cap, at, 'S');
                lclip = 0;
            }
            /* makeCigarBit (:334-372); getCigarOp(ref, read) (:316): equal -> M, else N if either is 'N', else X */
            char last = 0;
            long run = 0;
            for (long x = 0; x < this_len; ++x)
            {
                const char s = r[x], b = q[x];
                const char op = s == b ? 'M' : (s == 'N' || b == 'N') ? 'N' : 'X';
                if (op != last)
                {
                    if (run)
                    {
                        at = put_num(cigar, cap, at, run);
                        at = put_chr(cigar, cap, at, last);
                        if (last == 'M')
                            score += (int)run;
                    }
                    last = op;
                    run = 0;
                }
                ++run;
            }
            if (run)
            {
                at = put_num(cigar, cap, at, run);
                at = put_chr(cigar, cap, at, last);
                if (last == 'M')
                    score += (int)run;
            }
            q += this_len;
            if (right && this_len == left_len)
            {
                at = put_num(cigar, cap, at, right);
                at = put_chr(cigar, cap, at, 'S');
            }
            at = put_chr(cigar, cap, at, ']');
        }
        left_len -= this_len;
        this_start = 0;
    }
    if (cap > 0)
        cigar[at < cap - 1 ? at : cap - 1] = 0;
    *score_out = score;
    return at;
}

/* first element with the fewest mismatches in [from, n) -- std::min_element; n when the range is empty */
static int min_element_(const cand* c, int from, int n)
{
    if (from >= n)
        return n;
    int best = from;
    for (int i = from + 1; i < n; ++i)
        if (less_mm(&c[i], &c[best]))
            best = i;
    return best;
}

/* KmerAlignerImpl::alignRead (KmerAligner.cpp:519-536).
 * out8 = {status (0 UNMAPPED, 1 MAPPED, 2 BAD_ALIGN), graph_pos, score, unique, mapq, is_graph_reverse_strand,
 *         cigar_strlen, 0}; out_bases receives the read's bases after the call (reverse complement when the best
 * candidate is on the reverse strand). */
int pgo_kmer_align_read(const struct pgo_kmer_index* ix, const char* bases, int L, int is_reverse_strand, int32_t* out8,
                        char* out_bases, char* cigar, int cigar_cap)
{
    if (!ix || !bases || L < 0 || !out8)
        return PGO_E_ARG;
    memset(out8, 0, 8 * sizeof(int32_t));
    out8[5] = 0;
    if (out_bases)
        memcpy(out_bases, bases, (size_t)L);
    if (cigar && cigar_cap > 0)
        cigar[0] = 0;
    const int capacity = ix->n_paths + 2; /* :306-309 */
    cand* cands = (cand*)malloc(sizeof(cand) * (size_t)capacity);
    char* rv = (char*)malloc((size_t)L + 1);
    kpos* fk = (kpos*)malloc(sizeof(kpos) * (size_t)(L + 1));
    kpos* rk = (kpos*)malloc(sizeof(kpos) * (size_t)(L + 1));
    for (int i = 0; i < L; ++i)
        rv[i] = comp_base(bases[L - 1 - i]);
    rv[L] = 0;
    const int nf = L > 0 ? make_kmers(bases, L, ix->k, fk) : 0;
    const int nr = L > 0 ? make_kmers(rv, L, ix->k, rk) : 0;
    int max_path = 1;
    for (int p = 0; p < ix->n_paths; ++p)
        if (ix->paths[p].len > max_path)
            max_path = ix->paths[p].len;
    /* a read k-mer meets at most every path position once */
    int* seeds = (int*)malloc(sizeof(int) * ((size_t)max_path + 1));
    int n = 0;
    for (int p = 0; p < ix->n_paths; ++p)
    {
        seed_path(ix, p, 0, bases, L, fk, nf, cands, &n, capacity, seeds);
        seed_path(ix, p, 1, rv, L, rk, nr, cands, &n, capacity, seeds);
    }
    if (n > 0) /* pickBest (:479-517) */
    {
        const int b = min_element_(cands, 0, n);
        if (cands[b].mm <= 2)
        {
            int pos = 0, score = 0;
            const int clen = update_alignment(ix, &cands[b], cands[b].rev ? rv : bases, L, &pos, &score, cigar, cigar_cap);
            int unique = 1, status = 1;
            char* c2 = (char*)malloc((size_t)(clen > 0 ? clen : 0) + 64 + 16 * (size_t)(L + 4));
            const int c2cap = (clen > 0 ? clen : 0) + 64 + 16 * (L + 4);
            for (int s = min_element_(cands, b + 1, n); s < n; s = min_element_(cands, s + 1, n))
            {
                if (cands[s].mm != cands[b].mm)
                    break; /* no more as good candidates */
                int pos2 = 0, score2 = 0;
                const int l2 = update_alignment(ix, &cands[s], cands[s].rev ? rv : bases, L, &pos2, &score2, c2, c2cap);
                int same = pos2 == pos && l2 == clen;
                if (same && cigar && clen < cigar_cap)
                    same = memcmp(c2, cigar, (size_t)clen) == 0;
                if (!same)
                {
                    unique = 0;
                    status = 2;
                    break;
                }
            }
            free(c2);
            out8[0] = status;
            out8[1] = pos;
            out8[2] = score;
            out8[3] = unique;
            out8[4] = unique ? 60 : 0;
            out8[5] = cands[b].rev ? !is_reverse_strand : (is_reverse_strand != 0);
            out8[6] = clen;
            if (cands[b].rev && out_bases)
                memcpy(out_bases, rv, (size_t)L);
        }
    }
    free(seeds);
    free(rk);
    free(fk);
    free(rv);
    free(cands);
    return PGO_OK;
}

int pgo_kmer_align_batch(const struct pgo_kmer_index* ix, int n_reads, const char* blob, const int32_t* off,
                         const uint8_t* is_rev, int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride)
{
    for (int i = 0; i < n_reads; ++i)
    {
        const int rc = pgo_kmer_align_read(ix, blob + off[i], off[i + 1] - off[i], is_rev ? is_rev[i] : 0, out8 + 8 * i,
                                           out_bases_blob ? out_bases_blob + off[i] : NULL,
                                           cigars ? cigars + (size_t)i * cigar_stride : NULL, cigar_stride);
        if (rc != PGO_OK)
            return rc;
    }
    return PGO_OK;
}
