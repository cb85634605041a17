/* TEST INFRASTRUCTURE ONLY -- see pg_oracle.h.
 *
 * Plain-C restatement of the reference's read->graph alignment path:
 *   external/gssw/gssw.c            (striped 8-bit Smith-Waterman over graph nodes, traceback)
 *   src/c++/lib/grm/GraphAligner.cpp (4 fills per read, uniqueness, strand choice, CIGAR string)
 * The SSE2 code is restated lane by lane (16 unsigned-saturating byte lanes, Farrar striping,
 * the lazy-F loop with its global termination test) so that mH/mE/mF come out byte-identical,
 * including the quirks (E derived from the pre-lazy-F H, stored F only where the lazy loop
 * went, padded lanes).  Each function cites the reference lines it follows.
 */
#include "pg_oracle.h"

#include <ctype.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define NL 16 /* byte lanes of an __m128i */

/* ------------------------------------------------------------------ scoring tables */

/* gssw_create_nt_table, gssw.c:4206-4220 (A0 C1 G2 T3, 'U'->0, everything else 4) */
static int8_t nt_code(unsigned char c)
{
    switch (c)
    {
    case 'A': case 'a': case 'U': case 'u': return 0;
    case 'C': case 'c': return 1;
    case 'G': case 'g': return 2;
    case 'T': case 't': return 3;
    default: return 4;
    }
}

/* gssw_create_score_matrix(match=1, mismatch=4), gssw.c:4188-4204; GraphAligner.cpp:229-233 */
static int8_t sub_score(int a, int b)
{
    if (a == 4 || b == 4)
        return 0;
    return a == b ? 1 : -4;
}

enum { GAP_OPEN = 6, GAP_EXT = 1, BIAS = 4 /* |min(mat)|, gssw.c:818-823 */ };

static inline uint8_t adds8(uint8_t a, uint8_t b) { unsigned s = (unsigned)a + b; return (uint8_t)(s > 255 ? 255 : s); }
static inline uint8_t subs8(uint8_t a, uint8_t b) { return (uint8_t)(a > b ? a - b : 0); }
static inline uint8_t max8(uint8_t a, uint8_t b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ graph */

typedef struct
{
    int n;
    char** seq; /* upper-cased, NUL terminated */
    int8_t** num;
    int* len;
    int* npred;
    int** pred; /* ascending ids */
} g1;

struct pgo_graph
{
    g1 fwd, rev;
};

static int cmp_int(const void* a, const void* b) { return *(const int*)a - *(const int*)b; }

static int g1_build(g1* g, int n, char** seqs, const int* lens, int ne, const int32_t* ef, const int32_t* et)
{
    g->n = n;
    g->seq = (char**)calloc((size_t)n, sizeof(char*));
    g->num = (int8_t**)calloc((size_t)n, sizeof(int8_t*));
    g->len = (int*)calloc((size_t)n, sizeof(int));
    g->npred = (int*)calloc((size_t)n, sizeof(int));
    g->pred = (int**)calloc((size_t)n, sizeof(int*));
    for (int i = 0; i < n; ++i)
    {
        g->len[i] = lens[i];
        g->seq[i] = (char*)malloc((size_t)lens[i] + 1);
        g->num[i] = (int8_t*)malloc((size_t)lens[i] + 1);
        for (int k = 0; k < lens[i]; ++k)
        {
            g->seq[i][k] = (char)toupper((unsigned char)seqs[i][k]); /* GraphAligner.cpp:126,138 */
            g->num[i][k] = nt_code((unsigned char)g->seq[i][k]);     /* gssw_node_create -> gssw_create_num */
        }
        g->seq[i][lens[i]] = 0;
        g->pred[i] = (int*)malloc(sizeof(int) * (size_t)(ne > 0 ? ne : 1));
    }
    for (int e = 0; e < ne; ++e)
    {
        if (ef[e] < 0 || et[e] >= n || ef[e] >= et[e])
            return PGO_E_ARG;
        g->pred[et[e]][g->npred[et[e]]++] = ef[e];
    }
    for (int i = 0; i < n; ++i)
        qsort(g->pred[i], (size_t)g->npred[i], sizeof(int), cmp_int); /* std::set order, GraphAligner.cpp:147 */
    return PGO_OK;
}

static void g1_free(g1* g)
{
    for (int i = 0; i < g->n; ++i)
    {
        free(g->seq[i]);
        free(g->num[i]);
        free(g->pred[i]);
    }
    free(g->seq);
    free(g->num);
    free(g->len);
    free(g->npred);
    free(g->pred);
}

pgo_graph* pgo_graph_create(int n, const char* blob, const int32_t* off, int ne, const int32_t* ef, const int32_t* et)
{
    if (n <= 0)
        return NULL;
    pgo_graph* G = (pgo_graph*)calloc(1, sizeof(pgo_graph));
    char** seqs = (char**)malloc(sizeof(char*) * (size_t)n);
    char** rseqs = (char**)malloc(sizeof(char*) * (size_t)n);
    int* lens = (int*)malloc(sizeof(int) * (size_t)n);
    int* rlens = (int*)malloc(sizeof(int) * (size_t)n);
    int32_t* ref = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne > 0 ? ne : 1));
    int32_t* ret = (int32_t*)malloc(sizeof(int32_t) * (size_t)(ne > 0 ? ne : 1));
    for (int i = 0; i < n; ++i)
    {
        lens[i] = off[i + 1] - off[i];
        seqs[i] = (char*)blob + off[i];
    }
    /* reverseGraph(graph, complement=false): node i -> n-1-i, sequence reversed, edges flipped
     * (graph-tools src/graphcore/GraphOperations.cpp:38-60) */
    for (int i = 0; i < n; ++i)
    {
        int r = n - 1 - i;
        rlens[r] = lens[i];
        rseqs[r] = (char*)malloc((size_t)lens[i] + 1);
        for (int k = 0; k < lens[i]; ++k)
            rseqs[r][k] = seqs[i][lens[i] - 1 - k];
    }
    for (int e = 0; e < ne; ++e)
    {
        ref[e] = n - 1 - et[e];
        ret[e] = n - 1 - ef[e];
    }
    int rc = g1_build(&G->fwd, n, seqs, lens, ne, ef, et);
    int rc2 = g1_build(&G->rev, n, rseqs, rlens, ne, ref, ret);
    for (int i = 0; i < n; ++i)
        free(rseqs[i]);
    free(seqs);
    free(rseqs);
    free(lens);
    free(rlens);
    free(ref);
    free(ret);
    if (rc != PGO_OK || rc2 != PGO_OK)
    {
        pgo_graph_destroy(G);
        return NULL;
    }
    return G;
}

void pgo_graph_destroy(pgo_graph* G)
{
    if (!G)
        return;
    g1_free(&G->fwd);
    g1_free(&G->rev);
    free(G);
}

/* ------------------------------------------------------------------ fill */

typedef struct
{
    /* len x L, row = reference position (gssw.c:395-430 byte mode, :706-740 word mode).  Held as 16-bit cells in
     * both modes; in byte mode every value is <= 255 (the reference's arrays are uint8_t there). */
    uint16_t *mH, *mE, *mF;
    uint8_t *seedH, *seedE;      /* byte mode: striped [segLen][16]  (alignment->seed, gssw.c:442-443) */
    uint16_t *seedH16, *seedE16; /* word mode: striped [segLen8][8]  (gssw.c:745-746) */
    int score1, ref_end1, read_end1;
    int is_byte; /* alignment->is_byte (gssw.c:226, 598) */
} naln;

/* gssw_qP_byte, gssw.c:72-98: prof[nt][i][lane] = j>=L ? bias : mat[nt][read[j]]+bias, j = i + lane*segLen */
static uint8_t* make_profile(const char* read, int L, int segLen)
{
    uint8_t* p = (uint8_t*)malloc((size_t)5 * segLen * NL);
    uint8_t* t = p;
    for (int nt = 0; nt < 5; ++nt)
        for (int i = 0; i < segLen; ++i)
        {
            int j = i;
            for (int s = 0; s < NL; ++s, j += segLen)
                *t++ = (uint8_t)(j >= L ? BIAS : sub_score(nt, nt_code((unsigned char)read[j])) + BIAS);
        }
    return p;
}

static void shl1(uint8_t* v) /* _mm_slli_si128(v, 1) */
{
    for (int k = NL - 1; k > 0; --k)
        v[k] = v[k - 1];
    v[0] = 0;
}

/* 1 when every lane has vF <= sat(vH - gapO), i.e. the lazy-F loop stops (gssw.c:335-339, 363-366) */
static int lazy_done(const uint8_t* vF, const uint8_t* vH)
{
    for (int k = 0; k < NL; ++k)
        if (subs8(vF[k], subs8(vH[k], GAP_OPEN)) != 0)
            return 0;
    return 1;
}

/* gssw_sw_sse2_byte, gssw.c:153-473 (ref_dir = 0).  Returns 255 on byte overflow. */
static int fill_node_byte(const int8_t* ref, int refLen, int L, const uint8_t* prof, const uint8_t* seedH,
                          const uint8_t* seedE, naln* out)
{
    const int segLen = (L + 15) / 16;
    const size_t vb = (size_t)segLen * NL;
    uint8_t* HS = (uint8_t*)calloc(vb, 1);
    uint8_t* HL = (uint8_t*)calloc(vb, 1);
    uint8_t* Hmax = (uint8_t*)calloc(vb, 1);
    uint8_t* E = (uint8_t*)calloc(vb, 1);
    uint8_t* ES = (uint8_t*)calloc(vb, 1);
    uint8_t* FS = (uint8_t*)calloc(vb, 1);
    out->mH = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mE = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mF = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->seedH = (uint8_t*)calloc(vb, 1);
    out->seedE = (uint8_t*)calloc(vb, 1);
    out->is_byte = 1;
    if (seedH) /* gssw.c:215-218 */
    {
        memcpy(E, seedE, vb);
        memcpy(HS, seedH, vb);
    }
    uint8_t max = 0;
    int end_read = L - 1, end_ref = -1, overflow = 0;
    uint8_t vMaxScore[NL] = { 0 }, vMaxMark[NL] = { 0 };

    for (int i = 0; i < refLen; ++i)
    {
        uint8_t e[NL], vF[NL] = { 0 }, vMaxColumn[NL] = { 0 }, vH[NL];
        memcpy(vH, HS + (size_t)(segLen - 1) * NL, NL); /* gssw.c:263-264 */
        shl1(vH);
        const uint8_t* vP = prof + (size_t)ref[i] * segLen * NL;
        uint8_t* pv = HL; /* swap, gssw.c:268-270 */
        HL = HS;
        HS = pv;
        for (int j = 0; j < segLen; ++j) /* gssw.c:273-322 */
        {
            for (int k = 0; k < NL; ++k)
            {
                uint8_t h = adds8(vH[k], vP[j * NL + k]);
                h = subs8(h, BIAS);
                e[k] = E[j * NL + k];
                h = max8(h, e[k]);
                h = max8(h, vF[k]);
                vMaxColumn[k] = max8(vMaxColumn[k], h);
                HS[j * NL + k] = h;
                ES[j * NL + k] = e[k];
                FS[j * NL + k] = vF[k];
                h = subs8(h, GAP_OPEN);
                e[k] = max8(subs8(e[k], GAP_EXT), h);
                vF[k] = max8(subs8(vF[k], GAP_EXT), h);
                E[j * NL + k] = e[k];
                vH[k] = HL[j * NL + k];
            }
        }
        /* lazy-F, gssw.c:325-367 */
        {
            int j = 0;
            memcpy(vH, HS, NL);
            shl1(vF);
            while (!lazy_done(vF, vH))
            {
                for (int k = 0; k < NL; ++k)
                {
                    vH[k] = max8(vH[k], vF[k]);
                    vMaxColumn[k] = max8(vMaxColumn[k], vH[k]);
                    HS[j * NL + k] = vH[k];
                    FS[j * NL + k] = max8(FS[j * NL + k], vF[k]);
                    vF[k] = subs8(vF[k], GAP_EXT);
                }
                if (++j >= segLen)
                {
                    j = 0;
                    shl1(vF);
                }
                memcpy(vH, HS + (size_t)j * NL, NL);
            }
        }
        /* running maximum and end column snapshot, gssw.c:369-387 */
        for (int k = 0; k < NL; ++k)
            vMaxScore[k] = max8(vMaxScore[k], vMaxColumn[k]);
        if (memcmp(vMaxMark, vMaxScore, NL) != 0)
        {
            uint8_t temp = 0;
            memcpy(vMaxMark, vMaxScore, NL);
            for (int k = 0; k < NL; ++k)
                temp = max8(temp, vMaxScore[k]);
            if (temp > max)
            {
                max = temp;
                if (max + BIAS >= 255)
                {
                    overflow = 1;
                    break;
                }
                end_ref = i;
                memcpy(Hmax, HS, vb);
            }
        }
        /* de-stripe, gssw.c:395-430.  Padded positions p >= L land in the next row and are
         * overwritten by it (or fall into allocation slack for the last row): skip them. */
        for (int j = 0; j < segLen; ++j)
            for (int ti = 0; ti < NL; ++ti)
            {
                int p = ti * segLen + j;
                if (p < L)
                {
                    out->mH[(size_t)i * L + p] = HS[j * NL + ti];
                    out->mE[(size_t)i * L + p] = ES[j * NL + ti];
                    out->mF[(size_t)i * L + p] = FS[j * NL + ti];
                }
            }
    }
    memcpy(out->seedE, E, vb); /* gssw.c:442-443 */
    memcpy(out->seedH, HS, vb);
    for (int idx = 0; idx < segLen * NL; ++idx) /* gssw.c:446-454 */
        if (Hmax[idx] == max)
        {
            int temp = idx / NL + idx % NL * segLen;
            if (temp < end_read)
                end_read = temp;
        }
    free(HS);
    free(HL);
    free(Hmax);
    free(E);
    free(ES);
    free(FS);
    out->score1 = overflow ? 255 : max; /* gssw.c:467 */
    out->ref_end1 = end_ref;
    out->read_end1 = end_read;
    return out->score1;
}

/* ------------------------------------------------------------------ 16-bit ("word") mode
 * gssw_graph_fill_internal redoes the whole graph with 8 x int16 lanes as soon as a byte-mode node fill reports
 * overflow (score + bias >= 255, i.e. a score >= 251: gssw.c:380, 467, 4001-4013, 4098-4102). */
enum { NLW = 8 };
static inline int16_t adds16(int16_t a, int16_t b)
{
    int v = (int)a + (int)b;
    return (int16_t)(v > 32767 ? 32767 : (v < -32768 ? -32768 : v));
}
static inline int16_t subsu16(int16_t a, int16_t b) /* _mm_subs_epu16 on the bit patterns */
{
    uint16_t x = (uint16_t)a, y = (uint16_t)b;
    return (int16_t)(x > y ? x - y : 0);
}
static inline int16_t max16(int16_t a, int16_t b) { return a > b ? a : b; }

/* gssw_qP_word, gssw.c:475-498: prof[nt][i][lane] = j >= L ? 0 : mat[nt][read[j]], j = i + lane*segLen */
static int16_t* make_profile_word(const char* read, int L, int segLen)
{
    int16_t* p = (int16_t*)malloc((size_t)5 * segLen * NLW * sizeof(int16_t));
    int16_t* t = p;
    for (int nt = 0; nt < 5; ++nt)
        for (int i = 0; i < segLen; ++i)
        {
            int j = i;
            for (int s = 0; s < NLW; ++s, j += segLen)
                *t++ = (int16_t)(j >= L ? 0 : sub_score(nt, nt_code((unsigned char)read[j])));
        }
    return p;
}

static void shl1w(int16_t* v) /* _mm_slli_si128(v, 2) */
{
    for (int k = NLW - 1; k > 0; --k)
        v[k] = v[k - 1];
    v[0] = 0;
}

/* gssw_sw_sse2_word, gssw.c:527-786 (ref_dir = 0) */
static int fill_node_word(const int8_t* ref, int refLen, int L, const int16_t* prof, const uint16_t* seedH,
                          const uint16_t* seedE, naln* out)
{
    const int segLen = (L + 7) / 8;
    const size_t vn = (size_t)segLen * NLW;
    int16_t* HS = (int16_t*)calloc(vn, sizeof(int16_t));
    int16_t* HL = (int16_t*)calloc(vn, sizeof(int16_t));
    int16_t* Hmax = (int16_t*)calloc(vn, sizeof(int16_t));
    int16_t* E = (int16_t*)calloc(vn, sizeof(int16_t));
    int16_t* ES = (int16_t*)calloc(vn, sizeof(int16_t));
    int16_t* FS = (int16_t*)calloc(vn, sizeof(int16_t));
    out->mH = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mE = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mF = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->seedH16 = (uint16_t*)calloc(vn, sizeof(uint16_t));
    out->seedE16 = (uint16_t*)calloc(vn, sizeof(uint16_t));
    out->is_byte = 0;
    if (seedH) /* gssw.c:588-591 */
    {
        memcpy(E, seedE, vn * sizeof(int16_t));
        memcpy(HS, seedH, vn * sizeof(int16_t));
    }
    uint16_t max = 0;
    int end_read = L - 1, end_ref = 0; /* NB: 0, not -1 as in byte mode (gssw.c:543) */
    int16_t vMaxScore[NLW] = { 0 }, vMaxMark[NLW] = { 0 };

    for (int i = 0; i < refLen; ++i)
    {
        int16_t e[NLW], vF[NLW] = { 0 }, vMaxColumn[NLW] = { 0 }, vH[NLW];
        memcpy(vH, HS + (size_t)(segLen - 1) * NLW, sizeof vH); /* gssw.c:625-626 */
        shl1w(vH);
        const int16_t* vP = prof + (size_t)ref[i] * segLen * NLW;
        int16_t* pv = HL; /* swap, gssw.c:629-635 */
        HL = HS;
        HS = pv;
        for (int j = 0; j < segLen; ++j) /* gssw.c:638-668 */
        {
            for (int k = 0; k < NLW; ++k)
            {
                int16_t h = adds16(vH[k], vP[j * NLW + k]);
                e[k] = E[j * NLW + k];
                h = max16(h, e[k]);
                h = max16(h, vF[k]);
                vMaxColumn[k] = max16(vMaxColumn[k], h);
                HS[j * NLW + k] = h;
                ES[j * NLW + k] = e[k];
                FS[j * NLW + k] = vF[k];
                h = subsu16(h, GAP_OPEN);
                e[k] = max16(subsu16(e[k], GAP_EXT), h);
                E[j * NLW + k] = e[k];
                vF[k] = max16(subsu16(vF[k], GAP_EXT), h);
                vH[k] = HL[j * NLW + k];
            }
        }
        /* lazy-F, gssw.c:671-693: up to 8 passes; stops as soon as no lane has vF > H - gapO */
        {
            int done = 0;
            for (int k8 = 0; k8 < NLW && !done; ++k8)
            {
                shl1w(vF);
                for (int j = 0; j < segLen && !done; ++j)
                {
                    int any = 0;
                    for (int k = 0; k < NLW; ++k)
                    {
                        int16_t h = max16(HS[j * NLW + k], vF[k]);
                        HS[j * NLW + k] = h;
                        FS[j * NLW + k] = max16(FS[j * NLW + k], vF[k]);
                        h = subsu16(h, GAP_OPEN);
                        vF[k] = subsu16(vF[k], GAP_EXT);
                        if (vF[k] > h)
                            any = 1;
                    }
                    if (!any)
                        done = 1;
                }
            }
        }
        /* running maximum and end column snapshot, gssw.c:696-710 (vMaxColumn is NOT updated by the lazy-F loop here) */
        for (int k = 0; k < NLW; ++k)
            vMaxScore[k] = max16(vMaxScore[k], vMaxColumn[k]);
        if (memcmp(vMaxMark, vMaxScore, sizeof vMaxMark) != 0)
        {
            int16_t temp = vMaxScore[0];
            memcpy(vMaxMark, vMaxScore, sizeof vMaxMark);
            for (int k = 1; k < NLW; ++k)
                temp = max16(temp, vMaxScore[k]);
            if ((uint16_t)temp > max)
            {
                max = (uint16_t)temp;
                end_ref = i;
                memcpy(Hmax, HS, vn * sizeof(int16_t));
            }
        }
        /* de-stripe, gssw.c:713-748; padded positions p >= L land in the next row and are overwritten by it */
        for (int j = 0; j < segLen; ++j)
            for (int ti = 0; ti < NLW; ++ti)
            {
                int p = ti * segLen + j;
                if (p < L)
                {
                    out->mH[(size_t)i * L + p] = (uint16_t)HS[j * NLW + ti];
                    out->mE[(size_t)i * L + p] = (uint16_t)ES[j * NLW + ti];
                    out->mF[(size_t)i * L + p] = (uint16_t)FS[j * NLW + ti];
                }
            }
    }
    memcpy(out->seedE16, E, vn * sizeof(int16_t)); /* gssw.c:756-757 */
    memcpy(out->seedH16, HS, vn * sizeof(int16_t));
    for (int idx = 0; idx < segLen * NLW; ++idx) /* gssw.c:761-769 */
        if ((uint16_t)Hmax[idx] == max)
        {
            int temp = idx / NLW + idx % NLW * segLen;
            if (temp < end_read)
                end_read = temp;
        }
    free(HS);
    free(HL);
    free(Hmax);
    free(E);
    free(ES);
    free(FS);
    out->score1 = max;
    out->ref_end1 = end_ref;
    out->read_end1 = end_read;
    return max;
}

/* ------------------------------------------------------------------ model variants (not the reference)
 * The CUDA path does not restate Farrar striping.  It computes the plain affine-gap recurrence
 *     t = max(Hdiag + s, E, 0);  H = max(t, F);  F' = max(F - ge, t - go);
 *     variant 1 (textbook):  E' = max(E - ge, H - go)
 *     variant 2 (kernel):    E' = max(E - ge, t - go)      -- never opens a deletion out of an insertion
 * and claims that H is cell-identical to the reference's mH and that the reference's traceback
 * takes the same decisions on these E/F as on its own striped mE/mF ("equivalence").
 * pgo_set_fill_variant() lets tests/ run the restated traceback over those matrices and compare
 * with oracle/_ref; variant 0 (default) is the faithful restatement. */
static int g_fill_variant = 0;
void pgo_set_fill_variant(int v) { g_fill_variant = v; }

/* seeds of the model variants are linear (row p at index p), 16-bit; word != 0 = "the reference is in 16-bit mode":
 * no overflow report, and is_byte = 0 for the uniqueness scan (aligns_end_at_mult_nodes). */
static int fill_node_model(const int8_t* ref, int refLen, const char* read, int L, const uint16_t* seedH,
                           const uint16_t* seedE, int word, naln* out)
{
    int* Hp = (int*)calloc((size_t)L + 1, sizeof(int));
    int* Hc = (int*)calloc((size_t)L + 1, sizeof(int));
    int* E = (int*)calloc((size_t)L + 1, sizeof(int));
    out->mH = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mE = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->mF = (uint16_t*)calloc((size_t)refLen * L + 1, sizeof(uint16_t));
    out->seedH16 = (uint16_t*)calloc((size_t)L + 1, sizeof(uint16_t));
    out->seedE16 = (uint16_t*)calloc((size_t)L + 1, sizeof(uint16_t));
    out->is_byte = !word;
    for (int p = 0; p < L; ++p)
    {
        Hp[p] = seedH ? seedH[p] : 0;
        E[p] = seedE ? seedE[p] : 0;
    }
    int max = 0, end_ref = word ? 0 : -1, end_read = L - 1;
    for (int i = 0; i < refLen; ++i)
    {
        int F = 0, colmax = 0;
        for (int p = 0; p < L; ++p)
        {
            int s = sub_score(ref[i], nt_code((unsigned char)read[p]));
            int t = (p > 0 ? Hp[p - 1] : 0) + s;
            if (E[p] > t) t = E[p];
            if (t < 0) t = 0;
            int h = t > F ? t : F;
            Hc[p] = h;
            out->mH[(size_t)i * L + p] = (uint16_t)h;
            out->mE[(size_t)i * L + p] = (uint16_t)(E[p] > 0 ? E[p] : 0);
            out->mF[(size_t)i * L + p] = (uint16_t)(F > 0 ? F : 0);
            if (h > colmax) colmax = h;
            int open = (g_fill_variant == 1 ? h : t) - GAP_OPEN;
            E[p] = E[p] - GAP_EXT > open ? E[p] - GAP_EXT : open;
            F = F - GAP_EXT > t - GAP_OPEN ? F - GAP_EXT : t - GAP_OPEN;
        }
        if (colmax > max) /* first column reaching a new maximum; smallest row in it */
        {
            max = colmax;
            end_ref = i;
            for (int p = L - 1; p >= 0; --p)
                if (Hc[p] == max) end_read = p;
        }
        int* sw = Hp; Hp = Hc; Hc = sw;
    }
    for (int p = 0; p < L; ++p)
    {
        out->seedH16[p] = (uint16_t)Hp[p];
        out->seedE16[p] = (uint16_t)(E[p] > 0 ? E[p] : 0);
    }
    free(Hp); free(Hc); free(E);
    out->score1 = max; out->ref_end1 = end_ref; out->read_end1 = end_read;
    return (!word && max >= 251) ? 255 : max;
}

static void naln_clear(naln* a)
{
    free(a->mH);
    free(a->mE);
    free(a->mF);
    free(a->seedH);
    free(a->seedE);
    free(a->seedH16);
    free(a->seedE16);
    memset(a, 0, sizeof(*a));
}

static void naln_free(naln* a, int n)
{
    if (!a)
        return;
    for (int i = 0; i < n; ++i)
        naln_clear(&a[i]);
    free(a);
}

/* One pass of gssw_graph_fill_internal over all nodes in one mode.  Returns 255 when a byte-mode node fill
 * overflowed (the caller redoes everything in word mode), else 0. */
static int graph_fill_mode(const g1* g, const char* read, int L, int word, naln* a, int* max_node)
{
    const int segLen = word ? (L + 7) / 8 : (L + 15) / 16;
    const size_t vn = g_fill_variant ? (size_t)L + 1 : (size_t)segLen * (word ? NLW : NL);
    uint8_t* prof8 = (!word && !g_fill_variant) ? make_profile(read, L, segLen) : NULL;
    int16_t* prof16 = (word && !g_fill_variant) ? make_profile_word(read, L, segLen) : NULL;
    uint8_t* sH = (uint8_t*)malloc(vn);
    uint8_t* sE = (uint8_t*)malloc(vn);
    uint16_t* sH16 = (uint16_t*)malloc(vn * sizeof(uint16_t));
    uint16_t* sE16 = (uint16_t*)malloc(vn * sizeof(uint16_t));
    int max_score = 0, rc = 0;
    *max_node = -1;
    for (int i = 0; i < g->n; ++i)
    {
        /* gssw_create_seed_byte / _word, gssw.c:3897-3962: lane-wise max over predecessors, zeros if none */
        memset(sH, 0, vn);
        memset(sE, 0, vn);
        memset(sH16, 0, vn * sizeof(uint16_t));
        memset(sE16, 0, vn * sizeof(uint16_t));
        for (int k = 0; k < g->npred[i]; ++k)
        {
            const naln* p = &a[g->pred[i][k]];
            for (size_t x = 0; x < vn; ++x)
            {
                if (p->seedH)
                {
                    sH[x] = max8(sH[x], p->seedH[x]);
                    sE[x] = max8(sE[x], p->seedE[x]);
                }
                else
                {
                    sH16[x] = sH16[x] > p->seedH16[x] ? sH16[x] : p->seedH16[x];
                    sE16[x] = sE16[x] > p->seedE16[x] ? sE16[x] : p->seedE16[x];
                }
            }
        }
        int sc;
        if (g_fill_variant)
            sc = fill_node_model(g->num[i], g->len[i], read, L, sH16, sE16, word, &a[i]);
        else if (word)
            sc = fill_node_word(g->num[i], g->len[i], L, prof16, sH16, sE16, &a[i]);
        else
            sc = fill_node_byte(g->num[i], g->len[i], L, prof8, sH, sE, &a[i]);
        if (!word && sc == 255)
        {
            rc = 255; /* gssw.c:4001-4013 */
            break;
        }
        if (sc > max_score) /* gssw.c:4015-4018 (max_score restarts at 0 in the word-mode re-run) */
        {
            *max_node = i;
            max_score = sc;
        }
    }
    free(prof8);
    free(prof16);
    free(sH);
    free(sE);
    free(sH16);
    free(sE16);
    return rc;
}

/* gssw_graph_fill_internal, gssw.c:3964-4028.  *max_node = first node (array order) whose score1
 * strictly exceeds every earlier one, starting from 0 (:4015-4018); -1 when no node scores > 0
 * (the reference then keeps a stale/first max_node whose fresh score1 is 0 -- same observable
 * result: score 0, ref_end -1, empty CIGAR).  Byte mode first (score_size 2, GraphAligner.cpp:220); when a node
 * overflows, every node is refilled in word mode (:4001-4013). */
static int graph_fill(const g1* g, const char* read, int L, naln** out, int* max_node)
{
    naln* a = (naln*)calloc((size_t)g->n, sizeof(naln));
    if (graph_fill_mode(g, read, L, 0, a, max_node) == 255)
    {
        for (int i = 0; i < g->n; ++i)
            naln_clear(&a[i]);
        graph_fill_mode(g, read, L, 1, a, max_node);
    }
    *out = a;
    return PGO_OK;
}

/* ------------------------------------------------------------------ CIGAR containers */

typedef struct
{
    char* type;
    int* len;
    int n, cap;
} cig;

static void cig_reverse(cig* c) /* gssw_reverse_cigar, gssw.c:3726-3745 */
{
    for (int s = 0, e = c->n - 1; s < e; ++s, --e)
    {
        char t = c->type[s];
        c->type[s] = c->type[e];
        c->type[e] = t;
        int l = c->len[s];
        c->len[s] = c->len[e];
        c->len[e] = l;
    }
}

static void cig_push_back(cig* c, char type, int len) /* gssw_cigar_push_back, gssw.c:3679-3695 */
{
    if (c->n > 0 && c->type[c->n - 1] == type)
    {
        c->len[c->n - 1] += len;
        return;
    }
    if (c->n == c->cap)
    {
        c->cap = c->cap ? 2 * c->cap : 8;
        c->type = (char*)realloc(c->type, (size_t)c->cap);
        c->len = (int*)realloc(c->len, sizeof(int) * (size_t)c->cap);
    }
    c->type[c->n] = type;
    c->len[c->n] = len;
    c->n++;
}

static void cig_push_front(cig* c, char type, int len) /* gssw_cigar_push_front, gssw.c:3697-3700 */
{
    cig_reverse(c);
    cig_push_back(c, type, len);
    cig_reverse(c);
}

static char match_op(char refc, char readc) /* gssw.c:1601-1622 */
{
    if (refc == 'N' || readc == 'N')
        return 'N';
    return refc == readc ? 'M' : 'X';
}

/* gssw_alignment_trace_back_byte with final_traceback = 1 and no deflections, gssw.c:1112-1818 */
static void node_trace_back(const g1* g, const naln* a, int n, const char* read, int L, int* score, int* refEnd,
                            int* readEnd, int* gRef, int* gRead, cig* result)
{
    const uint16_t *mH = a[n].mH, *mE = a[n].mE, *mF = a[n].mF;
    const char* ref = g->seq[n];
    int i = *refEnd, j = *readEnd;
    int inE = *gRead, inF = *gRef;
    int scoreHere = inE ? mE[(size_t)L * i + j] : (inF ? mF[(size_t)L * i + j] : mH[(size_t)L * i + j]); /* :1198-1208 */
    long guard = 0;
    while (scoreHere > 0 && i >= 0 && j >= 0)
    {
        if (++guard > 8L * (L + g->len[n] + 8)) /* the reference would spin forever ("Stuck"); never seen */
            break;
        if (inE) /* :1341-1455 */
        {
            if (i > 0)
            {
                if (scoreHere == (int)mH[(size_t)L * (i - 1) + j] - GAP_OPEN)
                {
                    cig_push_back(result, 'D', 1);
                    scoreHere += GAP_OPEN;
                    --i;
                    inE = 0;
                    continue;
                }
                if (scoreHere == (int)mE[(size_t)L * (i - 1) + j] - GAP_EXT)
                {
                    cig_push_back(result, 'D', 1);
                    scoreHere += GAP_EXT;
                    --i;
                    continue;
                }
                continue; /* "Stuck in read gap": asserts are compiled out; loops (guarded above) */
            }
            break; /* i == 0: leave through the left edge, :1440-1448 */
        }
        if (inF) /* :1456-1561 */
        {
            if (j > 0)
            {
                if (scoreHere == (int)mH[(size_t)L * i + (j - 1)] - GAP_OPEN)
                {
                    cig_push_back(result, 'I', 1);
                    scoreHere += GAP_OPEN;
                    --j;
                    inF = 0;
                    continue;
                }
                if (scoreHere == (int)mF[(size_t)L * i + (j - 1)] - GAP_EXT)
                {
                    cig_push_back(result, 'I', 1);
                    scoreHere += GAP_EXT;
                    --j;
                    continue;
                }
            }
            continue; /* stuck (guarded) */
        }
        /* H state, :1562-1801 */
        {
            int s = sub_score(nt_code((unsigned char)ref[i]), nt_code((unsigned char)read[j]));
            if (i > 0 && j > 0)
            {
                if (scoreHere == (int)mH[(size_t)L * (i - 1) + (j - 1)] + s) /* :1591-1637 */
                {
                    cig_push_back(result, match_op(ref[i], read[j]), 1);
                    scoreHere -= s;
                    --i;
                    --j;
                    continue;
                }
            }
            else if (scoreHere == s) /* alignment start, :1655-1690 */
            {
                if (ref[i] == 'N' || read[j] == 'N')
                    cig_push_back(result, 'N', 1);
                else if (ref[i] == read[j])
                    cig_push_back(result, 'M', 1);
                --i;
                --j;
                scoreHere -= s;
                continue;
            }
            if (j > 0 && scoreHere == (int)mF[(size_t)L * i + j]) /* :1709-1729 */
            {
                inF = 1;
                continue;
            }
            if (scoreHere == (int)mE[(size_t)L * i + j]) /* :1747-1768 */
            {
                inE = 1;
                continue;
            }
            if (i == 0) /* :1787-1794 */
                break;
            continue; /* "Stuck in main matrix" (guarded) */
        }
    }
    *score = scoreHere;
    *refEnd = i;
    *readEnd = j;
    *gRef = inF;
    *gRead = inE;
    cig_reverse(result); /* :1815 */
}

typedef struct
{
    int* node;
    cig* c;
    int n, cap;
    int position, score;
} gmap;

static void gmap_free(gmap* m)
{
    for (int i = 0; i < m->n; ++i)
    {
        free(m->c[i].type);
        free(m->c[i].len);
    }
    free(m->node);
    free(m->c);
}

/* gssw_graph_trace_back_internal, num_tracebacks = 1, byte mode: gssw.c:2621-3537 */
static void graph_trace_back(const g1* g, const naln* a, int max_node, const char* read, int L, gmap* gm)
{
    memset(gm, 0, sizeof(*gm));
    if (max_node < 0) /* every score1 == 0: ref_end1 = -1 -> score 0, no nodes, position 0 (:2728-2732, :3528) */
        return;
    int n = max_node;
    int refEnd = a[n].ref_end1, readEnd = a[n].read_end1;
    gm->score = a[n].score1;
    int score = (readEnd < 0 || refEnd < 0) ? 0 : a[n].mH[(size_t)L * refEnd + readEnd];
    int gapInRef = 0, gapInRead = 0;
    int end_soft_clip = L - readEnd - 1; /* :2766-2773 */
    while (score > 0)
    {
        if (gm->n == gm->cap)
        {
            gm->cap = gm->cap ? 2 * gm->cap : 16;
            gm->node = (int*)realloc(gm->node, sizeof(int) * (size_t)gm->cap);
            gm->c = (cig*)realloc(gm->c, sizeof(cig) * (size_t)gm->cap);
        }
        cig* nc = &gm->c[gm->n];
        memset(nc, 0, sizeof(*nc));
        node_trace_back(g, a, n, read, L, &score, &refEnd, &readEnd, &gapInRef, &gapInRead, nc);
        if (end_soft_clip) /* :2814-2820 */
        {
            cig_push_back(nc, 'S', end_soft_clip);
            end_soft_clip = 0;
        }
        gm->node[gm->n++] = n;
        if (score != 0 && refEnd > 0) /* :2824-2828 */
        {
            gm->score = -1;
            break;
        }
        if (score == 0) /* :2836-2844 */
        {
            if (readEnd > -1)
                cig_push_front(nc, 'S', readEnd + 1);
            break;
        }
        int best_prev = -1;
        for (int k = 0; k < g->npred[n]; ++k) /* :2966-3148, first hit wins */
        {
            int cn = g->pred[n][k];
            size_t last = (size_t)L * (g->len[cn] - 1);
            if (!gapInRead)
            {
                int s = sub_score(nt_code((unsigned char)g->seq[n][refEnd]), nt_code((unsigned char)read[readEnd]));
                /* readEnd == 0 reads index last-1 in the reference as well (:2974) */
                int diag = (last + (size_t)readEnd) >= 1 ? a[cn].mH[last + readEnd - 1] : 0;
                if (score == diag + s)
                {
                    best_prev = cn;
                    cig_push_front(nc, match_op(g->seq[n][refEnd], read[readEnd]), 1);
                    score -= s;
                    --readEnd;
                    break;
                }
            }
            else
            {
                if (score == (int)a[cn].mH[last + readEnd] - GAP_OPEN) /* :3089-3110 */
                {
                    best_prev = cn;
                    cig_push_front(nc, 'D', 1);
                    score += GAP_OPEN;
                    gapInRead = 0;
                    break;
                }
                if (score == (int)a[cn].mE[last + readEnd] - GAP_EXT) /* :3122-3136 */
                {
                    best_prev = cn;
                    cig_push_front(nc, 'D', 1);
                    score += GAP_EXT;
                    break;
                }
            }
        }
        if (best_prev >= 0) /* :3486-3499 */
        {
            n = best_prev;
            refEnd = g->len[n] - 1;
        }
        else /* :3500-3517 (assert compiled out) */
        {
            if (readEnd > -1)
                cig_push_front(nc, 'S', readEnd + 1);
            break;
        }
    }
    /* gssw_reverse_graph_cigar, :3526 */
    for (int s = 0, e = gm->n - 1; s < e; ++s, --e)
    {
        int t = gm->node[s];
        gm->node[s] = gm->node[e];
        gm->node[e] = t;
        cig c = gm->c[s];
        gm->c[s] = gm->c[e];
        gm->c[e] = c;
    }
    gm->position = refEnd + 1 < 0 ? 0 : refEnd + 1; /* :3528 */
}

/* alignsEndAtMultNodes, GraphAligner.cpp:170-212 (gssw node id == graph node id: Graph{n,false}).
 * The reference scans `(uint8_t*)alignment->mH` over len*L BYTES whatever the mode (:177-186).  In byte mode that is
 * the matrix.  In word mode it is the first half of the int16 matrix seen as bytes (little-endian: low byte, high
 * byte, ...), each compared with the uint16 top score: nothing matches a top score >= 256, and for 251..255 only
 * the low bytes of cells with linear index < ceil(len*L/2) can.  Restated literally. */
static int aligns_end_at_mult_nodes(const g1* g, const naln* a, int max_node, int L)
{
    int top = max_node >= 0 ? a[max_node].score1 : 0;
    int cnt = 0;
    for (int n = 0; n < g->n; ++n)
    {
        int found = 0;
        size_t tot = (size_t)g->len[n] * L;
        if (a[n].is_byte)
        {
            for (size_t x = 0; x < tot && !found; ++x)
                found = a[n].mH[x] == top;
        }
        else
        {
            for (size_t x = 0; x < tot && !found; ++x)
            {
                uint16_t cell = a[n].mH[x >> 1];
                uint8_t byte = (uint8_t)((x & 1) ? (cell >> 8) : (cell & 0xff));
                found = byte == top;
            }
        }
        cnt += found;
        if (cnt > 1)
            return 1;
    }
    return 0;
}

/* GraphAlignerImpl::extractCigar, GraphAligner.cpp:88-108 */
static int cigar_string(const gmap* gm, char* out, int cap)
{
    int len = 0;
    char tmp[32];
#define EMIT(str)                                                                                                      \
    for (const char* q_ = (str); *q_; ++q_)                                                                            \
    {                                                                                                                  \
        if (out && len < cap - 1)                                                                                      \
            out[len] = *q_;                                                                                            \
        ++len;                                                                                                         \
    }
    for (int i = 0; i < gm->n; ++i)
    {
        snprintf(tmp, sizeof tmp, "%d[", gm->node[i]);
        EMIT(tmp);
        for (int k = 0; k < gm->c[i].n; ++k)
        {
            snprintf(tmp, sizeof tmp, "%d%c", gm->c[i].len[k], gm->c[i].type[k]);
            EMIT(tmp);
        }
        EMIT("]");
    }
#undef EMIT
    if (out && cap > 0)
        out[len < cap - 1 ? len : cap - 1] = 0;
    return len;
}

/* GraphAlignerImpl::alignString, GraphAligner.cpp:214-227 (str already upper-cased by caller) */
static int align_string(const g1* g, const char* str, int L, gmap* gm, int* multi)
{
    naln* a = NULL;
    int max_node = -1;
    int rc = graph_fill(g, str, L, &a, &max_node);
    if (rc != PGO_OK)
    {
        naln_free(a, g->n);
        memset(gm, 0, sizeof(*gm));
        return rc;
    }
    graph_trace_back(g, a, max_node, str, L, gm);
    *multi = aligns_end_at_mult_nodes(g, a, max_node, L);
    naln_free(a, g->n);
    return PGO_OK;
}

static int fill_trace_impl(const pgo_graph* G, int reversed_graph, const char* read, int L, int32_t* node_stats,
                           uint8_t* mats8, uint16_t* mats16, int32_t* res3, int32_t* multi, char* cigar, int cigar_cap)
{
    const g1* g = reversed_graph ? &G->rev : &G->fwd;
    naln* a = NULL;
    int max_node = -1;
    int rc = graph_fill(g, read, L, &a, &max_node);
    if (rc != PGO_OK)
    {
        naln_free(a, g->n);
        return rc;
    }
    size_t off = 0;
    for (int i = 0; i < g->n; ++i)
    {
        node_stats[4 * i + 0] = a[i].score1;
        node_stats[4 * i + 1] = a[i].ref_end1;
        node_stats[4 * i + 2] = a[i].read_end1;
        node_stats[4 * i + 3] = a[i].is_byte;
        size_t sz = (size_t)g->len[i] * L;
        const uint16_t* m3[3] = { a[i].mH, a[i].mE, a[i].mF };
        for (int k = 0; k < 3; ++k)
            for (size_t x = 0; x < sz; ++x)
            {
                if (mats8 && a[i].is_byte)
                    mats8[off + k * sz + x] = (uint8_t)m3[k][x];
                if (mats16)
                    mats16[off + k * sz + x] = m3[k][x];
            }
        off += 3 * sz;
    }
    gmap gm;
    graph_trace_back(g, a, max_node, read, L, &gm);
    res3[0] = max_node;
    res3[1] = gm.position;
    res3[2] = gm.score;
    if (multi)
        *multi = aligns_end_at_mult_nodes(g, a, max_node, L);
    rc = cigar_string(&gm, cigar, cigar_cap);
    gmap_free(&gm);
    naln_free(a, g->n);
    return rc;
}

int pgo_fill_trace(const pgo_graph* G, int reversed_graph, const char* read, int L, int32_t* node_stats, uint8_t* mats,
                   int32_t* res3, int32_t* multi, char* cigar, int cigar_cap)
{
    return fill_trace_impl(G, reversed_graph, read, L, node_stats, mats, NULL, res3, multi, cigar, cigar_cap);
}

int pgo_fill_trace16(const pgo_graph* G, int reversed_graph, const char* read, int L, int32_t* node_stats,
                     uint16_t* mats, int32_t* res3, int32_t* multi, char* cigar, int cigar_cap)
{
    return fill_trace_impl(G, reversed_graph, read, L, node_stats, NULL, mats, res3, multi, cigar, cigar_cap);
}

static char complement_base(char b) /* graph-tools src/graphutils/SequenceOperations.cpp:66-81 (case-sensitive) */
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

/* GraphAligner::alignRead, GraphAligner.cpp:308-404 */
int pgo_align_read(const pgo_graph* G, const char* bases, int L, int is_reverse_strand, unsigned flags, int32_t* out6,
                   char* out_bases, char* cigar, int cigar_cap)
{
    if (!G || L <= 0)
        return PGO_E_ARG;
    char* s_fwd = (char*)malloc((size_t)L + 1);
    char* rev_cmp = (char*)malloc((size_t)L + 1); /* reverseComplement(read.bases()), :315 */
    char* tmp = (char*)malloc((size_t)L + 1);
    for (int i = 0; i < L; ++i)
    {
        s_fwd[i] = (char)toupper((unsigned char)bases[i]); /* alignString: toUpper(str), :218 */
        rev_cmp[i] = complement_base(bases[L - 1 - i]);
    }
    s_fwd[L] = rev_cmp[L] = tmp[L] = 0;

    gmap gm_fwd, gm_rev, gm_tmp;
    memset(&gm_rev, 0, sizeof gm_rev);
    int fwd_multi = 0, rev_multi = 0, rfwd_multi = 0, rrev_multi = 0, have_rev = 0;
    int rc = align_string(&G->fwd, s_fwd, L, &gm_fwd, &fwd_multi); /* :317-318 */
    if (rc == PGO_OK && (flags & PGO_AF_BOTH_STRANDS))              /* :319-321 */
    {
        for (int i = 0; i < L; ++i)
            tmp[i] = (char)toupper((unsigned char)rev_cmp[i]);
        rc = align_string(&G->fwd, tmp, L, &gm_rev, &rev_multi);
        have_rev = rc == PGO_OK;
    }
    if (rc == PGO_OK && (flags & PGO_AF_REVERSE_GRAPH)) /* :326-338 */
    {
        for (int i = 0; i < L; ++i)
            tmp[i] = (char)toupper((unsigned char)bases[L - 1 - i]);
        rc = align_string(&G->rev, tmp, L, &gm_tmp, &rfwd_multi);
        gmap_free(&gm_tmp);
        if (rc == PGO_OK && (flags & PGO_AF_BOTH_STRANDS))
        {
            /* reverseComplement(bases_rev) = complement of the bases in original order */
            for (int i = 0; i < L; ++i)
                tmp[i] = (char)toupper((unsigned char)complement_base(bases[i]));
            rc = align_string(&G->rev, tmp, L, &gm_tmp, &rrev_multi);
            gmap_free(&gm_tmp);
        }
    }
    if (rc == PGO_OK)
    {
        int fwd_unique = !fwd_multi && !rfwd_multi; /* :340-341 */
        int rev_unique = !rev_multi && !rrev_multi;
        int return_reverse = 0; /* :344-356 */
        if (!fwd_unique && rev_unique && have_rev)
            return_reverse = 1;
        else if (fwd_unique && !rev_unique)
            return_reverse = 0;
        else if (have_rev)
            return_reverse = (int16_t)gm_fwd.score < (int16_t)gm_rev.score;
        const gmap* gm = return_reverse ? &gm_rev : &gm_fwd;
        int uniq = return_reverse ? rev_unique : fwd_unique;
        out6[0] = gm->position;
        out6[1] = (int16_t)gm->score;
        out6[2] = uniq;
        out6[3] = uniq ? 60 : 0;
        out6[4] = (is_reverse_strand != 0) != return_reverse; /* :358-359 */
        out6[5] = 0;
        if (flags & PGO_AF_CIGAR)
            out6[5] = cigar_string(gm, cigar, cigar_cap);
        else if (cigar && cigar_cap > 0)
            cigar[0] = 0;
        if (out_bases)
            memcpy(out_bases, return_reverse ? rev_cmp : bases, (size_t)L); /* :375 */
    }
    gmap_free(&gm_fwd);
    gmap_free(&gm_rev);
    free(s_fwd);
    free(rev_cmp);
    free(tmp);
    return rc;
}

int pgo_align_batch(const pgo_graph* G, int n_reads, const char* blob, const int32_t* off, const uint8_t* is_rev,
                    unsigned flags, int32_t* out6, char* out_bases_blob, char* cigars, int cigar_stride)
{
    int worst = PGO_OK;
    for (int i = 0; i < n_reads; ++i)
    {
        int rc = pgo_align_read(G, blob + off[i], off[i + 1] - off[i], is_rev ? is_rev[i] : 0, flags, out6 + 6 * i,
                                out_bases_blob ? out_bases_blob + off[i] : NULL,
                                cigars ? cigars + (size_t)i * cigar_stride : NULL, cigar_stride);
        if (rc != PGO_OK)
        {
            worst = rc;
            memset(out6 + 6 * i, 0xFF, 6 * sizeof(int32_t));
        }
    }
    return worst;
}

/* readfilters::BadAlign (src/c++/lib/paragraph/readfilters/BadAlign.hh:62-73): decode the graph CIGAR, sum the
 * clipped query bases, filter when aligned < round(frac * queryLength).  queryLength counts M X N I S ops.
 * Returns 1 if filtered; *clipped_out receives the number of soft-clipped query bases. */
int pgo_bad_align(const char* cigar, double bad_align_frac, int* clipped_out)
{
    long qlen = 0, clipped = 0, num = 0;
    for (const char* p = cigar; *p; ++p)
    {
        if (*p >= '0' && *p <= '9')
            num = num * 10 + (*p - '0');
        else if (*p == '[') /* the number before '[' is the node id */
            num = 0;
        else if (*p == ']')
            num = 0;
        else
        {
            if (*p == 'M' || *p == 'X' || *p == 'N' || *p == 'I' || *p == 'S')
                qlen += num;
            if (*p == 'S')
                clipped += num;
            num = 0;
        }
    }
    if (clipped_out)
        *clipped_out = (int)clipped;
    double thr = bad_align_frac * (double)qlen;
    double r = (double)(long)(thr + 0.5); /* round(): half away from zero, thr >= 0 */
    return (double)(qlen - clipped) < r;
}
