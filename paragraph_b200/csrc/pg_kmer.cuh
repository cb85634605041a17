// paragraph_b200 -- the k-mer stage of the cascade: grm::KmerAligner<K>
// (src/c++/lib/grm/KmerAligner.cpp; second stage of grm::CompositeAligner, lib/grm/CompositeAligner.cpp:105-126; K = 16).
// Shared verbatim between the sm_100a kernel (pg_kernels.cu: pg_kmer_kernel, one warp per read) and the CPU emulator
// (tests/emu): the algorithm is written once for a GROUP of lanes -- 32 on the device, 1 in the emulator -- with the
// group primitives (barrier, sum, bit set) abstracted below.
//
// Reference algorithm: gapless alignment of the read, on both strands, to the sequence of one of the graph's PATHS
// (the "paths" of the graph JSON, GraphInput.cpp:168-197).  Every k-mer the read shares with a path proposes the
// offset that lines the two occurrences up (:246-273; a k-mer that repeats in the READ only counts at its first
// occurrence, a consequence of the single merge walk over the two sorted k-mer lists); the distinct offsets of a
// (path, strand), in ascending order, are scored by their number of mismatching characters and pushed into a bounded
// heap that keeps the n_paths + 1 candidates with fewest mismatches (:283-293); pickBest (:479-517) maps the read to
// the FIRST minimal element of the heap array if it has at most two mismatches -- uniquely, unless an equally good
// candidate that follows it in the array gives a different (position, CIGAR): then mapq 0 / BAD_ALIGN, and the read
// goes on to the next stage with the bases this stage left behind.
// The heap is std::push_heap / std::pop_heap of GNU libstdc++ under a comparator that only sees the mismatch count, so
// which of several equal candidates survives and comes first is decided by that implementation; it is restated here
// operation by operation (kmer_push_heap / kmer_pop_heap) -- see oracle/pg_oracle_kmer.c for the pinning.
#pragma once
#include "pg_core.cuh"

namespace pg
{

constexpr int KMER_WIN = 2048;     // offsets handled per pass over the seeds (bits of the candidate bitmap)
constexpr int KMER_MAX_PATHS = 62; // heap capacity n_paths + 2 <= 64 entries of shared memory

struct KmerPathDev // one path of one site
{
    int32_t seq_off;   // byte offset of the path sequence (node sequences as given, concatenated)
    int32_t len;
    int32_t n_nodes;
    int32_t nodes_off; // int offset: node ids [n_nodes], then the offset of each node in the path sequence [n_nodes]
    int32_t kmers_off; // offset of the path's k-mers (value, position), sorted by (value, position)
    int32_t n_kmers;
};
struct KmerSiteDev
{
    int32_t n_paths;
    int32_t path0; // first KmerPathDev of the site
};
struct KmerPos
{
    uint32_t kmer;
    int32_t pos;
};
struct KmerView
{
    const KmerSiteDev* sites;
    const KmerPathDev* paths;
    const uint8_t* seqs;
    const int32_t* nodes;
    const KmerPos* kmers;
    int32_t k;
};
struct KmerCand
{
    int32_t path, pos, rev;
    uint32_t mm;
};
struct KmerScratch // per group, in shared memory (device) or on the heap (emulator)
{
    uint8_t* seq[2];    // [L] the read as given / graphtools::reverseComplement of it
    uint32_t* km[2];    // [L] k-mer value starting at position i
    uint8_t* valid[2];  // [L] the k characters from i on are all ACGT (any case)
    uint8_t* first[2];  // [L] valid, and no earlier position of the read has the same k-mer
    uint32_t* bitmap;   // [KMER_WIN / 32]
    KmerCand* heap;     // [heap_cap]
    uint32_t* ops_best; // [ops_cap]
    uint32_t* ops_tmp;  // [ops_cap]
    int ops_cap;
};
struct KmerResult
{
    int status; // 0 unmapped (no candidate with <= 2 mismatches), 1 mapped, 2 mapped but not unique (BAD_ALIGN)
    int pos, score, rev, clipped, n_ops;
};

// ---- group primitives -------------------------------------------------------------------------------------------
#if defined(__CUDA_ARCH__)
PG_HD void kg_sync() { __syncwarp(); }
PG_HD int kg_sum(int v)
{
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1)
        v += __shfl_xor_sync(0xffffffffu, v, d);
    return v;
}
PG_HD void kg_set_bit(uint32_t* w, uint32_t bit) { atomicOr(w, bit); }
PG_HD int kg_ffs(uint32_t x) { return __ffs((int)x) - 1; }
#else
PG_HD void kg_sync() {}
PG_HD int kg_sum(int v) { return v; }
PG_HD void kg_set_bit(uint32_t* w, uint32_t bit) { *w |= bit; }
PG_HD int kg_ffs(uint32_t x)
{
    int b = 0;
    while (!((x >> b) & 1u))
        ++b;
    return b;
}
#endif

// oligo::Translator<> (Nucleotides.hh:59-341): A/a 0, C/c 1, G/g 2, T/t 3, everything else invalid
PG_HD int kmer_base_value(uint8_t c)
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

// ---- GNU libstdc++ <bits/stl_heap.h> with Candidate::lessMismatches (KmerAligner.cpp:78-83) -----------------------
PG_HD void kmer_sift_up(KmerCand* first, int hole, int top, KmerCand value) // std::__push_heap
{
    int parent = (hole - 1) / 2;
    while (hole > top && first[parent].mm < value.mm)
    {
        first[hole] = first[parent];
        hole = parent;
        parent = (hole - 1) / 2;
    }
    first[hole] = value;
}
PG_HD void kmer_push_heap(KmerCand* first, int n) { kmer_sift_up(first, n - 1, 0, first[n - 1]); } // new element = first[n - 1]
PG_HD void kmer_pop_heap(KmerCand* first, int n) // afterwards the element with most mismatches is first[n - 1]
{
    if (n <= 1)
        return;
    const KmerCand value = first[n - 1];
    first[n - 1] = first[0];
    const int len = n - 1; // std::__adjust_heap(first, 0, len, value)
    int hole = 0, child = 0;
    while (child < (len - 1) / 2)
    {
        child = 2 * (child + 1);
        if (first[child].mm < first[child - 1].mm)
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
    kmer_sift_up(first, hole, 0, value);
}
PG_HD int kmer_min_element(const KmerCand* c, int from, int n) // std::min_element: the first minimal one; n if empty
{
    if (from >= n)
        return n;
    int best = from;
    for (int i = from + 1; i < n; ++i)
        if (c[i].mm < c[best].mm)
            best = i;
    return best;
}

// updateAlignment + buildCigar + makeCigarBit (KmerAligner.cpp:318-478) as op words: leading / trailing 'N' of the PATH
// (N-filled source / sink) are soft-clipped, then per path node the runs of M / N / X (equal characters -> M, else N
// if either is 'N', else X); score = matching bases; graph_pos = offset in the first node after clipping.
// Returns the number of op words (written while they fit `cap`).
PG_HD int kmer_build_ops(const KmerView& v, const KmerPathDev& kp, int pos0, const uint8_t* q, int L, uint32_t* ops, int cap,
                         int& gpos, int& score, int& clipped)
{
    const uint8_t* ref = v.seqs + kp.seq_off + pos0;
    const int32_t* node_id = v.nodes + kp.nodes_off;
    const int32_t* node_start = node_id + kp.n_nodes;
    int left = 0;
    while (left < L && ref[left] == 'N')
        ++left;
    int right = 0;
    while (right < L - left && ref[L - 1 - right] == 'N')
        ++right;
    const int pos = pos0 + left;
    int sn = kp.n_nodes - 1; // findStartNode (:162-178)
    for (int i = 0; i < kp.n_nodes; ++i)
        if (node_start[i] >= pos)
        {
            sn = node_start[i] > pos ? i - 1 : i;
            break;
        }
    if (sn < 0)
        sn = 0;
    int this_start = pos - node_start[sn], left_len = L - left - right, lclip = left, n = 0;
    gpos = this_start;
    score = 0;
    clipped = left + right;
    const uint8_t* qq = q + left;
    const uint8_t* seq = v.seqs + kp.seq_off;
    for (int i = sn; i < kp.n_nodes && left_len > 0; ++i)
    {
        int this_len = left_len;
        if (i + 1 < kp.n_nodes)
        {
            const int room = node_start[i + 1] - node_start[i] - this_start;
            if (room < this_len)
                this_len = room;
        }
        if (this_len > 0)
        {
            const uint8_t* r = seq + this_start + node_start[i];
            const int node = node_id[i];
            if (lclip)
            {
                if (n < cap)
                    ops[n] = cigar_word(node, OP_S, lclip);
                ++n;
                lclip = 0;
            }
            int last = -1, run = 0;
            for (int x = 0; x < this_len; ++x)
            {
                const uint8_t s = r[x], b = qq[x];
                const int op = s == b ? OP_M : ((s == 'N' || b == 'N') ? OP_N : OP_X);
                if (op != last)
                {
                    if (run)
                    {
                        if (n < cap)
                            ops[n] = cigar_word(node, last, run);
                        ++n;
                        if (last == OP_M)
                            score += run;
                    }
                    last = op;
                    run = 0;
                }
                ++run;
            }
            if (run)
            {
                if (n < cap)
                    ops[n] = cigar_word(node, last, run);
                ++n;
                if (last == OP_M)
                    score += run;
            }
            qq += this_len;
            if (right && this_len == left_len)
            {
                if (n < cap)
                    ops[n] = cigar_word(node, OP_S, right);
                ++n;
            }
        }
        left_len -= this_len;
        this_start = 0;
    }
    return n;
}

// KmerAlignerImpl::alignRead (KmerAligner.cpp:519-536) for one read by one group of `nl` lanes (this one is `lane`).
// All lanes return the same result; the op words of a mapped read are in sc.ops_best[0 .. n_ops).
PG_HD KmerResult kmer_align_read(const KmerView& v, int site, const uint8_t* bases, int L, int lane, int nl, const KmerScratch& sc)
{
    const int k = v.k;
    const KmerSiteDev ks = v.sites[site];
    KmerResult res;
    res.status = res.pos = res.score = res.rev = res.clipped = res.n_ops = 0;
    for (int j = lane; j < L; j += nl)
    {
        sc.seq[0][j] = bases[j];
        sc.seq[1][j] = complement_base(bases[L - 1 - j]);
    }
    kg_sync();
    for (int s = 0; s < 2; ++s)
        for (int i = lane; i < L; i += nl) // makeKmers (:120-133): 2 bits per base, windows with an invalid character skipped
        {
            uint32_t val = 0;
            bool ok = i + k <= L;
            for (int x = 0; ok && x < k; ++x)
            {
                const int b = kmer_base_value(sc.seq[s][i + x]);
                ok = b < 4;
                val = (val << 2) | (uint32_t)(b & 3);
            }
            sc.km[s][i] = val;
            sc.valid[s][i] = ok ? 1 : 0;
        }
    kg_sync();
    for (int s = 0; s < 2; ++s)
        for (int i = lane; i < L; i += nl)
        {
            bool first = sc.valid[s][i] != 0;
            for (int j = 0; first && j < i; ++j)
                first = !(sc.valid[s][j] && sc.km[s][j] == sc.km[s][i]);
            sc.first[s][i] = first ? 1 : 0;
        }
    kg_sync();
    const int capacity = ks.n_paths + 2; // :306-309
    int n = 0;                           // candidates in the heap (every lane counts along; lane 0 owns the array)
    for (int p = 0; p < ks.n_paths; ++p)
    {
        const KmerPathDev kp = v.paths[ks.path0 + p];
        const KmerPos* pk_ = v.kmers + kp.kmers_off;
        const uint8_t* pseq = v.seqs + kp.seq_off;
        const int maxoff = kp.len - L; // candidates that overhang the path are ignored (:263-267)
        for (int s = 0; s < 2 && maxoff >= 0; ++s)
            for (int w0 = 0; w0 <= maxoff; w0 += KMER_WIN)
            {
                for (int x = lane; x < KMER_WIN / 32; x += nl)
                    sc.bitmap[x] = 0u;
                kg_sync();
                for (int i = lane; i < L; i += nl)
                {
                    if (!sc.first[s][i])
                        continue;
                    const uint32_t val = sc.km[s][i];
                    int lo = 0, hi = kp.n_kmers; // first path k-mer >= val
                    while (lo < hi)
                    {
                        const int mid = (lo + hi) >> 1;
                        if (pk_[mid].kmer < val)
                            lo = mid + 1;
                        else
                            hi = mid;
                    }
                    for (; lo < kp.n_kmers && pk_[lo].kmer == val; ++lo)
                    {
                        const int off = pk_[lo].pos - i;
                        if (off >= w0 && off < w0 + KMER_WIN && off <= maxoff)
                            kg_set_bit(sc.bitmap + ((off - w0) >> 5), 1u << ((off - w0) & 31));
                    }
                }
                kg_sync();
                for (int wi = 0; wi < KMER_WIN / 32; ++wi) // the distinct offsets, ascending (:274-281)
                {
                    uint32_t word = sc.bitmap[wi];
                    while (word)
                    {
                        const int b = kg_ffs(word);
                        word &= word - 1u;
                        const int off = w0 + wi * 32 + b;
                        int mm = 0; // countMismatches (:232-240): plain character comparison
                        for (int x = lane; x < L; x += nl)
                            mm += sc.seq[s][x] != pseq[off + x];
                        mm = kg_sum(mm);
                        if (lane == 0)
                        {
                            KmerCand c;
                            c.path = p;
                            c.pos = off;
                            c.rev = s;
                            c.mm = (uint32_t)mm;
                            sc.heap[n] = c;
                            kmer_push_heap(sc.heap, n + 1);
                            if (capacity == n + 1) // full: the candidate with most mismatches goes (:286-292)
                                kmer_pop_heap(sc.heap, n + 1);
                        }
                        if (capacity != n + 1)
                            ++n;
                    }
                }
                kg_sync();
            }
    }
    kg_sync();
    if (n == 0)
        return res;
    // pickBest (:479-517).  Every lane runs it on the shared heap (read-only from here on); lane 0 writes the op words.
    const int b = kmer_min_element(sc.heap, 0, n);
    const KmerCand best = sc.heap[b];
    if (best.mm > 2u)
        return res;
    int gpos = 0, score = 0, clipped = 0;
    int nb = 0;
    if (lane == 0)
        nb = kmer_build_ops(v, v.paths[ks.path0 + best.path], best.pos, sc.seq[best.rev], L, sc.ops_best, sc.ops_cap, gpos, score,
                            clipped);
    int status = 1;
    if (lane == 0)
        for (int s2 = kmer_min_element(sc.heap, b + 1, n); s2 < n; s2 = kmer_min_element(sc.heap, s2 + 1, n))
        {
            const KmerCand c2 = sc.heap[s2];
            if (c2.mm != best.mm)
                break; // no more as good candidates
            int gpos2 = 0, score2 = 0, clipped2 = 0;
            const int n2 = kmer_build_ops(v, v.paths[ks.path0 + c2.path], c2.pos, sc.seq[c2.rev], L, sc.ops_tmp, sc.ops_cap, gpos2,
                                          score2, clipped2);
            bool same = gpos2 == gpos && n2 == nb;
            for (int x = 0; same && x < nb && x < sc.ops_cap; ++x)
                same = sc.ops_tmp[x] == sc.ops_best[x];
            if (!same)
            {
                status = 2;
                break;
            }
        }
#if defined(__CUDA_ARCH__)
    status = __shfl_sync(0xffffffffu, status, 0);
    gpos = __shfl_sync(0xffffffffu, gpos, 0);
    score = __shfl_sync(0xffffffffu, score, 0);
    clipped = __shfl_sync(0xffffffffu, clipped, 0);
    nb = __shfl_sync(0xffffffffu, nb, 0);
#endif
    res.status = status;
    res.pos = gpos;
    res.score = score;
    res.rev = best.rev;
    res.clipped = clipped;
    res.n_ops = nb;
    return res;
}

} // namespace pg
