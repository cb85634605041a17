// paragraph_b200 -- core of the read->graph alignment path, shared verbatim between the sm_100a kernels
// (pg_kernels.cu) and the host-side lane emulator used by the CPU test-suite (tests/emu/pg_emu.cpp).
//
// What is computed (reference: external/gssw/gssw.c gssw_graph_fill / gssw_graph_trace_back driven by
// src/c++/lib/grm/GraphAligner.cpp:214-404):
//   * affine-gap local Smith-Waterman of a read against the topologically ordered nodes of a variant
//     graph; column -1 of a node = element-wise max over the last columns of its predecessors
//     (gssw.c:3897-3931), gaps in the reference do not cross nodes (gssw.h:61-65);
//   * best cell = first cell in (node, column, row) order holding the global maximum (gssw.c:378-386,
//     446-454, 4015-4018); uniqueness = the maximum occurs in exactly one node (GraphAligner.cpp:170-212);
//   * traceback with the reference's positional decision order (gssw.c:1112-1818, 2621-3537).
//
// How it is laid out on a warp ("wavefront"): lane t owns read rows [R*t, R*t+R); at step k it processes
// graph column q = k - t, so the vertical dependency (F, and the diagonal into the lane's first row)
// comes from lane t-1's previous step through one warp shuffle.  Two alignment problems that share the
// column sequence are packed in the two int16 halves of every register (forward read / reverse-
// complemented read), so the recurrence runs on the packed-int16 DPX instructions (VIADDMNMX.S16x2,
// VIMNMX3.S16x2).  Recurrence per cell (d = H of the previous column one row up, s = substitution score):
//     t  = max(d + s, E, 0)         E' = max(E - ge, t - go)
//     H  = max(t, F)                F' = max(F - ge, t - go)      (F runs down the rows of one column)
// H is cell-identical to gssw's mH; E/F differ from gssw's striped mE/mF only in cells where the
// traceback cannot look (DESIGN.md "equivalence", proven by tests/test_model_equivalence.py).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define PG_HD __host__ __device__ __forceinline__
#define PG_HD_COLD __host__ __device__ __forceinline__
#define PG_UNROLL _Pragma("unroll")
#define PG_NOUNROLL _Pragma("unroll 1")
#ifndef PG_OUTLINE_GENERAL
#define PG_OUTLINE_GENERAL 0 // 1: seed_general as a real call -- measured: 135 registers, a stack frame, fill + 14 %
#endif
#if PG_OUTLINE_GENERAL
#define PG_COLD_CALL __host__ __device__ __noinline__
#else
#define PG_COLD_CALL __host__ __device__ __forceinline__
#endif
#else
#define PG_HD inline
#define PG_HD_COLD inline
#define PG_UNROLL
#define PG_NOUNROLL
#define PG_COLD_CALL inline
#endif

namespace pg
{

constexpr int GAP_OPEN = 6; // GraphAligner.cpp:231
constexpr int GAP_EXT = 1;  // GraphAligner.cpp:232
constexpr int NEG = -16384; // substitution score of sentinel columns / padded rows
#ifndef PG_CK
#define PG_CK 32 // 16 / 32 / 64 measured on the B200: 32 halves the scratch written by the fill for 16 steps of run-up in half of the tracebacks
#endif
constexpr int CK = PG_CK;   // checkpoint interval of the fill, in wavefront steps
#ifndef PG_TS
#define PG_TS 16
#endif
constexpr int TS = PG_TS;   // traceback tile size in steps: a tile is recomputed from the checkpoint at or before its first
                            // step (CK a multiple of TS: the steps in between are run without keeping their cells)
static_assert(CK % TS == 0, "the checkpoint interval must be a multiple of the tile size");
constexpr int SENT = 32;    // sentinel columns (code 5) before and after every column sequence
constexpr int NCODE = 6;    // A C G T other sentinel
constexpr int MAX_READ_LEN = 1024; // R = 32 rows per lane x 32 lanes
// gssw fills in 8-bit mode until a score reaches 251 (score + bias(4) >= 255, gssw.c:380, 467) and then redoes the
// whole graph in 16-bit mode (gssw.c:4001-4013).  Scores are the same in both modes (this path computes in int16
// throughout); what changes is GraphAligner's uniqueness scan (finalize_task) and what fits a byte in the scratch
// layouts: geometries whose reads can exceed BYTE_MAX_SCORE use the WIDE checkpoint / tile-cell packing.
constexpr int BYTE_MAX_SCORE = 250;
constexpr int BYTE_MAX_READ_LEN = 250; // longest read a non-WIDE geometry takes
PG_HD constexpr bool is_wide(int R, int W) { return R * W > 256; }

// ---------------------------------------------------------------------------------------------
// packed int16x2 arithmetic: DPX on the device, plain C on the host (emulator)
// ---------------------------------------------------------------------------------------------
PG_HD uint32_t pk(int lo, int hi) { return (uint32_t)(uint16_t)(int16_t)lo | ((uint32_t)(uint16_t)(int16_t)hi << 16); }
PG_HD int lo16(uint32_t x) { return (int)(int16_t)(uint16_t)(x & 0xffffu); }
PG_HD int hi16(uint32_t x) { return (int)(int16_t)(uint16_t)(x >> 16); }
PG_HD int half16(uint32_t x, int h) { return h ? hi16(x) : lo16(x); }
PG_HD int imax0(int a) { return a > 0 ? a : 0; }

#if defined(__CUDA_ARCH__)
PG_HD uint32_t addmax_relu2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2_relu(a, b, c); }
PG_HD uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c) { return __viaddmax_s16x2(a, b, c); }
PG_HD uint32_t max2(uint32_t a, uint32_t b) { return __vmaxs2(a, b); }
PG_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return __vimax3_s16x2(a, b, c); }
PG_HD uint32_t add2(uint32_t a, uint32_t b) { return __vadd2(a, b); }
PG_HD uint32_t sub2(uint32_t a, uint32_t b) { return __vsub2(a, b); }
PG_HD uint32_t min2(uint32_t a, uint32_t b) { return __vmins2(a, b); }
// max(a, b) per half plus "a >= b" per half (one VIMNMX.S16x2 with two predicate outputs)
PG_HD uint32_t max2_ge(uint32_t a, uint32_t b, bool& hi_ge, bool& lo_ge) { return __vibmax_s16x2(a, b, &hi_ge, &lo_ge); }
#else
PG_HD int imax_(int a, int b) { return a > b ? a : b; }
PG_HD uint32_t addmax_relu2(uint32_t a, uint32_t b, uint32_t c)
{
    return pk(imax_(imax_(lo16(a) + lo16(b), lo16(c)), 0), imax_(imax_(hi16(a) + hi16(b), hi16(c)), 0));
}
PG_HD uint32_t addmax2(uint32_t a, uint32_t b, uint32_t c)
{
    return pk(imax_(lo16(a) + lo16(b), lo16(c)), imax_(hi16(a) + hi16(b), hi16(c)));
}
PG_HD uint32_t max2(uint32_t a, uint32_t b) { return pk(imax_(lo16(a), lo16(b)), imax_(hi16(a), hi16(b))); }
PG_HD uint32_t max3(uint32_t a, uint32_t b, uint32_t c) { return max2(max2(a, b), c); }
PG_HD uint32_t add2(uint32_t a, uint32_t b) { return pk(lo16(a) + lo16(b), hi16(a) + hi16(b)); }
PG_HD uint32_t sub2(uint32_t a, uint32_t b) { return pk(lo16(a) - lo16(b), hi16(a) - hi16(b)); }
PG_HD uint32_t min2(uint32_t a, uint32_t b) { return pk(lo16(a) < lo16(b) ? lo16(a) : lo16(b), hi16(a) < hi16(b) ? hi16(a) : hi16(b)); }
PG_HD uint32_t max2_ge(uint32_t a, uint32_t b, bool& hi_ge, bool& lo_ge)
{
    hi_ge = hi16(a) >= hi16(b);
    lo_ge = lo16(a) >= lo16(b);
    return max2(a, b);
}
#endif

// ---------------------------------------------------------------------------------------------
// sequence coding
// ---------------------------------------------------------------------------------------------
// gssw_create_nt_table (gssw.c:4206-4220): A0 C1 G2 T3, 'U' -> 0 (sic), everything else 4.
PG_HD int nt_code(uint8_t c)
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
// std::toupper in the "C" locale (common/StringUtil.hh:178 via GraphAligner.cpp:218)
PG_HD uint8_t to_upper(uint8_t c) { return (c >= 'a' && c <= 'z') ? (uint8_t)(c - 32) : c; }
// graphtools complementBase: case-sensitive, anything but ACGT -> 'N' (SequenceOperations.cpp:66-81)
PG_HD uint8_t complement_base(uint8_t c)
{
    switch (c)
    {
    case 'A': return 'T';
    case 'C': return 'G';
    case 'G': return 'C';
    case 'T': return 'A';
    default: return 'N';
    }
}
// gssw_create_score_matrix(1, 4) (gssw.c:4188-4204)
PG_HD int sub_score(int a, int b) { return (a == 4 || b == 4) ? 0 : (a == b ? 1 : -4); }

// The four strings GraphAligner::alignRead aligns (GraphAligner.cpp:315-337), by (graph orientation o,
// packed half h), as the character the reference would see at row j after its toUpper():
//   o=0,h=0: bases                      o=0,h=1: reverseComplement(bases)
//   o=1,h=0: reverse(bases)             o=1,h=1: reverseComplement(reverse(bases)) = complement(bases)
PG_HD int read_index(int L, int o, int h, int j) // where row j of that string sits in `bases`
{
    const bool rev = (o == 0) ? (h == 1) : (h == 0);
    return rev ? (L - 1 - j) : j;
}
PG_HD uint8_t read_char_of(uint8_t raw, int h) { return to_upper(h ? complement_base(raw) : raw); }
PG_HD uint8_t read_char(const uint8_t* bases, int L, int o, int h, int j)
{
    return read_char_of(bases[read_index(L, o, h, j)], h);
}

// ---------------------------------------------------------------------------------------------
// graph view (one orientation of one site); all arrays live in device (or emulator) memory
// ---------------------------------------------------------------------------------------------
struct GraphView
{
    const uint8_t* codes;      // codes[-SENT .. G+SENT): column code 0..4, 5 outside [0,G)
    const int32_t* node_start; // [n_nodes] first column of each node (nodes concatenated in id order)
    const int32_t* node_len;   // [n_nodes]
    const int32_t* pred_ptr;   // [n_nodes+1] CSR of predecessor ids, ascending (std::set order, GraphAligner.cpp:147)
    const int32_t* pred_idx;
    int32_t n_nodes;
    int32_t G; // total number of columns
};

// Device-resident description of one site (graph), both orientations (built by pg_host.hpp):
//   bytes blob: per orientation [SENT sentinels][G column codes][SENT sentinels]; forward orientation also the
//               G upper-cased graph characters (for the M/X/N decision, gssw.c:1601-1622);
//   ints blob : per orientation node_start[n], node_len[n], pred_ptr[n+1], pred_idx[n_edges].
// Orientation 1 is graphtools::reverseGraph (GraphOperations.cpp:38-60): node i -> n-1-i, sequences reversed,
// edges flipped -- its column sequence is the forward one read backwards.
struct SiteDev
{
    int32_t n_nodes, G, n_edges;
    int32_t codes_off[2]; // byte offset of column 0 (sentinels precede it)
    int32_t chars_off;    // byte offset of the forward graph characters
    int32_t tab_off[2];   // int offset of the orientation's tables
    int32_t tab_ints;     // ints in one orientation's tables: 3 n_nodes + 1 + distinct edges
    int32_t raw_off;      // byte offset of the RAW (as given) forward graph characters (exact-match stage, pg_path.cuh)
};

// bytes of one orientation's column codes as staged into shared memory: leading sentinels + G + trailing sentinels,
// rounded up to the 16-byte granularity of cp.async.bulk (pg_host.hpp lays the blob out accordingly)
PG_HD uint32_t code_span_bytes(int G) { return ((uint32_t)SENT + (uint32_t)G + SENT + CK + 4 + 15u) & ~15u; }

// view over one orientation's int tables t (in HBM, or a copy staged in shared memory)
PG_HD GraphView make_view_at(const SiteDev& sd, const uint8_t* codes, const int32_t* t)
{
    GraphView g;
    g.codes = codes;
    g.node_start = t;
    g.node_len = t + sd.n_nodes;
    g.pred_ptr = t + 2 * sd.n_nodes;
    g.pred_idx = t + 3 * sd.n_nodes + 1;
    g.n_nodes = sd.n_nodes;
    g.G = sd.G;
    return g;
}

PG_HD GraphView make_view(const SiteDev& sd, const uint8_t* bytes, const int32_t* ints, int o)
{
    GraphView g;
    g.codes = bytes + sd.codes_off[o];
    const int32_t* t = ints + sd.tab_off[o];
    g.node_start = t;
    g.node_len = t + sd.n_nodes;
    g.pred_ptr = t + 2 * sd.n_nodes;
    g.pred_idx = t + 3 * sd.n_nodes + 1;
    g.n_nodes = sd.n_nodes;
    g.G = sd.G;
    return g;
}

// ---------------------------------------------------------------------------------------------
// per-lane wavefront state
// ---------------------------------------------------------------------------------------------
template <int R> struct Lane
{
    uint32_t Hp[R];    // H of the previous column, this lane's rows
    uint32_t E[R];     // E entering the next column
    uint32_t hupPrev;  // H (bottom row of lane t-1) received at the previous step = this step's diagonal
    uint32_t hbotLast; // H of this lane's bottom row at the column just processed (to send down)
    uint32_t foutLast; // F leaving this lane's bottom row at the column just processed (to send down)
};

// words per lane in a checkpoint / in a saved node last column
// Geometry: a task (one column sequence, two packed problems) is processed by a group of W consecutive lanes
// (W = 32, 16 or 8: 1, 2 or 4 tasks per warp), each owning R read rows; W * R >= read length.
template <int R, int W = 32> struct Sizes
{
    static constexpr bool WIDE = is_wide(R, W);
    // checkpoint words per lane: Hp[R], E[R], hupPrev, foutLast -- 4 byte values per word, or (WIDE) the 2R+2 packed registers as they are
    static constexpr int CKW = WIDE ? 2 * R + 2 : R + 1;
    static constexpr int INFOW = WIDE ? 4 : 3; // node maximum, first step per half [, region maximum]
    // The node table: ONE ROW per (node, lane) = that lane's part of the node's last column -- H and the E leaving it,
    // i.e. the seed of the node's successors -- followed by the lane's node maximum and the steps it was first
    // reached at.  Byte-packed like the checkpoints unless WIDE: word r = (H.lo, H.hi, E.lo, E.hi) of row r (H >= 0
    // always, a negative E is as good as 0).  Rows are lane-major and a multiple of four words long, so that the lane
    // that reaches a node boundary stores its row with two 128-bit stores (R = 5: 5 + 3 = 8 words; before: 13 scalar
    // stores into [node][2R + 3][lane] tables that took 1.6 KB of shared memory per node instead of 1 KB -- graphs
    // with many nodes keep more warps resident).
    static constexpr int SEEDV = WIDE ? 2 * R : R;          // value words of a row
    static constexpr int ROWW = (SEEDV + INFOW + 3) & ~3;   // row stride in words
    static constexpr int ROWS = W * R;
    static constexpr int NT = 32 / W;      // tasks per warp
};

template <int R> PG_HD void lane_zero(Lane<R>& s)
{
    for (int r = 0; r < R; ++r)
    {
        s.Hp[r] = 0;
        s.E[r] = 0;
    }
    s.hupPrev = 0;
    s.hbotLast = 0;
    s.foutLast = pk(-1, -1); // no insertion running (any F <= 0 is equivalent; negative lets lazy-F skip)
}

// One wavefront step of one lane.  recvH/recvF = hbotLast/foutLast of lane t-1 after ITS previous step
// (zero / any value <= 0 for lane 0).  pf(r) = packed substitution score of the step's column code against this
// lane's row r.  If KEEP, the values the traceback needs are returned per row:
// Hc = H(i,j), Ec = E(i,j) as used for H (gssw mE), Fc = F(i,j) as used for H (gssw mF; <= 0 when not live).
// Returns max over this lane's rows of t - GAP_OPEN (MBIAS below: the node maxima are tracked with that offset; max t
// == max H since an F-derived H never sets a maximum).
//
// Two passes.  Pass 1: t, E' and the lane maximum for the R rows.  Pass 2: the F chain down the rows, H = max(t, F),
// F' = max(F - ge, t - go).  With LAZY, pass 2 is skipped when it cannot change anything: every t - go of the step
// is negative and the incoming F is negative -- then H = t in every row and the outgoing F is negative too (any
// negative F is as good as another: H >= 0 always wins and F only decays or is replaced by a t - go).  Away from a
// real alignment, i.e. in most columns of the graph, that is the common case.  On the device the decision is taken
// per warp (vote) so that the skip is a uniform branch.
#ifndef PG_LAZY_F
#define PG_LAZY_F 0
#endif
constexpr int MBIAS = GAP_OPEN; // offset of the tracked maxima: Mnode = max(t) - MBIAS, "no score yet" = -MBIAS
template <int W> struct ProfPtr // profile row accessor: plain pointer (host emulator, traceback kernel)
{
    const uint32_t* p;
    PG_HD uint32_t operator()(int r) const { return p[r * W]; }
};
PG_HD bool warp_any(bool x)
{
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, x);
#else
    return x; // the emulator steps lane by lane: a per-lane decision is exact as well
#endif
}

template <int R, bool KEEP, bool LAZY, class PF>
PG_HD uint32_t lane_step_pf(Lane<R>& s, uint32_t recvH, uint32_t recvF, const PF& pf, uint32_t* Hc, uint32_t* Ec,
                            uint32_t* Fc, uint32_t* TgOut = nullptr)
{
    const uint32_t mGO = pk(-GAP_OPEN, -GAP_OPEN), mGE = pk(-GAP_EXT, -GAP_EXT);
    uint32_t d = s.hupPrev; // diagonal for row 0
    s.hupPrev = recvH;
    uint32_t tgLocal[R];
    uint32_t* tg = TgOut ? TgOut : tgLocal; // per-row t - go (the WIDE fill looks at them, track_region)
    uint32_t mg = mGO;
PG_UNROLL
    for (int r = 0; r < R; ++r)
    {
        const uint32_t sc = pf(r);
        const uint32_t e = s.E[r];
        const uint32_t t = addmax_relu2(d, sc, e);
        tg[r] = add2(t, mGO);
        if (KEEP)
            Ec[r] = e;
        s.E[r] = addmax2(e, mGE, tg[r]);
        d = s.Hp[r];
        s.Hp[r] = t;
        mg = max2(mg, tg[r]);
    }
    bool need = true;
    if (LAZY) // both halves of max(t - go, F in) negative in every lane: nothing for the F chain to do
        need = warp_any((~max2(mg, recvF) & 0x80008000u) != 0u);
    if (need)
    {
        uint32_t F = recvF;
PG_UNROLL
        for (int r = 0; r < R; ++r)
        {
            const uint32_t h = max2(s.Hp[r], F);
            if (KEEP)
            {
                Hc[r] = h;
                Fc[r] = F;
            }
            s.Hp[r] = h;
            F = addmax2(F, mGE, tg[r]);
        }
        s.foutLast = F;
    }
    else
        s.foutLast = mGE;
    s.hbotLast = s.Hp[R - 1];
    return mg;
}

// ---- steps with no gap alive ---------------------------------------------------------------------------------
// While no E and no F of a lane group is positive, the recurrence collapses to t = H = max(diag + s, 0): E and F only
// ever enter as max(., E) / max(., F) against values >= 0, a value <= 0 behaves like any other value <= 0 (the
// checkpoints already rely on that, ck_pack), and they stay <= 0 as long as every t of the step is <= GAP_OPEN
// (E' = max(E - ge, t - go), F' likewise).  Away from a real alignment that is the rule: t >= go + 1 takes seven
// matching bases in a row.  So a block of steps may be run SPECULATIVELY with lane_step_dead -- one DPX operation
// per packed cell pair instead of five, no F shuffle -- when gaps_alive() is false for every lane at its start; if
// afterwards some t of the block turned out > GAP_OPEN (its maximum is tracked anyway), the block's start state is
// restored and the block is redone with the full step.  Either way every H and every positive E / F is what the full
// recurrence gives; E / F registers that are <= 0 keep a stale value <= 0.
template <int R> PG_HD bool gaps_alive(const Lane<R>& s)
{
    uint32_t m = s.foutLast; // the F this lane hands to the next one
PG_UNROLL
    for (int r = 0; r < R; ++r)
        m = max2(m, s.E[r]);
    return max2(m, 0u) != 0u; // some half > 0
}
template <int R> PG_HD bool e_alive(const Lane<R>& s) // the E half of gaps_alive
{
    uint32_t m = 0u;
PG_UNROLL
    for (int r = 0; r < R; ++r)
        m = max2(m, s.E[r]);
    return m != 0u;
}
// one step of one lane under that premise; returns the maximum of t over the lane's rows (NOT offset by MBIAS).
// zero = 0, on the device in a register the compiler cannot see through (a literal 0 as the third DPX operand makes
// it materialise a zero register per operation).
template <int R, class PF> PG_HD uint32_t lane_step_dead(Lane<R>& s, uint32_t recvH, const PF& pf, uint32_t zero = 0u, uint32_t acc = 0u)
{
    uint32_t d = s.hupPrev;
    s.hupPrev = recvH;
    uint32_t mt = acc; // a caller that only needs the maximum over a whole block passes the running one in (acc >= 0)
PG_UNROLL
    for (int r = 0; r < R; ++r)
    {
        const uint32_t t = addmax_relu2(d, pf(r), zero);
        d = s.Hp[r];
        s.Hp[r] = t;
        mt = max2(mt, t);
    }
    s.hbotLast = s.Hp[R - 1];
    return mt;
}
// A speculative block need not track WHERE a lane's node maximum was first reached: everything it computes is <= GAP_OPEN
// (else it is redone), and the first-reached steps only matter for the cells that hold the fill's top score.  So a
// block just folds its maximum into the lane's node maximum; a fill whose top score ends up within 1 .. GAP_OPEN (a read
// that matches nowhere) is done once more with exact bookkeeping (dead_range_score; fill kernel and emulator alike).
PG_HD bool dead_range_score(int S) { return S > 0 && S <= GAP_OPEN; }
// did a speculative block stay within the premise?  Mt = maximum of lane_step_dead's results over the block
PG_HD bool dead_block_broken(uint32_t Mt) { return max2(add2(Mt, pk(-GAP_OPEN, -GAP_OPEN)), 0u) != 0u; }
// what a speculative block changes and a redo has to put back (E, foutLast are not touched by lane_step_dead)
template <int R> struct DeadSave
{
    uint32_t Hp[R], hupPrev, hbotLast;
};
template <int R> PG_HD void dead_save(const Lane<R>& s, DeadSave<R>& b)
{
PG_UNROLL
    for (int r = 0; r < R; ++r)
        b.Hp[r] = s.Hp[r];
    b.hupPrev = s.hupPrev;
    b.hbotLast = s.hbotLast;
}
template <int R> PG_HD void dead_restore(Lane<R>& s, const DeadSave<R>& b)
{
PG_UNROLL
    for (int r = 0; r < R; ++r)
        s.Hp[r] = b.Hp[r];
    s.hupPrev = b.hupPrev;
    s.hbotLast = b.hbotLast;
}
// ---- upper-bound pruning of gaps (experiment) ---------------------------------------------
// A gap value v reaching row j cannot end above v + rem, rem = L - 1 - j (one match per remaining read row).  With
// Sb = a lower bound of the fill's final best score (the best seen so far), a gap with v + rem < Sb can change
// neither score nor ties nor end cell nor traceback, and may be dropped.  All arguments are packed halves; rem may
// be negative (padding rows: never relevant).
// Some half: v > 0, rem >= 0 and v + rem >= Sb?
PG_HD bool gap_relevant(uint32_t v, uint32_t rem, uint32_t Sb)
{
    const uint32_t neg = (sub2(add2(v, rem), Sb) | add2(v, pk(-1, -1)) | rem) & 0x80008000u; // per half: sign of any of the three
    return neg != 0x80008000u;
}
// rem0 = rem of the lane's first row, remNext = rem of the next lane's first row (where foutLast arrives)
template <int R> PG_HD bool lane_gaps_relevant(const Lane<R>& s, uint32_t rem0, uint32_t remNext, uint32_t Sb)
{
    bool rel = gap_relevant(s.foutLast, remNext, Sb);
PG_UNROLL
    for (int r = 0; r < R; ++r)
        rel = rel || gap_relevant(s.E[r], sub2(rem0, pk(r, r)), Sb);
    return rel;
}
template <int R> PG_HD void lane_gaps_drop(Lane<R>& s)
{
PG_UNROLL
    for (int r = 0; r < R; ++r)
        s.E[r] = min2(s.E[r], 0u);
    s.foutLast = min2(s.foutLast, 0u);
}
// a dead block whose maximum Mt broke the premise: were the gaps it would have opened (at most Mt - go, in the
// lane's first row at best) all droppable?
PG_HD bool dead_block_broken_pruned(uint32_t Mt, uint32_t rem0, uint32_t Sb)
{
    return gap_relevant(add2(Mt, pk(-GAP_OPEN, -GAP_OPEN)), rem0, Sb);
}

#ifndef PG_SPEC_STEPS
#define PG_SPEC_STEPS 8
#endif
static_assert(CK % PG_SPEC_STEPS == 0, "PG_SPEC_STEPS must divide the checkpoint interval");
constexpr int SPEC_STEPS = PG_SPEC_STEPS; // steps per speculative block; CK must be a multiple (4 / 8 / 16: shorter blocks
                                          // are redone less often but pay the vote and the state copy more often)

// prof = this group's profile, word (c*R + r)*W + lane = packed score of column code c against this lane's row r.
template <int R, bool KEEP, int W = 32>
PG_HD uint32_t lane_step(Lane<R>& s, uint32_t recvH, uint32_t recvF, const uint32_t* prof, int code, int lane,
                         uint32_t* Hc, uint32_t* Ec, uint32_t* Fc)
{
    const ProfPtr<W> pf = { prof + (code * R) * W + lane };
    return lane_step_pf<R, KEEP, (PG_LAZY_F != 0) && !KEEP>(s, recvH, recvF, pf, Hc, Ec, Fc);
}

// Build this lane's part of the warp profile: rows [R*lane, R*lane+R) x 6 column codes, both halves.
// Rows >= L and the sentinel code get NEG so that they can never reach a maximum (DESIGN.md "padding").
template <int R, int W = 32>
PG_HD void build_profile_pair(uint32_t* prof, const uint8_t* bases0, int L0, int o0, int h0, const uint8_t* bases1, int L1,
                              int o1, int h1, int lane);
template <int R, int W = 32> PG_HD void build_profile(uint32_t* prof, const uint8_t* bases, int L, int orient, int lane)
{
    build_profile_pair<R, W>(prof, bases, L, orient, 0, bases, L, orient, 1, lane);
}

// The same with the two packed halves taken from two different reads of one site: half x is the string
// (orientation ox, half hx) of read x, or absent (bases_x == nullptr: all NEG).  Used by the paired reversed-graph
// tasks, where each half is the one reversed-graph fill some read still needs (rev_plan).
// All characters are fetched first and decoded afterwards: the 2R byte loads of a lane are independent and go out
// back to back (one memory latency per task instead of 2R -- decoding between them had them serialised, 12 % of the
// fill kernel's stall samples sat here).
template <int R, int W>
PG_HD void build_profile_pair(uint32_t* prof, const uint8_t* bases0, int L0, int o0, int h0, const uint8_t* bases1, int L1,
                              int o1, int h1, int lane)
{
    uint8_t raw0[R], raw1[R];
    bool has0[R], has1[R];
PG_UNROLL
    for (int r = 0; r < R; ++r)
    {
        const int j = R * lane + r;
        has0[r] = bases0 && j < L0;
        has1[r] = bases1 && j < L1;
        raw0[r] = has0[r] ? bases0[read_index(L0, o0, h0, j)] : (uint8_t)0;
        raw1[r] = has1[r] ? bases1[read_index(L1, o1, h1, j)] : (uint8_t)0;
    }
PG_UNROLL
    for (int r = 0; r < R; ++r)
    {
        const int c0 = has0[r] ? nt_code(read_char_of(raw0[r], h0)) : -1;
        const int c1 = has1[r] ? nt_code(read_char_of(raw1[r], h1)) : -1;
        for (int c = 0; c < NCODE; ++c)
        {
            const int s0 = (c0 < 0 || c == 5) ? NEG : sub_score(c, c0);
            const int s1 = (c1 < 0 || c == 5) ? NEG : sub_score(c, c1);
            prof[(c * R + r) * W + lane] = pk(s0, s1);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// per-lane bookkeeping around the step: which node the lane is in, node maxima, node boundaries
// ---------------------------------------------------------------------------------------------
constexpr int COLS_INF = 0x3fffffff;

struct LaneCtl
{
    int node;      // node the lane is currently in (n_nodes once past the end)
    int colsLeft;  // columns of that node still to process, counted at the top of a step (0 -> node just ended)
    uint32_t Mnode; // packed maximum of t - MBIAS over this lane's rows within the current node
    int first[2];  // wavefront step at which Mnode's half first reached its current value
    // WIDE geometries only: the same maximum restricted to the cells GraphAligner's uniqueness scan sees once gssw
    // is in 16-bit mode (finalize_task): the node's first ceil(len * L / 2) cells in (column, row) order.
    uint32_t Mreg;
    int regLeft;   // columns still fully inside that region (0: the next column is the partial one, < 0: past it)
};

// number of leading columns of a node of length len that lie fully inside the scanned region, and the number of
// rows of the following column that do (GraphAligner.cpp:177-186 reads len * L BYTES of the int16 matrix)
PG_HD void scan_region(int len, int L, int& full_cols, int& part_rows)
{
    const int cells = (len * L + 1) >> 1;
    full_cols = cells / L;
    part_rows = cells - full_cols * L;
}

// Control state of `lane` at the top of step k0, before the node event of that step has run.
PG_HD void ctl_at_step(LaneCtl& c, const GraphView& g, int k0, int lane)
{
    const int q = k0 - lane; // column about to be processed
    c.Mnode = pk(-MBIAS, -MBIAS);
    c.first[0] = c.first[1] = 0;
    c.Mreg = pk(-MBIAS, -MBIAS);
    c.regLeft = -1; // set by region_begin() in the fill of a WIDE geometry
    if (q <= 0)
    {
        c.node = 0;
        c.colsLeft = g.node_len[0] - q;
        return;
    }
    if (q >= g.G)
    {
        c.node = (q == g.G) ? g.n_nodes - 1 : g.n_nodes;
        c.colsLeft = (q == g.G) ? 0 : COLS_INF;
        return;
    }
    int n = 0;
    while (q >= g.node_start[n] + g.node_len[n])
        ++n;
    if (q == g.node_start[n]) // first column of node n > 0: node n-1 has just ended, its event is pending
    {
        c.node = n - 1;
        c.colsLeft = 0;
    }
    else
    {
        c.node = n;
        c.colsLeft = g.node_start[n] + g.node_len[n] - q;
    }
}

PG_HD void track_max(LaneCtl& c, uint32_t m, int k)
{
    bool hi_ge, lo_ge; // old maximum >= this step's value: nothing new
    c.Mnode = max2_ge(c.Mnode, m, hi_ge, lo_ge);
    if (!lo_ge)
        c.first[0] = k;
    if (!hi_ge)
        c.first[1] = k;
}

// The same for a run of steps whose maxima come in the t domain (lane_step_dead: max t, not max t - MBIAS): the node
// maximum is moved to that domain once (track_t_begin), compared and updated there, and moved back at the end
// (track_t_end) -- m - MBIAS > Mnode <=> m > Mnode + MBIAS -- which saves the subtraction per step.
PG_HD uint32_t track_t_begin(const LaneCtl& c) { return add2(c.Mnode, pk(MBIAS, MBIAS)); }
PG_HD void track_t(LaneCtl& c, uint32_t& Mn, uint32_t mt, int k)
{
    bool hi_ge, lo_ge;
    Mn = max2_ge(Mn, mt, hi_ge, lo_ge);
    if (!lo_ge)
        c.first[0] = k;
    if (!hi_ge)
        c.first[1] = k;
}
PG_HD void track_t_end(LaneCtl& c, uint32_t Mn) { c.Mnode = add2(Mn, pk(-MBIAS, -MBIAS)); }

// WIDE fill, once per step after lane_step: fold the step into the region-restricted maximum.  tg = the step's
// per-row t - go (lane_step's optional output), mg their maximum.
template <int R> PG_HD void track_region(LaneCtl& c, uint32_t mg, const uint32_t* tg, const GraphView& g, int L, int lane)
{
    if (c.regLeft > 0)
    {
        c.Mreg = max2(c.Mreg, mg);
        --c.regLeft;
    }
    else if (c.regLeft == 0) // the column the region ends in: only its first part_rows rows count
    {
        int full, part;
        scan_region(g.node_len[c.node], L, full, part);
        for (int r = 0; r < R; ++r)
            if (R * lane + r < part)
                c.Mreg = max2(c.Mreg, tg[r]);
        c.regLeft = -1;
    }
}
// start of the fill (step 0): lane t first runs through t sentinel columns, which count for nothing
PG_HD void region_begin(LaneCtl& c, const GraphView& g, int L, int lane)
{
    int full, part;
    scan_region(g.node_len[0], L, full, part);
    c.regLeft = full + lane;
    c.Mreg = pk(-MBIAS, -MBIAS);
}

// row of (node n, lane) in a node table
template <int R, int W> PG_HD uint32_t* node_row(uint32_t* tab, int n, int lane) { return tab + ((size_t)n * W + lane) * Sizes<R, W>::ROWW; }
template <int R, int W> PG_HD const uint32_t* node_row(const uint32_t* tab, int n, int lane)
{
    return tab + ((size_t)n * W + lane) * Sizes<R, W>::ROWW;
}
PG_HD uint32_t ck_pack(uint32_t v0, uint32_t v1);
PG_HD void ck_unpack(uint32_t w, uint32_t& v0, uint32_t& v1);
// ck_pack of (H, E) when H is known to be >= 0 in both halves (H = max(t, F) with t >= 0): only E needs the clamp
PG_HD uint32_t pack_he(uint32_t h, uint32_t e)
{
    e = max2(e, 0u);
#if defined(__CUDA_ARCH__)
    return __byte_perm(h, e, 0x6420);
#else
    return (h & 0xffu) | (((h >> 16) & 0xffu) << 8) | ((e & 0xffu) << 16) | (((e >> 16) & 0xffu) << 24);
#endif
}

// fold the last column saved in `row` into (H, E): element-wise maximum
template <int R, int W> PG_HD void seed_max(const uint32_t* row, uint32_t* H, uint32_t* E)
{
    constexpr int V = Sizes<R, W>::SEEDV;
    uint32_t v[V];
#if defined(__CUDA_ARCH__)
    static_assert(Sizes<R, W>::ROWW % 4 == 0, "rows are read with 128-bit loads");
PG_UNROLL
    for (int x = 0; x < (V + 3) / 4; ++x) // (the words after the values belong to the same row: reading them is harmless)
    {
        const uint4 q = reinterpret_cast<const uint4*>(row)[x];
        if (4 * x + 0 < V) v[4 * x + 0] = q.x;
        if (4 * x + 1 < V) v[4 * x + 1] = q.y;
        if (4 * x + 2 < V) v[4 * x + 2] = q.z;
        if (4 * x + 3 < V) v[4 * x + 3] = q.w;
    }
#else
    for (int x = 0; x < V; ++x)
        v[x] = row[x];
#endif
PG_UNROLL
    for (int r = 0; r < R; ++r)
    {
        uint32_t h, e;
        if (Sizes<R, W>::WIDE)
        {
            h = v[r];
            e = v[R + r];
        }
        else
            ck_unpack(v[r], h, e);
        H[r] = max2(H[r], h);
        E[r] = max2(E[r], e);
    }
}
// H of the bottom row of `row`'s lane (the diagonal into the first row of the lane below)
template <int R, int W> PG_HD uint32_t seed_bottom_h(const uint32_t* row)
{
    if (Sizes<R, W>::WIDE)
        return row[R - 1];
    uint32_t h, e;
    ck_unpack(row[R - 1], h, e);
    return h;
}

// Node boundary handling at the top of a step (rare, per lane: lanes reach a boundary at different steps).
//   FILL = true  (fill kernel): save the finished node's row (last column H, E leaving it, the lane's node maximum /
//                first steps) into the warp-private node table in shared memory; then load the next node's seed.
//                (The fill kernel copies the table to HBM once, at the end of a forward-graph task, for the traceback.)
//   FILL = false (tile recomputation in the traceback kernel): only load the next node's seed, from that copy.
// Seed of a node = element-wise max over its predecessors' last columns (gssw_create_seed_byte,
// gssw.c:3897-3931), zeros for a source; if the only predecessor is the node just finished the state
// simply carries over, and when it is one of several the registers stand for it (they ARE its last column).  The
// diagonal into this lane's first row comes from the seed row just above it, i.e. the bottom row of lane-1's row of
// each predecessor (written by lane-1 at least one step earlier).
template <int R, bool FILL, int W = 32>
PG_HD_COLD void node_event(Lane<R>& s, LaneCtl& c, const GraphView& g, int lane, uint32_t* tab, int L = 0)
{
    constexpr int V = Sizes<R, W>::SEEDV, RW = Sizes<R, W>::ROWW;
    if (c.colsLeft == 0)
    {
        const int n = c.node;
        if (FILL)
        {
            uint32_t v[RW];
PG_UNROLL
            for (int x = 0; x < RW; ++x)
                v[x] = 0u;
PG_UNROLL
            for (int r = 0; r < R; ++r)
            {
                if (Sizes<R, W>::WIDE)
                {
                    v[r] = s.Hp[r];
                    v[R + r] = s.E[r];
                }
                else
                    v[r] = pack_he(s.Hp[r], s.E[r]);
            }
            v[V + 0] = c.Mnode;
            v[V + 1] = (uint32_t)c.first[0];
            v[V + 2] = (uint32_t)c.first[1];
            c.Mnode = pk(-MBIAS, -MBIAS);
            if (Sizes<R, W>::WIDE)
            {
                v[V + 3] = c.Mreg;
                c.Mreg = pk(-MBIAS, -MBIAS);
                c.regLeft = -1;
                if (n + 1 < g.n_nodes)
                {
                    int part;
                    scan_region(g.node_len[n + 1], L, c.regLeft, part);
                }
            }
            uint32_t* row = node_row<R, W>(tab, n, lane);
#if defined(__CUDA_ARCH__)
PG_UNROLL
            for (int x = 0; x < RW / 4; ++x)
                reinterpret_cast<uint4*>(row)[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
#else
            for (int x = 0; x < RW; ++x)
                row[x] = v[x];
#endif
        }
        c.node = n + 1;
        if (n + 1 < g.n_nodes)
        {
            c.colsLeft = g.node_len[n + 1];
            const int p0 = g.pred_ptr[n + 1], p1 = g.pred_ptr[n + 2];
            if (!(p1 - p0 == 1 && g.pred_idx[p0] == n))
            {
                uint32_t H[R], E[R], hup = 0;
                for (int r = 0; r < R; ++r)
                    H[r] = E[r] = 0;
PG_NOUNROLL
                for (int e = p0; e < p1; ++e)
                {
                    const int p = g.pred_idx[e];
                    if (p == n) // the node just finished: its last column is what the registers hold
                    {
PG_UNROLL
                        for (int r = 0; r < R; ++r)
                        {
                            H[r] = max2(H[r], s.Hp[r]);
                            E[r] = max2(E[r], s.E[r]);
                        }
                        hup = max2(hup, s.hupPrev);
                        continue;
                    }
                    const uint32_t* src = node_row<R, W>(tab, p, lane);
                    seed_max<R, W>(src, H, E);
                    if (lane > 0)
                        hup = max2(hup, seed_bottom_h<R, W>(src - RW));
                }
                for (int r = 0; r < R; ++r)
                {
                    s.Hp[r] = H[r];
                    s.E[r] = E[r];
                }
                s.hupPrev = hup;
            }
        }
        else
            c.colsLeft = COLS_INF;
    }
    --c.colsLeft;
}

// ---- node boundaries inside the fill's boundary sub-blocks: entry words + prefetched seeds -----------------------
// A boundary reaches the lanes of a group on consecutive steps, so node_event runs with ONE active lane and costs the
// warp its whole instruction count every time: 60 (chain) to 130 (merge of two predecessors) issue slots per lane and
// boundary -- 11 % of the fill's instructions on 3-node graphs, nearly half on 6-9-node vcf2paragraph-shaped ones
// (profiles/r02d_fill_fwd_source.csv.gz: the rows executed with one thread).  Two things make it lean:
//   * entry_word(m): what a lane needs to know when it enters node m from node m - 1, precomputed once per task into
//     shared memory -- the node's length, whether m - 1 is among its predecessors (EV_PREV) and whether anything has
//     to be merged at all (EV_MERGE; clear for a chain link, where the state just carries over);
//   * seed_prefetch: at the start of a boundary sub-block, ALL lanes that will enter a merging node within it fold
//     the last columns of that node's OLDER predecessors (ids < m - 1) at once -- rows they wrote themselves at
//     earlier boundaries, so they are complete -- which leaves the lane-at-a-time event with the row store and
//     a register-only maximum (node_event_pre).  A lane that crosses a second boundary within the same sub-block
//     (nodes shorter than it) takes the general path for that one.
constexpr uint32_t EV_LEN = 0x3fffffffu, EV_PREV = 0x40000000u, EV_MERGE = 0x80000000u;
PG_HD uint32_t entry_word(const GraphView& g, int m) // m >= 1
{
    const int p0 = g.pred_ptr[m], p1 = g.pred_ptr[m + 1];
    const bool prev = p1 > p0 && g.pred_idx[p1 - 1] == m - 1; // ids ascending and < m: m - 1 can only be the last one
    uint32_t w = (uint32_t)g.node_len[m] & EV_LEN;
    if (prev)
        w |= EV_PREV;
    if (!(prev && p1 - p0 == 1))
        w |= EV_MERGE;
    return w;
}
template <int R> struct SeedPre
{
    uint32_t H[R], E[R], hup; // element-wise maximum over the older predecessors' last columns (>= 0)
    int node;                 // the node they are the seed part of; -1: nothing prefetched
    bool twice;               // the lane crosses a second boundary within the horizon (nodes shorter than it)
};
// does the prefetched seed part bring a live gap (an E > 0) into the node?
template <int R> PG_HD bool seed_live(const SeedPre<R>& p)
{
    if (p.node < 0)
        return false;
    uint32_t m = 0u;
PG_UNROLL
    for (int r = 0; r < R; ++r)
        m = max2(m, p.E[r]);
    return m != 0u; // (all E here are >= 0)
}
template <int R, int W>
PG_HD void seed_prefetch(SeedPre<R>& p, const LaneCtl& c, const GraphView& g, const uint32_t* entry, int lane,
                         const uint32_t* tab, int horizon)
{
    constexpr int RW = Sizes<R, W>::ROWW;
    p.node = -1;
    p.twice = false;
    const int m = c.node + 1;
    if (c.colsLeft >= horizon || m >= g.n_nodes)
        return;
    const uint32_t w = entry[m];
    p.twice = c.colsLeft + (int)(w & EV_LEN) < horizon;
    if (!(w & EV_MERGE))
        return;
    p.node = m;
PG_UNROLL
    for (int r = 0; r < R; ++r)
        p.H[r] = p.E[r] = 0;
    p.hup = 0;
    const int p0 = g.pred_ptr[m], p1 = g.pred_ptr[m + 1] - ((w & EV_PREV) ? 1 : 0);
PG_NOUNROLL
    for (int e = p0; e < p1; ++e)
    {
        const uint32_t* src = node_row<R, W>(tab, g.pred_idx[e], lane);
        seed_max<R, W>(src, p.H, p.E);
        if (lane > 0)
            p.hup = max2(p.hup, seed_bottom_h<R, W>(src - RW));
    }
}
// The general statement of a merge (what node_event does): every predecessor of node n + 1 from the table, the registers
// standing for the node just finished.  Rare here (see above).
template <int R, int W>
PG_COLD_CALL void seed_general(Lane<R>& s, const GraphView& g, int n, int lane, const uint32_t* tab)
{
    constexpr int RW = Sizes<R, W>::ROWW;
    uint32_t H[R], E[R], hup = 0;
    for (int r = 0; r < R; ++r)
        H[r] = E[r] = 0;
    const int p0 = g.pred_ptr[n + 1], p1 = g.pred_ptr[n + 2];
PG_NOUNROLL
    for (int e = p0; e < p1; ++e)
    {
        const int p = g.pred_idx[e];
        if (p == n)
        {
PG_UNROLL
            for (int r = 0; r < R; ++r)
            {
                H[r] = max2(H[r], s.Hp[r]);
                E[r] = max2(E[r], s.E[r]);
            }
            hup = max2(hup, s.hupPrev);
            continue;
        }
        const uint32_t* src = node_row<R, W>(tab, p, lane);
        seed_max<R, W>(src, H, E);
        if (lane > 0)
            hup = max2(hup, seed_bottom_h<R, W>(src - RW));
    }
    for (int r = 0; r < R; ++r)
    {
        s.Hp[r] = H[r];
        s.E[r] = E[r];
    }
    s.hupPrev = hup;
}
// node_event<R, true, W> of a non-WIDE geometry with the entry words and (where it applies) the prefetched seed part
template <int R, int W>
PG_HD_COLD void node_event_pre(Lane<R>& s, LaneCtl& c, const GraphView& g, const uint32_t* entry, const SeedPre<R>& pre,
                               int lane, uint32_t* tab)
{
    // (byte-packed rows: non-WIDE geometries only)
    constexpr int V = Sizes<R, W>::SEEDV, RW = Sizes<R, W>::ROWW;
    const int n = c.node;
    {
        uint32_t v[RW];
PG_UNROLL
        for (int x = 0; x < RW; ++x)
            v[x] = 0u;
PG_UNROLL
        for (int r = 0; r < R; ++r)
            v[r] = pack_he(s.Hp[r], s.E[r]);
        v[V + 0] = c.Mnode;
        v[V + 1] = (uint32_t)c.first[0];
        v[V + 2] = (uint32_t)c.first[1];
        c.Mnode = pk(-MBIAS, -MBIAS);
        uint32_t* row = node_row<R, W>(tab, n, lane);
#if defined(__CUDA_ARCH__)
PG_UNROLL
        for (int x = 0; x < RW / 4; ++x)
            reinterpret_cast<uint4*>(row)[x] = make_uint4(v[4 * x], v[4 * x + 1], v[4 * x + 2], v[4 * x + 3]);
#else
        for (int x = 0; x < RW; ++x)
            row[x] = v[x];
#endif
    }
    c.node = n + 1;
    if (n + 1 < g.n_nodes)
    {
        const uint32_t w = entry[n + 1];
        c.colsLeft = (int)(w & EV_LEN) - 1;
        if (w & EV_MERGE)
        {
            if (pre.node == n + 1)
            {
                if (w & EV_PREV) // the node just finished is a predecessor too: the registers are its last column
                {
PG_UNROLL
                    for (int r = 0; r < R; ++r)
                    {
                        s.Hp[r] = max2(s.Hp[r], pre.H[r]);
                        s.E[r] = max2(s.E[r], pre.E[r]);
                    }
                    s.hupPrev = max2(s.hupPrev, pre.hup);
                }
                else
                {
PG_UNROLL
                    for (int r = 0; r < R; ++r)
                    {
                        s.Hp[r] = pre.H[r];
                        s.E[r] = pre.E[r];
                    }
                    s.hupPrev = pre.hup;
                }
            }
            else // second boundary of this lane within one sub-block: the general statement
                seed_general<R, W>(s, g, n, lane, tab);
        }
    }
    else
        c.colsLeft = COLS_INF - 1;
}

// Checkpoint = lane state at the top of a step, byte-packed: every value is two halves in [0, 255] once clamped at
// 0 (scores <= MAX_READ_LEN; negative E / F are equivalent to 0), so two packed registers fit
// one 32-bit word: (lo0, hi0, lo1, hi1).  R + 1 words per lane instead of 2R + 2.
PG_HD uint32_t ck_pack(uint32_t v0, uint32_t v1)
{
    v0 = max2(v0, 0u);
    v1 = max2(v1, 0u);
#if defined(__CUDA_ARCH__)
    return __byte_perm(v0, v1, 0x6420);
#else
    return (v0 & 0xffu) | (((v0 >> 16) & 0xffu) << 8) | ((v1 & 0xffu) << 16) | (((v1 >> 16) & 0xffu) << 24);
#endif
}
PG_HD void ck_unpack(uint32_t w, uint32_t& v0, uint32_t& v1)
{
#if defined(__CUDA_ARCH__)
    v0 = __byte_perm(w, 0u, 0x4140);
    v1 = __byte_perm(w, 0u, 0x4342);
#else
    v0 = (w & 0xffu) | (((w >> 8) & 0xffu) << 16);
    v1 = ((w >> 16) & 0xffu) | (((w >> 24) & 0xffu) << 16);
#endif
}
template <int R, int W = 32> PG_HD void ckpt_store(const Lane<R>& s, uint32_t* ck, int lane)
{
    uint32_t v[2 * R + 2];
    for (int r = 0; r < R; ++r)
    {
        v[r] = s.Hp[r];
        v[R + r] = s.E[r];
    }
    v[2 * R] = s.hupPrev;
    v[2 * R + 1] = s.foutLast;
    if (Sizes<R, W>::WIDE) // scores may exceed a byte: keep the packed int16 pairs
        for (int x = 0; x < 2 * R + 2; ++x)
            ck[x * W + lane] = v[x];
    else
        for (int x = 0; x < R + 1; ++x)
            ck[x * W + lane] = ck_pack(v[2 * x], v[2 * x + 1]);
}
template <int R, int W = 32> PG_HD void ckpt_load(Lane<R>& s, const uint32_t* ck, int lane)
{
    uint32_t v[2 * R + 2];
    if (Sizes<R, W>::WIDE)
        for (int x = 0; x < 2 * R + 2; ++x)
            v[x] = ck[x * W + lane];
    else
        for (int x = 0; x < R + 1; ++x)
            ck_unpack(ck[x * W + lane], v[2 * x], v[2 * x + 1]);
    for (int r = 0; r < R; ++r)
    {
        s.Hp[r] = v[r];
        s.E[r] = v[R + r];
    }
    s.hupPrev = v[2 * R];
    s.foutLast = v[2 * R + 1];
    s.hbotLast = s.Hp[R - 1];
}

// Traceback tiles keep, per cell, one word H | E << 8 | F << 16 (chosen half; E/F clamped at 0 like gssw's unsigned
// saturation; all three fit a byte because scores are <= MAX_READ_LEN) and only for a band of BAND_LANES lanes
// (BAND_LANES * R read rows): within the CK steps of a tile the walk moves through ~CK rows, so rows outside
// the band are never read.  A cell outside the resident bands is simply a miss (tile recomputed around it).
template <int R> struct TileGeom
{
    // rows the walk can climb within one tile (<= 1 per step on a diagonal, + insertions) + lane alignment slack
    static constexpr int BAND_LANES = (TS + 19 + R - 1) / R + 1;
    static constexpr int BAND_ROWS = BAND_LANES * R;
    static constexpr int SLOT_WORDS = TS * BAND_ROWS;
};

// PRMT: result byte i = byte (selector nibble i) of the eight bytes {a: 0-3, b: 4-7} (selector nibbles < 8 only)
PG_HD uint32_t bperm(uint32_t a, uint32_t b, uint32_t sel)
{
#if defined(__CUDA_ARCH__)
    return __byte_perm(a, b, sel);
#else
    const uint64_t v = (uint64_t)a | ((uint64_t)b << 32);
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i)
        r |= (uint32_t)((v >> (8 * ((sel >> (4 * i)) & 7u))) & 0xffu) << (8 * i);
    return r;
#endif
}
template <bool WIDE> PG_HD uint32_t pack_cell(uint32_t h, uint32_t e, uint32_t f, int half)
{
    if (WIDE) // H in 11 bits (scores <= MAX_READ_LEN = 1024), E and F in 10 bits each (a gap value is <= score - GAP_OPEN)
    {
        e = max2(e, 0u);
        f = max2(f, 0u);
        return (uint32_t)half16(h, half) | ((uint32_t)half16(e, half) << 11) | ((uint32_t)half16(f, half) << 21);
    }
    // three operations: the chosen halves of E and F side by side as one int16 pair, one clamp for both, then
    // bytes b0 = H.byte(2*half), b1 = E.byte0, b2 = F.byte0, b3 = F.byte1 (= 0: values <= 255 after the clamp)
    const uint32_t ef = max2(bperm(e, f, half ? 0x7632u : 0x5410u), 0u);
    return bperm(h, ef, half ? 0x7642u : 0x7640u);
}
template <bool WIDE> PG_HD int cellH(uint32_t w) { return (int)(WIDE ? (w & 0x7ffu) : (w & 0xffu)); }
template <bool WIDE> PG_HD int cellE(uint32_t w) { return (int)(WIDE ? ((w >> 11) & 0x3ffu) : ((w >> 8) & 0xffu)); }
template <bool WIDE> PG_HD int cellF(uint32_t w) { return (int)(WIDE ? ((w >> 21) & 0x3ffu) : ((w >> 16) & 0xffu)); }
static_assert(MAX_READ_LEN <= 1024 && MAX_READ_LEN - GAP_OPEN < 1024, "WIDE tile cells: 11 + 10 + 10 bits");

// one step of a traceback tile: lanes [blo, blo + BAND_LANES) store their R cells
template <int R, bool WIDE = false>
PG_HD void tile_store(uint32_t* tstep, int lane, int blo, const uint32_t* Hc, const uint32_t* Ec, const uint32_t* Fc,
                      int half)
{
    const int bl = lane - blo;
    if (bl < 0 || bl >= TileGeom<R>::BAND_LANES)
        return;
    for (int r = 0; r < R; ++r)
        tstep[R * bl + r] = pack_cell<WIDE>(Hc[r], Ec[r], Fc[r], half);
}

// ---------------------------------------------------------------------------------------------
// per-task scratch layout (32-bit words, [..][32 lanes] innermost so that a warp store is one 128 B line)
// ---------------------------------------------------------------------------------------------
//   node table [n_nodes][32][ROWW] : per node and lane one row (Sizes): the node's last column (H, E leaving it = seed),
//                                  byte-packed, then the packed node maximum and the first step reaching it per half;
//                                  copied to HBM as `last` by forward-graph tasks
//   ckpt  [n_ck][R+1][32]       : byte-packed lane state before step c*CK           (forward-graph tasks only)
PG_HD int num_steps(int G, int W) { return G + W; } // lane W-1 ends column G-1 at step G+W-2; its node event runs at step G+W-1
PG_HD int num_ckpt(int G, int W) { return (num_steps(G, W) + CK - 1) / CK; }

struct TaskOut // result of one fill (two packed problems)
{
    int32_t score[2];    // global maximum S
    int32_t n_top[2];    // number of nodes whose real cells contain S (capped at 2)
    int32_t max_node[2]; // first node containing S, -1 when S == 0
    int32_t end_step[2]; // wavefront step at which (end_lane) first held S in max_node
    int32_t end_lane[2];
};

// Serial reduction of the per-(node, lane) maxima written by the fill (run by one lane at the end of a task).
// Best cell per gssw: first node in array order whose maximum is the global one, first column in it, smallest
// row in that column (gssw.c:378-386, 446-454, 4015-4018) -> min column = min(step - lane), ties -> smaller lane.
PG_HD uint32_t ld_scratch(const uint32_t* p) { return *p; } // warp-private shared memory (after a __syncwarp)

// Uniqueness (GraphAligner.cpp:170-212): n_top = number of nodes "containing the top score" as the reference's scan
// sees them.  In 8-bit mode (S <= BYTE_MAX_SCORE) that is the plain count.  Once S >= 251 gssw has redone the graph
// in 16-bit mode and the scan, which walks len*L BYTES of each node's matrix through a uint8_t*, (a) can never match
// a top score >= 256 -> no node, "unique"; (b) for 251..255 matches the low byte of the cells in the first half of
// the node's matrix only -> the count over the region-restricted maxima (LaneCtl::Mreg, info word 3).
PG_HD int n_top_rule(int S, int plain, int region) { return S <= BYTE_MAX_SCORE ? plain : (S >= 256 ? 1 : region); }

template <int R, int W> PG_HD void finalize_task(const uint32_t* tab, int n_nodes, TaskOut& o)
{
    constexpr int IW = Sizes<R, W>::INFOW, V = Sizes<R, W>::SEEDV;
    auto info = [&](int n, int x, int t) { return ld_scratch(node_row<R, W>(tab, n, t) + V + x); };
    for (int h = 0; h < 2; ++h)
    {
        int S = 0;
        for (int n = 0; n < n_nodes; ++n)
            for (int t = 0; t < W; ++t)
            {
                const int v = half16(info(n, 0, t), h) + MBIAS;
                if (v > S)
                    S = v;
            }
        int ntop = 0, nreg = 0, mnode = -1, bestq = 0x7fffffff, blane = 0, bstep = 0;
        for (int n = 0; n < n_nodes; ++n)
        {
            bool has = false, hasreg = false;
            for (int t = 0; t < W; ++t)
            {
                if (IW > 3 && half16(info(n, 3, t), h) + MBIAS == S)
                    hasreg = true;
                if (half16(info(n, 0, t), h) + MBIAS != S)
                    continue;
                has = true;
                if (mnode == -1 || mnode == n)
                {
                    const int step = (int)info(n, 1 + h, t);
                    if (step - t < bestq)
                    {
                        bestq = step - t;
                        blane = t;
                        bstep = step;
                    }
                }
            }
            if (has)
            {
                if (mnode == -1)
                    mnode = n;
                if (ntop < 2)
                    ++ntop;
            }
            if (hasreg && nreg < 2)
                ++nreg;
        }
        o.score[h] = S;
        o.n_top[h] = n_top_rule(S, ntop, nreg);
        if (S == 0)
        {
            // every real cell is 0: gssw leaves ref_end = -1 -> empty CIGAR, position 0 (gssw.c:2728-2732)
            o.max_node[h] = -1;
            o.end_step[h] = 0;
            o.end_lane[h] = 0;
        }
        else
        {
            o.max_node[h] = mnode;
            o.end_step[h] = bstep;
            o.end_lane[h] = blane;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// strand choice (GraphAligner.cpp:340-356) from the forward-graph and reversed-graph fills of one read
// ---------------------------------------------------------------------------------------------
constexpr unsigned AF_CIGAR = 1u, AF_BOTH_STRANDS = 2u, AF_REVERSE_GRAPH = 4u;

struct Decision
{
    int half;    // which packed half of the forward-graph fill is reported (1 = reverse-complemented read)
    int unique;
    int score;
};

PG_HD Decision decide_strand(const TaskOut& fw, const TaskOut& rv, unsigned flags)
{
    const bool both = (flags & AF_BOTH_STRANDS) != 0, rg = (flags & AF_REVERSE_GRAPH) != 0;
    const bool fwd_multi = fw.n_top[0] > 1;
    const bool rev_multi = both ? fw.n_top[1] > 1 : false;
    const bool rfwd_multi = rg ? rv.n_top[0] > 1 : false;
    const bool rrev_multi = (rg && both) ? rv.n_top[1] > 1 : false;
    const bool fwd_unique = !fwd_multi && !rfwd_multi;
    const bool rev_unique = !rev_multi && !rrev_multi;
    bool ret_rev = false;
    if (!fwd_unique && rev_unique && both)
        ret_rev = true;
    else if (fwd_unique && !rev_unique)
        ret_rev = false;
    else if (both)
        ret_rev = fw.score[0] < fw.score[1];
    Decision d;
    d.half = ret_rev ? 1 : 0;
    d.unique = ret_rev ? rev_unique : fwd_unique;
    d.score = fw.score[d.half];
    return d;
}

// ---------------------------------------------------------------------------------------------
// Which reversed-graph fills does a read really need?  (pairing of reversed-graph tasks, pg_kernels.cu)
// ---------------------------------------------------------------------------------------------
// GraphAligner::alignRead fills the reversed graph twice per read only to learn rfwd_multi / rrev_multi, and the
// strand rule above often does not look at both: a strand the forward-graph fill already found non-unique stays
// non-unique whatever its reversed-graph fill says, and once the better-scoring strand turns out unique the other
// strand's uniqueness cannot change the answer.  rv_known[h] = -1 while the reversed-graph result of half h is unknown,
// else its n_top.  rev_plan returns the half to fill next, or -1 when the decision is settled: the decision is
// evaluated for every completion of the unknowns and settled iff they all agree.
PG_HD Decision decide_with(const TaskOut& fw, int n0, int n1, unsigned flags)
{
    TaskOut rv;
    rv.n_top[0] = n0;
    rv.n_top[1] = n1;
    return decide_strand(fw, rv, flags);
}
PG_HD int rev_plan(const TaskOut& fw, const int* rv_known, unsigned flags)
{
    if (!(flags & AF_REVERSE_GRAPH))
        return -1;
    const bool both = (flags & AF_BOTH_STRANDS) != 0;
    // completions: unknown -> {1 (not multi), 2 (multi)}
    bool first = true, settled = true;
    Decision ref;
    ref.half = ref.unique = ref.score = 0;
    for (int c0 = 0; c0 < 2; ++c0)
        for (int c1 = 0; c1 < 2; ++c1)
        {
            if ((rv_known[0] >= 0 && c0) || ((rv_known[1] >= 0 || !both) && c1))
                continue;
            const Decision d = decide_with(fw, rv_known[0] >= 0 ? rv_known[0] : 1 + c0,
                                           rv_known[1] >= 0 ? rv_known[1] : 1 + c1, flags);
            if (first)
            {
                ref = d;
                first = false;
            }
            else if (d.half != ref.half || d.unique != ref.unique)
                settled = false;
        }
    if (settled)
        return -1;
    // unknown halves that can still matter: the forward-graph fill did not already call that strand non-unique
    const bool want0 = rv_known[0] < 0 && fw.n_top[0] <= 1;
    const bool want1 = both && rv_known[1] < 0 && fw.n_top[1] <= 1;
    if (want0 && want1) // the better-scoring strand first: if it is unique the other one is never needed
        return fw.score[0] >= fw.score[1] ? 0 : 1;
    return want0 ? 0 : (want1 ? 1 : -1);
}

// ---------------------------------------------------------------------------------------------
// output records (these two structs ARE the C-ABI result layout, see include/pg_align.h)
// ---------------------------------------------------------------------------------------------
constexpr int STAGE_GSSW = 0, STAGE_PATH = 1, STAGE_GSSW_REV = 2, STAGE_KMER = 3, STAGE_KMER_REV = 4, STAGE_GSSW_REV2 = 5; // Record::mapped_by (include/pg_align.h)
constexpr int ST_UNMAPPED = 3;                // Record::status: no enabled stage mapped the read
struct Record
{
    int32_t graph_pos;
    int16_t score;          // gssw_graph_mapping::score is an int16_t too (gssw.h)
    uint16_t query_clipped; // soft-clipped query bases: what readfilters::BadAlign sums with numClipped() (BadAlign.hh:62-73)
    uint8_t unique;
    uint8_t chose_reverse;
    uint8_t status; // 0 ok, 1 traceback dead end (never seen; the reference would spin/assert), 2 cigar overflow, 3 unmapped
    uint8_t mapped_by; // stage of the CompositeAligner cascade that mapped the read: 0 gssw (DP), 1 PathAligner (pg_path.cuh)
    uint32_t cigar_off; // index of the first op in the cigar arena
    uint32_t cigar_len; // number of ops
};
// cigar op word: node id << 16 | length << 3 | op code
// OP_NONE (length 0) marks a node the path touches without emitting an op: gssw's "alignment start" step pushes
// nothing when the characters differ although their codes match ('U' is coded as A, gssw.c:4214, 1662-1676), and
// extractCigar still prints the node as "id[]".
enum Op { OP_M = 0, OP_X = 1, OP_N = 2, OP_I = 3, OP_D = 4, OP_S = 5, OP_NONE = 7 };
PG_HD uint32_t cigar_word(int node, int op, int len) { return ((uint32_t)node << 16) | ((uint32_t)len << 3) | (uint32_t)op; }

// Replay an op log (traceback order, one word per move) front to back of the path, merging runs of equal
// (node, op) like gssw_cigar_push_back/_front do within a node cigar (gssw.c:3679-3700).  An OP_NONE marker
// survives only if its node has no other op.  Returns the number of ops (writes at most cap of them).
PG_HD int emit_cigar(const uint32_t* oplog, int n, uint32_t* out, int cap)
{
    int m = 0;
    uint32_t prev = 0;
    bool have = false;
    for (int x = n - 1; x >= 0; --x)
    {
        const uint32_t e = oplog[x];
        if (have && (prev >> 16) == (e >> 16))
        {
            if ((e & 7u) == OP_NONE)
                continue;
            if ((prev & 7u) == OP_NONE)
            {
                prev = e;
                continue;
            }
            if ((prev & 7u) == (e & 7u))
            {
                prev += e & 0xFFF8u;
                continue;
            }
        }
        if (have)
        {
            if (m < cap)
                out[m] = prev;
            ++m;
        }
        prev = e;
        have = true;
    }
    if (have)
    {
        if (m < cap)
            out[m] = prev;
        ++m;
    }
    return m;
}

// ---------------------------------------------------------------------------------------------
// traceback over recomputed tiles
// ---------------------------------------------------------------------------------------------
// A tile holds H/E/F (bytes, the chosen half, clamped at 0 like gssw's unsigned saturation) of all rows for
// CK consecutive wavefront steps; cell (node n, column i in node, row j) lives at step node_start[n]+i+j/R.
// steps of tile T the walk can still look at after a miss at need_step (see the trace kernel's recompute phase)
PG_HD int tile_steps_needed(int need_step, int T)
{
    const int n = need_step - T * TS + 3;
    return n < TS ? n : TS;
}
template <int R> struct TileBuf
{
    uint32_t* mem;    // [2 slots][TS][BAND_ROWS] cell words
    int tile0, tile1; // tile index resident in each slot, -1 = empty
    int blo0, blo1;   // first lane of each slot's row band
    int lru;          // slot to evict next
    PG_HD const uint32_t* find(int step, int row) const
    {
        if (step < 0)
            return nullptr;
        const int T = step / TS, ln = row / R;
        int sl = -1, blo = 0;
        constexpr int BAND_LANES = TileGeom<R>::BAND_LANES;
        if (tile0 == T && ln >= blo0 && ln < blo0 + BAND_LANES)
        {
            sl = 0;
            blo = blo0;
        }
        else if (tile1 == T && ln >= blo1 && ln < blo1 + BAND_LANES)
        {
            sl = 1;
            blo = blo1;
        }
        if (sl < 0)
            return nullptr;
        return mem + (size_t)(sl * TS + (step - T * TS)) * TileGeom<R>::BAND_ROWS + (row - R * blo);
    }
    // slot that will receive tile T with a band ending at `row`'s lane (round-robin eviction)
    PG_HD int admit(int T, int row, int& blo)
    {
        blo = row / R - (TileGeom<R>::BAND_LANES - 1);
        if (blo < 0)
            blo = 0;
        const int sl = lru;
        lru ^= 1;
        if (sl == 0)
        {
            tile0 = T;
            blo0 = blo;
        }
        else
        {
            tile1 = T;
            blo1 = blo;
        }
        return sl;
    }
};

// a cell of a node's saved last column (the traceback's cross-node moves): H, and the E leaving it clamped at 0
template <int R, int W> PG_HD int last_h(const uint32_t* last, int node, int j, int half)
{
    const uint32_t w = node_row<R, W>(last, node, j / R)[j % R];
    return is_wide(R, W) ? half16(w, half) : (int)((w >> (8 * half)) & 0xffu);
}
template <int R, int W> PG_HD int last_e(const uint32_t* last, int node, int j, int half)
{
    if (is_wide(R, W))
        return imax0(half16(node_row<R, W>(last, node, j / R)[R + j % R], half));
    return (int)((node_row<R, W>(last, node, j / R)[j % R] >> (16 + 8 * half)) & 0xffu);
}

struct Walker // traceback state of one read (lane 0 only)
{
    int n, i, j;   // current node, column in node, read row
    int st;        // 0 = H, 1 = E (gap in read, 'D'), 2 = F (gap in reference, 'I')
    int v;         // running score == M[st][i][j]
    int phase;     // 0 locate end row, 1 walking, 2 done
    int end_clip;  // trailing soft clip still to emit
    int nops;      // raw ops pushed so far
    int status;
    int need_step; // on a miss: wavefront step whose tile must be made resident ...
    int need_row;  // ... with a row band ending at this row's lane
    int clipped;   // soft-clipped query bases (leading + trailing 'S'), for the BadAlign read filter
    int position;
    int cert_off;  // certified stretches (cert_stretch) are not attempted while set: the last attempt failed here
};

// op log: one cigar_word per traceback move (length 1) or soft clip, in traceback order (back to front)
PG_HD void push_op(Walker& w, uint32_t* oplog, int cap, int node, int op, int len)
{
    if (op == OP_S)
        w.clipped += len;
    if (w.nops < cap)
        oplog[w.nops] = cigar_word(node, op, len);
    else
        w.status = 2;
    ++w.nops;
}
PG_HD int match_op(uint8_t refc, uint8_t readc) { return (refc == 'N' || readc == 'N') ? OP_N : (refc == readc ? OP_M : OP_X); }

// Diagonal runs, W cells at a time.  Most traceback moves are diagonal (M/X/N); instead of one serial move per
// iteration every lane ell probes the cell (i-ell, j-ell): is it an interior cell with positive score whose H equals
// the diagonal neighbour plus the substitution score (gssw.c:1591-1637)?  The run is the number of leading lanes
// that say yes; their ops are logged in parallel.  flag: 0 = no, 1 = yes, 2 = a tile the probe needs is not resident.
template <int R, int W>
PG_HD void diag_probe(int ell, const Walker& w, const TileBuf<R>& tb, const GraphView& g, const uint8_t* chars,
                      const uint8_t* bases, int L, int half, int& flag, int& dval, int& op)
{
    constexpr bool WD = is_wide(R, W);
    const int ii = w.i - ell, jj = w.j - ell;
    flag = 0;
    dval = 0;
    op = OP_M;
    if (ii <= 0 || jj <= 0)
        return;
    const int k = g.node_start[w.n] + ii + jj / R;
    const int kd = k - 1 - ((jj % R) == 0 ? 1 : 0);
    const uint32_t* c0 = tb.find(k, jj);
    const uint32_t* cd = c0 ? tb.find(kd, jj - 1) : nullptr;
    if (!c0 || !cd)
    {
        flag = 2;
        return;
    }
    const int hv = ell == 0 ? w.v : cellH<WD>(*c0);
    if (hv <= 0)
        return;
    const uint8_t refc = chars[g.node_start[w.n] + ii];
    const uint8_t readc = read_char(bases, L, 0, half, jj);
    dval = cellH<WD>(*cd);
    op = match_op(refc, readc);
    flag = (hv == dval + sub_score(nt_code(refc), nt_code(readc))) ? 1 : 0;
}

// Returns the run length (0..W) and logs the run's ops run-length encoded; on 0, miss0 tells that the current
// cell's own probe missed a tile.  Group-uniform on the device: every lane of the W-lane group calls it with
// identical walker state and its own group-lane id; gmask = the group's lanes (other groups of the warp may be
// executing something else).  The host version loops over the W probes.
template <int R, int W>
PG_HD int diag_run(Walker& w, const TileBuf<R>& tb, const GraphView& g, const uint8_t* chars, const uint8_t* bases,
                   int L, int half, int lane, unsigned gmask, uint32_t* oplog, int oplog_cap, int& vnew, bool& miss0)
{
    int run = 0, nent = 0;
#if defined(__CUDA_ARCH__)
    constexpr unsigned WM = (W == 32) ? 0xffffffffu : ((1u << (W & 31)) - 1u);
    const int gshift = __ffs((int)gmask) - 1; // first lane of this group
    int flag, dval, op;
    diag_probe<R, W>(lane, w, tb, g, chars, bases, L, half, flag, dval, op);
    const unsigned yes = (__ballot_sync(gmask, flag == 1) >> gshift) & WM;
    run = (yes == WM) ? W : (__ffs((int)~yes) - 1);
    miss0 = (__shfl_sync(gmask, flag, 0, W) == 2);
    vnew = __shfl_sync(gmask, dval, run > 0 ? run - 1 : 0, W);
    // run-length encode: a lane starts an entry when its op differs from the previous lane's
    const int prev = __shfl_up_sync(gmask, op, 1, W);
    const bool in = lane < run;
    const unsigned starts = (__ballot_sync(gmask, in && (lane == 0 || op != prev)) >> gshift) & WM;
    if (in && ((starts >> lane) & 1u))
    {
        const unsigned above = starts & ~((2u << lane) - 1u); // entry starts after this lane
        const int nxt = above ? (__ffs((int)above) - 1) : run;
        const int idx = w.nops + __popc(starts & ((1u << lane) - 1u));
        if (idx < oplog_cap)
            oplog[idx] = cigar_word(w.n, op, nxt - lane);
    }
    __syncwarp(gmask);
    nent = __popc(starts);
#else
    (void)lane;
    (void)gmask;
    miss0 = false;
    vnew = 0;
    int last_op = -1;
    for (int ell = 0; ell < W; ++ell)
    {
        int flag, dval, op;
        diag_probe<R, W>(ell, w, tb, g, chars, bases, L, half, flag, dval, op);
        if (ell == 0 && flag == 2)
            miss0 = true;
        if (flag != 1)
            break;
        if (op == last_op)
        {
            if (w.nops + nent - 1 < oplog_cap)
                oplog[w.nops + nent - 1] += 1u << 3;
        }
        else
        {
            if (w.nops + nent < oplog_cap)
                oplog[w.nops + nent] = cigar_word(w.n, op, 1);
            ++nent;
            last_op = op;
        }
        vnew = dval;
        ++run;
    }
#endif
    if (run > 0)
    {
        if (w.nops + nent > oplog_cap)
            w.status = 2;
        w.nops += nent;
    }
    return run;
}

// ---------------------------------------------------------------------------------------------
// Certified diagonal stretches: traceback moves without the DP matrices
// ---------------------------------------------------------------------------------------------
// Most alignments are one long diagonal (matches and a few mismatches); recomputing a tile of the matrices every 16
// wavefront steps just to confirm "diagonal again" is most of the traceback kernel's work.  A stretch of diagonal
// moves can be CERTIFIED from the sequences alone, given exact values at its two ends.  Let the walker stand in state
// H at cell 0 = (i, j) of node n with the exact value v_0 = H(cell 0), let cell k = (i - k, j - k), s_k the
// substitution score of cell k and v_{k+1} = v_k - s_k.  For interior cells (i - k > 0, j - k > 0) the recurrence gives
// H(cell k) >= H(cell k+1) + s_k, hence by induction from cell 0:      H(cell k) <= v_k        (upper bounds).
// If further along there is an ANCHOR cell a whose value is known to be at least v_a, the same inequality read the other
// way gives H(cell k) >= H(cell k+1) + s_k >= v_k for all k < a        (lower bounds), so H(cell k) = v_k on the whole
// stretch, every H(cell k) equals H(cell k+1) + s_k, and the reference's traceback -- which tests the diagonal first
// in an interior cell (gssw.c:1591-1637) -- takes the diagonal at every one of them.  Anchors:
//   * zero: v_{a} = 0 with all earlier v_k > 0 (H >= 0 always) -- the alignment starts behind cell a - 1; when cell
//     a - 1 lies in the first row or column the move is the "alignment starts here" rule (v == s, gssw.c:1655-1690),
//     which needs H(cell a-1) >= s only: true for any diagonal input >= 0;
//   * first column: cell a = (0, j') reached with v_a > 0: H, E and F of the first column of a node follow from the
//     saved last columns of its predecessors alone (col0_exact: the seed is their element-wise maximum), so the
//     value is checked exactly, and the reference's first-column order -- start, F, E, then the predecessors in
//     ascending order (gssw.c:1655-1768, 2966-3040) -- is evaluated on exact numbers; the walk continues in the
//     predecessor that explains the score, where the argument starts over.
// Anything else (a partial sum below zero: some gap lies on the true path; the first row reached with score left; F or
// E explaining a first-column value; no predecessor explaining it) certifies nothing: the walker stays where it was
// and the tile walk below takes over from that exact state.  Exact by construction; the emulator and the GPU tests
// compare the result with the reference either way.
#ifndef PG_TRACE_CERT
#define PG_TRACE_CERT 1
#endif

// exact H, E (entering) and F (entering) of cell (0, jq) of node n for the chosen half, from the predecessors' saved
// last columns: t(0, r) = max(seedH(r - 1) + s(0, r), seedE(r), 0), F(jq) = max_{r < jq} (t(0, r) - go - (jq - 1 - r))
template <int R, int W>
PG_HD void col0_exact(const GraphView& g, const uint8_t* chars, const uint32_t* last, const uint8_t* bases, int L, int half,
                      int n, int jq, int lane, unsigned gmask, int& Hq, int& Eq, int& Fq)
{
    const int p0 = g.pred_ptr[n], p1 = g.pred_ptr[n + 1];
    const int rc = nt_code(chars[g.node_start[n]]);
    int best = -100000, tq = 0, eq = 0; // best = max over rows r < jq of t(0, r) + r
#if defined(__CUDA_ARCH__)
    const int r_lo = R * lane, r_hi = R * lane + R;
#else
    (void)lane;
    (void)gmask;
    const int r_lo = 0, r_hi = jq + 1;
#endif
    for (int r = r_lo; r < r_hi && r <= jq; ++r)
    {
        int sh = 0, se = 0;
        for (int e = p0; e < p1; ++e)
        {
            const int c = g.pred_idx[e];
            if (r > 0)
            {
                const int hv = last_h<R, W>(last, c, r - 1, half);
                sh = sh > hv ? sh : hv;
            }
            const int ev = last_e<R, W>(last, c, r, half);
            se = se > ev ? se : ev;
        }
        const int sc = sub_score(rc, nt_code(read_char(bases, L, 0, half, r)));
        int t = sh + sc;
        t = t > se ? t : se;
        t = imax0(t);
        if (r < jq)
            best = best > t + r ? best : t + r;
        else
        {
            tq = t;
            eq = se;
        }
    }
#if defined(__CUDA_ARCH__)
    for (int d = W / 2; d >= 1; d >>= 1)
    {
        const int o = __shfl_xor_sync(gmask, best, d, W);
        best = best > o ? best : o;
    }
    tq = __shfl_sync(gmask, tq, jq / R, W);
    eq = __shfl_sync(gmask, eq, jq / R, W);
#endif
    Fq = jq > 0 ? best - (GAP_OPEN - GAP_EXT) - jq : -100000;
    Eq = eq;
    Hq = tq > Fq ? tq : Fq;
}

// Advance the walker (state H, exact w.v > 0) over certified stretches.  Returns true when it moved (w updated, ops
// logged); false when nothing could be certified from here.  Group-uniform on the device like diag_run.
template <int R, int W>
PG_HD bool cert_stretch(Walker& w, const GraphView& g, const uint8_t* chars, const uint32_t* last, const uint8_t* bases, int L,
                        int half, uint32_t* oplog, int oplog_cap, int lane, unsigned gmask)
{
    bool moved = false;
    for (;;)
    {
        const int n = w.n, i = w.i, j = w.j, m = i < j ? i : j;
        const int col0 = g.node_start[n];
        int nops = w.nops;  // tentative until the stretch is certified
        int vb = w.v;       // value entering the chunk's first cell
        int k0 = -1;        // cell after which the value is 0 (zero anchor), -1 = none
        int sm = 0, opm = OP_M, vm = 0; // cell m: its score, its diagonal op, its value
        bool failed = false;
#if defined(__CUDA_ARCH__)
        constexpr unsigned WM = (W == 32) ? 0xffffffffu : ((1u << (W & 31)) - 1u);
        const int gshift = __ffs((int)gmask) - 1;
        for (int base = 0; base <= m && k0 < 0; base += W)
        {
            const int k = base + lane;
            const bool valid = k <= m;
            const uint8_t refc = valid ? chars[col0 + i - k] : (uint8_t)'A';
            const uint8_t readc = valid ? read_char(bases, L, 0, half, j - k) : (uint8_t)'A';
            const int sc = valid ? sub_score(nt_code(refc), nt_code(readc)) : 0;
            int cs = sc; // inclusive prefix sum over the group
            for (int d = 1; d < W; d <<= 1)
            {
                const int o = __shfl_up_sync(gmask, cs, d, W);
                if (lane >= d)
                    cs += o;
            }
            const int v1 = vb - cs; // value behind cell k
            const unsigned neg = (__ballot_sync(gmask, valid && v1 < 0) >> gshift) & WM;
            const unsigned zero = (__ballot_sync(gmask, valid && v1 == 0) >> gshift) & WM;
            int cnt = m - base + 1 < W ? m - base + 1 : W; // cells of this chunk that are moves
            if (neg | zero)
            {
                const int e = __ffs((int)(neg | zero)) - 1;
                if ((neg >> e) & 1u)
                {
                    failed = true;
                    break;
                }
                k0 = base + e;
                cnt = e + 1;
            }
            if (base + cnt - 1 == m) // the chunk holds cell m: its numbers for the first-row / first-column rules
            {
                const int lm = m - base;
                sm = __shfl_sync(gmask, sc, lm, W);
                vm = __shfl_sync(gmask, v1 + sc, lm, W);
                opm = __shfl_sync(gmask, match_op(refc, readc), lm, W);
                if (k0 == m || k0 < 0)
                    --cnt; // cell m is logged by the rule that applies to it
            }
            // run-length encode the diagonal ops of lanes [0, cnt)
            const int op = match_op(refc, readc);
            const int prev = __shfl_up_sync(gmask, op, 1, W);
            const bool in = lane < cnt;
            const unsigned starts = (__ballot_sync(gmask, in && (lane == 0 || op != prev)) >> gshift) & WM;
            if (in && ((starts >> lane) & 1u))
            {
                const unsigned above = starts & ~((2u << lane) - 1u);
                const int nxt = above ? (__ffs((int)above) - 1) : cnt;
                const int idx = nops + __popc(starts & ((1u << lane) - 1u));
                if (idx < oplog_cap)
                    oplog[idx] = cigar_word(n, op, nxt - lane);
            }
            __syncwarp(gmask); // these slots are written again (same words, or by a later push_op when nothing was certified)
            nops += __popc(starts);
            vb = __shfl_sync(gmask, v1, W - 1, W);
        }
#else
        (void)lane;
        (void)gmask;
        {
            int v = vb, last_op = -1;
            for (int k = 0; k <= m; ++k)
            {
                const uint8_t refc = chars[col0 + i - k];
                const uint8_t readc = read_char(bases, L, 0, half, j - k);
                const int sc = sub_score(nt_code(refc), nt_code(readc));
                const int v1 = v - sc;
                if (v1 < 0)
                {
                    failed = true;
                    break;
                }
                if (k == m)
                {
                    sm = sc;
                    vm = v;
                    opm = match_op(refc, readc);
                }
                if (v1 == 0)
                    k0 = k;
                if (k < m) // a logged diagonal move (cell m is logged by the rule that applies to it)
                {
                    const int op = match_op(refc, readc);
                    if (op == last_op)
                    {
                        if (nops - 1 < oplog_cap)
                            oplog[nops - 1] += 1u << 3;
                    }
                    else
                    {
                        if (nops < oplog_cap)
                            oplog[nops] = cigar_word(n, op, 1);
                        ++nops;
                        last_op = op;
                    }
                }
                v = v1;
                if (k0 >= 0)
                    break;
            }
        }
#endif
        if (failed)
            return moved;
        if (nops > oplog_cap)
            w.status = 2;
        if (k0 >= 0) // zero anchor: the alignment starts at cell k0
        {
            if (k0 == m) // ... which lies in the first row or column: "alignment starts here" (gssw.c:1655-1690)
            {
                const uint8_t refc = chars[col0 + i - m];
                const uint8_t readc = read_char(bases, L, 0, half, j - m);
                w.nops = nops;
                if (refc == 'N' || readc == 'N' || refc == readc)
                    push_op(w, oplog, oplog_cap, n, (refc == 'N' || readc == 'N') ? OP_N : OP_M, 1);
                else
                    push_op(w, oplog, oplog_cap, n, OP_NONE, 0);
            }
            else
                w.nops = nops;
            w.i = i - (k0 + 1);
            w.j = j - (k0 + 1);
            w.v = 0;
            return true;
        }
        // cell m reached with score left
        if (i - m != 0)
            return moved; // first row before first column: only E could explain it -- not certified
        const int jq = j - m;
        int Hq, Eq, Fq;
        col0_exact<R, W>(g, chars, last, bases, L, half, n, jq, lane, gmask, Hq, Eq, Fq);
        if (Hq != vm)
            return moved; // the diagonal hypothesis does not hold
        // the stretch (cells 0 .. m-1) is certified; the walker now stands at (0, jq) with the exact value vm
        w.nops = nops;
        w.i = 0;
        w.j = jq;
        w.v = vm;
        moved = moved || m > 0;
        if ((jq > 0 && vm == Fq) || vm == Eq || jq == 0)
            return moved; // a gap (or a dead end) at the node's first column: the tile walk handles it from here
        int best = -1;
        for (int e = g.pred_ptr[n]; e < g.pred_ptr[n + 1] && best < 0; ++e) // gssw.c:2966-3040
            if (vm == last_h<R, W>(last, g.pred_idx[e], jq - 1, half) + sm)
                best = g.pred_idx[e];
        if (best < 0)
            return moved;
        moved = true;
        push_op(w, oplog, oplog_cap, n, opm, 1);
        w.v = vm - sm;
        w.j = jq - 1;
        w.n = best;
        w.i = g.node_len[best] - 1;
        if (w.v <= 0)
            return true;
    }
}

// Walk as far as the resident tiles allow.  Returns true when finished (w.phase == 2), false on a tile miss
// (w.need_step set).  Mirrors gssw_alignment_trace_back_byte (gssw.c:1112-1818, final_traceback = 1, no
// deflections) and the cross-node part of gssw_graph_trace_back_internal (gssw.c:2836-3148, 3486-3528).
//   g      forward graph view;  chars = upper-cased graph characters (column-indexed like codes)
//   last   this read's node table (rows per (node, lane), Sizes<R, W>::ROWW words: the saved last columns)
template <int R, int W>
PG_HD bool walk(Walker& w, const TileBuf<R>& tb, const GraphView& g, const uint8_t* chars, const uint32_t* last,
                const uint8_t* bases, int L, int half, const TaskOut& fo, uint32_t* oplog, int oplog_cap, int lane,
                unsigned gmask)
{
    constexpr bool WD = is_wide(R, W);
    if (w.phase == 0)
    {
        // end cell: smallest row of end_lane holding S at end_step (gssw.c:446-454)
        const int S = fo.score[half];
        if (S <= 0)
        {
            w.position = 0;
            w.phase = 2;
            return true;
        }
        int row = -1;
        for (int r = 0; r < R; ++r)
        {
            const uint32_t* t0 = tb.find(fo.end_step[half], R * fo.end_lane[half] + r);
            if (!t0)
            {
                w.need_step = fo.end_step[half];
                w.need_row = R * fo.end_lane[half] + R - 1;
                return false;
            }
            if (cellH<WD>(*t0) == S)
            {
                row = R * fo.end_lane[half] + r;
                break;
            }
        }
        if (row < 0)
        {
            w.status = 1;
            w.phase = 2;
            return true;
        }
        w.n = fo.max_node[half];
        w.i = fo.end_step[half] - fo.end_lane[half] - g.node_start[w.n];
        w.j = row;
        w.st = 0;
        w.v = S;
        w.end_clip = L - 1 - row; // gssw.c:2766-2773
        if (w.end_clip > 0)
            push_op(w, oplog, oplog_cap, w.n, OP_S, w.end_clip);
        w.phase = 1;
    }
    while (true)
    {
        // ---------------- inside node w.n (gssw_alignment_trace_back_byte) ----------------
        bool leave = false; // left the node through its first column
        while (w.v > 0 && w.i >= 0 && w.j >= 0)
        {
            if (PG_TRACE_CERT && w.st == 0 && !w.cert_off)
            {
                if (cert_stretch<R, W>(w, g, chars, last, bases, L, half, oplog, oplog_cap, lane, gmask))
                    continue;
                w.cert_off = 1; // nothing certified from here: tiles, until the walk is past whatever is in the way
            }
            const int k = g.node_start[w.n] + w.i + w.j / R; // step of the current cell
            // neighbours: (i-1,j-1) -> step k-1 or k-2; (i-1,j) -> k-1; (i,j-1) -> k or k-1
            const uint32_t* c0 = tb.find(k, w.j);
            if (!c0)
            {
                w.need_step = k;
                w.need_row = w.j;
                return false;
            }
            const int kl = k - ((w.j % R) == 0 ? 1 : 0);     // step of (i, j-1)
            if (w.st == 1)
            {
                if (w.i == 0)
                {
                    leave = true;
                    break;
                }
                const uint32_t* c1 = tb.find(k - 1, w.j);
                if (!c1)
                {
                    w.need_step = k - 1;
                    w.need_row = w.j;
                    return false;
                }
                if (w.v == cellH<WD>(*c1) - GAP_OPEN) // gssw.c:1347-1383
                {
                    push_op(w, oplog, oplog_cap, w.n, OP_D, 1);
                    w.v += GAP_OPEN;
                    --w.i;
                    w.st = 0;
                    w.cert_off = 0; // past the gap: the rest may be one certified diagonal again
                    continue;
                }
                if (w.v == cellE<WD>(*c1) - GAP_EXT) // gssw.c:1400-1423
                {
                    push_op(w, oplog, oplog_cap, w.n, OP_D, 1);
                    w.v += GAP_EXT;
                    --w.i;
                    continue;
                }
                w.status = 1; // "Stuck in read gap" (gssw.c:1449-1454)
                w.phase = 2;
                return true;
            }
            if (w.st == 2)
            {
                if (w.j > 0)
                {
                    const uint32_t* cl = tb.find(kl, w.j - 1);
                    if (!cl)
                    {
                        w.need_step = kl;
                        w.need_row = w.j;
                        return false;
                    }
                    if (w.v == cellH<WD>(*cl) - GAP_OPEN) // gssw.c:1458-1493
                    {
                        push_op(w, oplog, oplog_cap, w.n, OP_I, 1);
                        w.v += GAP_OPEN;
                        --w.j;
                        w.st = 0;
                        w.cert_off = 0;
                        continue;
                    }
                    if (w.v == cellF<WD>(*cl) - GAP_EXT) // gssw.c:1510-1532
                    {
                        push_op(w, oplog, oplog_cap, w.n, OP_I, 1);
                        w.v += GAP_EXT;
                        --w.j;
                        continue;
                    }
                }
                w.status = 1; // "Ref gap stuck"
                w.phase = 2;
                return true;
            }
            // H state (gssw.c:1562-1801)
            const uint8_t refc = chars[g.node_start[w.n] + w.i];
            const uint8_t readc = read_char(bases, L, 0, half, w.j);
            const int s = sub_score(nt_code(refc), nt_code(readc));
            if (w.i > 0 && w.j > 0)
            {
                int vnew;
                bool miss0;
                const int run = diag_run<R, W>(w, tb, g, chars, bases, L, half, lane, gmask, oplog, oplog_cap, vnew, miss0);
                if (run > 0) // diagonal moves, gssw.c:1591-1637
                {
                    w.v = vnew;
                    w.i -= run;
                    w.j -= run;
                    continue;
                }
                if (miss0) // the diagonal neighbour (i-1, j-1) is not resident
                {
                    w.need_step = k - 1 - ((w.j % R) == 0 ? 1 : 0);
                    w.need_row = w.j;
                    return false;
                }
            }
            else if (w.v == s) // alignment starts here, gssw.c:1655-1690
            {
                if (refc == 'N' || readc == 'N' || refc == readc)
                    push_op(w, oplog, oplog_cap, w.n, (refc == 'N' || readc == 'N') ? OP_N : OP_M, 1);
                else
                    push_op(w, oplog, oplog_cap, w.n, OP_NONE, 0);
                --w.i;
                --w.j;
                w.v -= s;
                continue;
            }
            if (w.j > 0 && w.v == cellF<WD>(*c0)) // H == F, gssw.c:1709-1729
            {
                w.st = 2;
                continue;
            }
            if (w.v == cellE<WD>(*c0)) // H == E, gssw.c:1747-1768
            {
                w.st = 1;
                continue;
            }
            if (w.i == 0) // gssw.c:1787-1794
            {
                leave = true;
                break;
            }
            w.status = 1; // "Stuck in main matrix"
            w.phase = 2;
            return true;
        }
        // ---------------- between nodes (gssw_graph_trace_back_internal) ----------------
        if (!leave || w.v == 0)
        {
            // score exhausted (gssw.c:2836-2844): leading soft clip of readEnd+1, position = refEnd+1
            if (w.v != 0) // ran off the read/reference with score left: reference flags gm->score = -1 (:2824-2828)
                w.status = 1;
            if (w.j > -1)
                push_op(w, oplog, oplog_cap, w.n, OP_S, w.j + 1);
            w.position = w.i + 1 < 0 ? 0 : w.i + 1;
            w.phase = 2;
            return true;
        }
        // first predecessor (ascending id) that explains the score wins (gssw.c:2966-3148)
        int best = -1;
        const uint8_t refc = chars[g.node_start[w.n]];
        const uint8_t readc = read_char(bases, L, 0, half, w.j);
        const int s = sub_score(nt_code(refc), nt_code(readc));
        for (int e = g.pred_ptr[w.n]; e < g.pred_ptr[w.n + 1]; ++e)
        {
            const int c = g.pred_idx[e];
            if (w.st == 0)
            {
                // diagonal source = pred's last column at row j-1 (row -1 never matches: H(0,0) is start or E)
                const int dsrc = w.j > 0 ? last_h<R, W>(last, c, w.j - 1, half) : -1000;
                if (w.v == dsrc + s) // gssw.c:2999-3040
                {
                    best = c;
                    push_op(w, oplog, oplog_cap, w.n, match_op(refc, readc), 1);
                    w.v -= s;
                    --w.j;
                    break;
                }
            }
            else
            {
                const int hsrc = last_h<R, W>(last, c, w.j, half);
                if (w.v == hsrc - GAP_OPEN) // open, gssw.c:3089-3110
                {
                    best = c;
                    push_op(w, oplog, oplog_cap, w.n, OP_D, 1);
                    w.v += GAP_OPEN;
                    w.st = 0;
                    w.cert_off = 0;
                    break;
                }
                // extend: the reference tests v == E_c(last, j) - ge with E *entering* pred's last column
                // (gssw.c:3122-3136).  Saved is the seed E' = max(E - ge, t - go) >= E - ge, and v >= E' (v is the max
                // of the seeds), so v == E' is necessary; only then is the exact E read from pred's last-column tile.
                const int eseed = last_e<R, W>(last, c, w.j, half);
                if (w.v == eseed)
                {
                    const int kc = g.node_start[c] + g.node_len[c] - 1 + w.j / R;
                    const uint32_t* cc = tb.find(kc, w.j);
                    if (!cc)
                    {
                        w.need_step = kc;
                        w.need_row = w.j;
                        return false; // nothing has been changed yet: the predecessor scan restarts after the reload
                    }
                    if (w.v == cellE<WD>(*cc) - GAP_EXT)
                    {
                        best = c;
                        push_op(w, oplog, oplog_cap, w.n, OP_D, 1);
                        w.v += GAP_EXT;
                        break;
                    }
                }
            }
        }
        if (best < 0)
        {
            // no predecessor explains the score (reference: assert compiled out, soft-clips the rest, gssw.c:3500-3517)
            w.status = 1;
            if (w.j > -1)
                push_op(w, oplog, oplog_cap, w.n, OP_S, w.j + 1);
            w.position = w.i + 1 < 0 ? 0 : w.i + 1;
            w.phase = 2;
            return true;
        }
        w.n = best;
        w.i = g.node_len[best] - 1; // gssw.c:3486-3499
    }
}

} // namespace pg
