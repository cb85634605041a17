// TEST INFRASTRUCTURE ONLY -- CPU lane emulator of the CUDA path.
// Compiles paragraph_b200/csrc/pg_core.cuh (the exact device source: recurrence, node events,
// checkpoints, finalisation, strand choice, tile traceback) for the host and drives 32 emulated lanes in
// lock step, with the warp shuffles replaced by array reads.  It exists so that the kernel logic can be
// fuzzed against the oracle on machines without a GPU; nothing in the product loads it.
#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../paragraph_b200/csrc/pg_core.cuh"
#include "../../paragraph_b200/csrc/pg_count.cuh"
#include "../../paragraph_b200/csrc/pg_host.hpp"

using namespace pg;

static long g_rev_rounds = 0, g_rev_reads = 0; // reversed-graph halves rev_plan asked for / reads (see emu_align_one)
// Speculative "no gap alive" blocks (pg_core.cuh: lane_step_dead) -- what pg_fill_kernel does when built with
// PG_SPEC_DEAD=1 (the default build).  On by default, like in the kernels; pgemu_set_spec(0) gives the plain fill.
static int g_spec = 1;
static long g_spec_pruned = 0;
extern "C" long pgemu_spec_pruned() { return g_spec_pruned; }
static int g_dead_boundary = 0;                  // boundary sub-blocks may run dead as well (pg_fill_kernel: PG_DEAD_BOUNDARY builds)
static long g_dead_boundary_blocks[2] = { 0, 0 }; // run dead, redone
extern "C" void pgemu_set_dead_boundary(int on) { g_dead_boundary = on; }
extern "C" void pgemu_dead_boundary_stats(long* out) { out[0] = g_dead_boundary_blocks[0]; out[1] = g_dead_boundary_blocks[1]; }
static long g_spec_blocks[4] = { 0, 0, 0, 0 }; // blocks of SPEC_STEPS steps: run dead, redone, gaps alive, node boundary inside
extern "C" void pgemu_set_spec(int on) { g_spec = on; }
extern "C" void pgemu_spec_stats(long* o)
{
    for (int i = 0; i < 4; ++i)
        o[i] = g_spec_blocks[i];
}
extern "C" void pgemu_rev_stats(long* o)
{
    o[0] = g_rev_rounds;
    o[1] = g_rev_reads;
}

namespace
{

template <int R, int W>
void emu_fill_pass(const GraphView& g, const uint8_t* bases, int L, int orient, bool save_trace, std::vector<uint32_t>& info,
                   std::vector<uint32_t>& last, std::vector<uint32_t>& ckpt, TaskOut& out, bool precise);

// pg_fill_kernel's two passes: speculative blocks with the cheap bookkeeping first; a forward-graph fill whose top score
// lies in the range those blocks keep no first-reached steps for is done once more, exactly (pg_core.cuh: dead_range_score)
static long g_second_pass = 0;
extern "C" long pgemu_second_passes() { return g_second_pass; }
template <int R, int W>
void emu_fill(const GraphView& g, const uint8_t* bases, int L, int orient, bool save_trace, std::vector<uint32_t>& info,
              std::vector<uint32_t>& last, std::vector<uint32_t>& ckpt, TaskOut& out)
{
    emu_fill_pass<R, W>(g, bases, L, orient, save_trace, info, last, ckpt, out, false);
    if (g_spec == 1 && !Sizes<R, W>::WIDE && save_trace && (dead_range_score(out.score[0]) || dead_range_score(out.score[1])))
    {
        ++g_second_pass;
        emu_fill_pass<R, W>(g, bases, L, orient, save_trace, info, last, ckpt, out, true);
    }
}

template <int R, int W>
void emu_fill_pass(const GraphView& g, const uint8_t* bases, int L, int orient, bool save_trace, std::vector<uint32_t>& info,
                   std::vector<uint32_t>& last, std::vector<uint32_t>& ckpt, TaskOut& out, bool precise)
{
    std::vector<uint32_t> prof((size_t)NCODE * R * W);
    std::vector<uint32_t> seedS((size_t)g.n_nodes * Sizes<R, W>::ROWW * W, 0); // node table (pg_core.cuh: Sizes)
    for (int t = 0; t < W; ++t)
        build_profile<R, W>(prof.data(), bases, L, orient, t);
    constexpr bool WIDE = Sizes<R, W>::WIDE;
    constexpr int CKW = Sizes<R, W>::CKW;
    Lane<R> s[W];
    LaneCtl c[W];
    for (int t = 0; t < W; ++t)
    {
        lane_zero(s[t]);
        ctl_at_step(c[t], g, 0, t);
        if (WIDE)
            region_begin(c[t], g, L, t);
    }
    if (save_trace)
        ckpt.assign((size_t)num_ckpt(g.G, W) * CKW * W, 0);
    const int nck = num_ckpt(g.G, W);
    int no_spec_before = 0;
    // node boundaries inside boundary sub-blocks as the kernel handles them (pg_core.cuh: entry_word, seed_prefetch,
    // node_event_pre): entry words per node, the older predecessors' part of a seed folded at the sub-block's start
    std::vector<uint32_t> entry((size_t)g.n_nodes + 1, 0);
    for (int m = 1; m < g.n_nodes; ++m)
        entry[m] = entry_word(g, m);
    SeedPre<R> pre[W];
    for (int t = 0; t < W; ++t)
        pre[t].node = -1;
    const bool lean = g_spec && !WIDE;
    int sbest[2] = { 0, 0 }, pending[2] = { 0, 0 }; // EXPERIMENT (g_spec == 2): best score so far, folded in at sub-block starts
    for (int k = 0; k < nck * CK; ++k)
    {
        if (save_trace && k % CK == 0)
            for (int t = 0; t < W; ++t)
                ckpt_store<R, W>(s[t], ckpt.data() + (size_t)(k / CK) * CKW * W, t);
        if (g_spec && !precise && !WIDE && k % SPEC_STEPS == 0 && k >= no_spec_before)
        {
            bool flat = true, dead = true;
            for (int h = 0; h < 2; ++h)
                sbest[h] = std::max(sbest[h], pending[h]);
            for (int t = 0; t < W; ++t)
            {
                flat = flat && c[t].colsLeft >= SPEC_STEPS;
                dead = dead && !gaps_alive(s[t]);
            }
            if (g_spec == 2 && flat && !dead)
            {
                // upper-bound pruning (pg_core.cuh: gap_relevant), with the packed helpers the kernel would use
                const uint32_t Sb = pk(sbest[0], sbest[1]);
                bool relevant = false;
                for (int t = 0; t < W; ++t)
                    relevant = relevant
                        || lane_gaps_relevant<R>(s[t], pk(L - 1 - R * t, L - 1 - R * t), pk(L - 1 - R * (t + 1), L - 1 - R * (t + 1)), Sb);
                if (!relevant)
                {
                    for (int t = 0; t < W; ++t)
                        lane_gaps_drop<R>(s[t]);
                    dead = true;
                    ++g_spec_pruned;
                }
            }
            ++g_spec_blocks[!flat ? 3 : (!dead ? 2 : 0)];
            for (int t = 0; t < W; ++t)
            {
                pre[t].node = -1;
                pre[t].twice = false;
            }
            if (!flat)
                for (int t = 0; t < W; ++t)
                    seed_prefetch<R, W>(pre[t], c[t], g, entry.data(), t, seedS.data(), SPEC_STEPS);
            // a boundary sub-block with the collapsed recurrence (pg_fill_kernel, PG_DEAD_BOUNDARY): no live gap in any lane,
            // none brought by a merged seed, no lane crossing twice; the events run between the collapsed steps
            if (!flat && g_spec == 1 && g_dead_boundary)
            {
                bool bad = false;
                for (int t = 0; t < W; ++t)
                    bad = bad || gaps_alive(s[t]) || seed_live(pre[t]);
                if (!bad)
                {
                    DeadSave<R> keep[W];
                    LaneCtl keepc[W];
                    uint32_t Mt[W], Mall = 0u;
                    bool brk = false;
                    for (int t = 0; t < W; ++t)
                    {
                        dead_save(s[t], keep[t]);
                        keepc[t] = c[t];
                        Mt[t] = 0u;
                    }
                    for (int kk = 0; kk < SPEC_STEPS; ++kk)
                    {
                        for (int t = 0; t < W; ++t)
                            if (c[t].colsLeft == 0)
                            {
                                c[t].Mnode = max2(c[t].Mnode, add2(Mt[t], pk(-MBIAS, -MBIAS)));
                                Mall = max2(Mall, Mt[t]);
                                Mt[t] = 0u;
                                node_event_pre<R, W>(s[t], c[t], g, entry.data(), pre[t], t, seedS.data());
                                brk = brk || e_alive(s[t]);
                            }
                            else
                                --c[t].colsLeft;
                        uint32_t rh[W];
                        for (int t = 0; t < W; ++t)
                            rh[t] = t ? s[t - 1].hbotLast : 0;
                        for (int t = 0; t < W; ++t)
                        {
                            const ProfPtr<W> pf = { prof.data() + (g.codes[k + kk - t] * R) * W + t };
                            Mt[t] = lane_step_dead<R>(s[t], rh[t], pf, 0u, Mt[t]);
                            pending[0] = std::max(pending[0], lo16(Mt[t]));
                            pending[1] = std::max(pending[1], hi16(Mt[t]));
                        }
                    }
                    for (int t = 0; t < W; ++t)
                        Mall = max2(Mall, Mt[t]);
                    if (!brk && !dead_block_broken(Mall))
                    {
                        for (int t = 0; t < W; ++t)
                            c[t].Mnode = max2(c[t].Mnode, add2(Mt[t], pk(-MBIAS, -MBIAS)));
                        ++g_dead_boundary_blocks[0];
                        k += SPEC_STEPS - 1;
                        continue;
                    }
                    ++g_dead_boundary_blocks[1];
                    for (int t = 0; t < W; ++t)
                    {
                        dead_restore(s[t], keep[t]);
                        for (int r = 0; r < R; ++r)
                            s[t].E[r] = 0u;
                        c[t] = keepc[t];
                    }
                }
            }
            if (flat && dead)
            {
                DeadSave<R> keep[W];
                LaneCtl keepc[W];
                for (int t = 0; t < W; ++t)
                {
                    dead_save(s[t], keep[t]);
                    keepc[t] = c[t];
                }
                uint32_t Mt = 0u, Mn[W], Mlane[W];
                for (int t = 0; t < W; ++t)
                    Mn[t] = track_t_begin(c[t]);
                for (int t = 0; t < W; ++t)
                    Mlane[t] = 0u;
                for (int kk = 0; kk < SPEC_STEPS; ++kk)
                {
                    uint32_t rh[W];
                    for (int t = 0; t < W; ++t)
                    {
                        --c[t].colsLeft;
                        rh[t] = t ? s[t - 1].hbotLast : 0;
                    }
                    for (int t = 0; t < W; ++t)
                    {
                        const ProfPtr<W> pf = { prof.data() + (g.codes[k + kk - t] * R) * W + t };
                        const uint32_t mt = lane_step_dead<R>(s[t], rh[t], pf);
                        Mt = max2(Mt, mt);
                        Mlane[t] = max2(Mlane[t], mt);
                        pending[0] = std::max(pending[0], lo16(mt));
                        pending[1] = std::max(pending[1], hi16(mt));
                        if (g_spec == 2) // the pruning experiment lets blocks with t > gap_open stay "dead": exact bookkeeping
                            track_t(c[t], Mn[t], mt, k + kk);
                    }
                }
                bool broken = dead_block_broken(Mt);
                if (g_spec == 2 && broken)
                {
                    // new gaps t - go opened in the block are tolerated when they are irrelevant by the same bound
                    broken = false;
                    for (int t = 0; t < W; ++t)
                        broken = broken || dead_block_broken_pruned(Mlane[t], pk(L - 1 - R * t, L - 1 - R * t), pk(sbest[0], sbest[1]));
                }
                if (!broken)
                {
                    for (int t = 0; t < W; ++t)
                        if (g_spec == 2)
                            track_t_end(c[t], Mn[t]);
                        else // fold the block into the node maximum; no first-reached steps
                            c[t].Mnode = max2(c[t].Mnode, add2(Mlane[t], pk(-MBIAS, -MBIAS)));
                    k += SPEC_STEPS - 1;
                    continue;
                }
                --g_spec_blocks[0];
                ++g_spec_blocks[1];
                for (int t = 0; t < W; ++t)
                {
                    dead_restore(s[t], keep[t]);
                    c[t] = keepc[t];
                }
                no_spec_before = k + SPEC_STEPS;
            }
        }
        if (lean && precise && k % SPEC_STEPS == 0) // (the exact second pass takes the same boundary path in the kernel)
        {
            bool flat = true;
            for (int t = 0; t < W; ++t)
            {
                flat = flat && c[t].colsLeft >= SPEC_STEPS;
                pre[t].node = -1;
            }
            if (!flat)
                for (int t = 0; t < W; ++t)
                    seed_prefetch<R, W>(pre[t], c[t], g, entry.data(), t, seedS.data(), SPEC_STEPS);
        }
        for (int t = 0; t < W; ++t) // events read what lane t-1 wrote at an EARLIER step only
            if (c[t].colsLeft == 0)
            {
                if (lean)
                    node_event_pre<R, W>(s[t], c[t], g, entry.data(), pre[t], t, seedS.data());
                else
                    node_event<R, true, W>(s[t], c[t], g, t, seedS.data(), L);
            }
            else
                --c[t].colsLeft;
        uint32_t rh[W], rf[W];
        for (int t = 0; t < W; ++t)
        {
            rh[t] = t ? s[t - 1].hbotLast : 0;
            rf[t] = t ? s[t - 1].foutLast : pk(-1, -1);
        }
        for (int t = 0; t < W; ++t)
        {
            const int code = g.codes[k - t];
            const ProfPtr<W> pf = { prof.data() + (code * R) * W + t };
            uint32_t tg[R];
            const uint32_t m = lane_step_pf<R, false, PG_LAZY_F != 0>(s[t], rh[t], rf[t], pf, nullptr, nullptr, nullptr, tg);
            track_max(c[t], m, k);
            pending[0] = std::max(pending[0], lo16(m) + MBIAS);
            pending[1] = std::max(pending[1], hi16(m) + MBIAS);
            if (WIDE)
                track_region<R>(c[t], m, tg, g, L, t);
        }
    }
    if (save_trace)
        last = seedS; // the kernel copies its shared-memory seed table to HBM at the end of a forward-graph task
    finalize_task<R, W>(seedS.data(), g.n_nodes, out);
}

template <int R, int W>
void emu_tile(const GraphView& g, const std::vector<uint32_t>& prof, const std::vector<uint32_t>& ckpt,
              std::vector<uint32_t>& last, int T, int blo, uint32_t* dst, int half, int kend)
{
    // the trace kernel recomputes a tile only up to the step the walk asked for (it never moves to a later one): the
    // rest of the slot is poisoned here, so that a walk that did look there would not go unnoticed
    for (int x = kend * TileGeom<R>::BAND_ROWS; x < TileGeom<R>::SLOT_WORDS; ++x)
        dst[x] = 0xffffffffu;
    Lane<R> s[W];
    LaneCtl c[W];
    const int ck = T * TS / CK; // the checkpoint at or before the tile's first step; the steps in between are not kept
    for (int t = 0; t < W; ++t)
    {
        ckpt_load<R, W>(s[t], ckpt.data() + (size_t)ck * Sizes<R, W>::CKW * W, t);
        ctl_at_step(c[t], g, ck * CK, t);
    }
    for (int kk = ck * CK - T * TS; kk < kend; ++kk)
    {
        const int k = T * TS + kk;
        for (int t = 0; t < W; ++t)
            if (c[t].colsLeft == 0)
                node_event<R, false, W>(s[t], c[t], g, t, last.data());
            else
                --c[t].colsLeft;
        uint32_t rh[W], rf[W];
        for (int t = 0; t < W; ++t)
        {
            rh[t] = t ? s[t - 1].hbotLast : 0;
            rf[t] = t ? s[t - 1].foutLast : pk(-1, -1);
        }
        for (int t = 0; t < W; ++t)
        {
            uint32_t Hc[R], Ec[R], Fc[R];
            if (kk < 0)
            {
                lane_step<R, false, W>(s[t], rh[t], rf[t], prof.data(), g.codes[k - t], t, nullptr, nullptr, nullptr);
                continue;
            }
            lane_step<R, true, W>(s[t], rh[t], rf[t], prof.data(), g.codes[k - t], t, Hc, Ec, Fc);
            tile_store<R, Sizes<R, W>::WIDE>(dst + (size_t)kk * TileGeom<R>::BAND_ROWS, t, blo, Hc, Ec, Fc, half);
        }
    }
}

template <int R, int W>
int emu_align_one(const SiteDev& sd, const uint8_t* bytes, const int32_t* ints, const uint8_t* bases, int L,
                  unsigned flags, Record& rec, std::vector<uint32_t>& ops_out, int* n_tiles)
{
    const GraphView g0 = make_view(sd, bytes, ints, 0), g1 = make_view(sd, bytes, ints, 1);
    std::vector<uint32_t> info0, info1, last, ckpt, dummy1, dummy2;
    TaskOut fw, rv;
    memset(&rv, 0, sizeof rv);
    emu_fill<R, W>(g0, bases, L, 0, true, info0, last, ckpt, fw);
    if (flags & AF_REVERSE_GRAPH)
        emu_fill<R, W>(g1, bases, L, 1, false, info1, dummy1, dummy2, rv);
    Decision d = decide_strand(fw, rv, flags);
    {
        // the kernels fill only the reversed-graph halves rev_plan asks for (paired with other reads' halves): replay
        // that here -- reveal half by half -- and demand the same decision as with both halves known
        int known[2] = { -1, -1 }, rounds = 0;
        for (int h = rev_plan(fw, known, flags); h >= 0; h = rev_plan(fw, known, flags))
        {
            known[h] = rv.n_top[h];
            if (++rounds > 2)
                return -2;
        }
        const Decision dl = decide_with(fw, known[0] >= 0 ? known[0] : 0, known[1] >= 0 ? known[1] : 0, flags);
        if (dl.half != d.half || dl.unique != d.unique || dl.score != d.score)
            return -2;
        g_rev_rounds += rounds;
        ++g_rev_reads;
    }
    std::vector<uint32_t> prof((size_t)NCODE * R * W);
    for (int t = 0; t < W; ++t)
        build_profile<R, W>(prof.data(), bases, L, 0, t);
    std::vector<uint32_t> tiles((size_t)2 * TileGeom<R>::SLOT_WORDS, 0);
    TileBuf<R> tb;
    tb.mem = tiles.data();
    tb.tile0 = tb.tile1 = -1;
    tb.blo0 = tb.blo1 = 0;
    tb.lru = 0;
    Walker w;
    memset(&w, 0, sizeof w);
    std::vector<uint32_t> oplog((size_t)2 * L + 64);
    const uint8_t* chars = bytes + sd.chars_off;
    int guard = 0;
    while (!walk<R, W>(w, tb, g0, chars, last.data(), bases, L, d.half, fw, oplog.data(), (int)oplog.size(), 0, 0u))
    {
        const int T = w.need_step / TS;
        int blo;
        const int slot = tb.admit(T, w.need_row, blo);
        emu_tile<R, W>(g0, prof, ckpt, last, T, blo, tiles.data() + (size_t)slot * TileGeom<R>::SLOT_WORDS, d.half,
                       tile_steps_needed(w.need_step, T));
        if (n_tiles)
            ++*n_tiles;
        if (++guard > 100000)
            return -1;
    }
    rec.graph_pos = w.position;
    rec.score = d.score;
    rec.unique = (uint8_t)d.unique;
    rec.chose_reverse = (uint8_t)d.half;
    rec.status = (uint8_t)w.status;
    rec.query_clipped = (uint16_t)w.clipped;
    // replay the op log back to front, merging runs of equal (node, op)  (gssw_cigar_push_back/_front merging)
    const int n = w.nops < (int)oplog.size() ? w.nops : (int)oplog.size();
    ops_out.assign((size_t)n + 1, 0);
    ops_out.resize((size_t)emit_cigar(oplog.data(), n, ops_out.data(), n + 1));
    rec.cigar_off = 0;
    rec.cigar_len = (flags & AF_CIGAR) ? (uint32_t)ops_out.size() : 0;
    return 0;
}

} // namespace

namespace
{
// the k-mer stage (pg_kmer.cuh) for one read, the group being a single lane; ops = the op words of a mapped read
KmerResult emu_kmer_one(const KmerView& v, int site, const uint8_t* bases, int L, int max_path_nodes, std::vector<uint32_t>& ops,
                        std::vector<uint8_t>& rv_out)
{
    const int ops_cap = L + 2 * max_path_nodes + 8;
    std::vector<uint32_t> km0((size_t)L + 1), km1((size_t)L + 1), bitmap(KMER_WIN / 32), ob((size_t)ops_cap), ot((size_t)ops_cap);
    std::vector<uint8_t> bytes((size_t)6 * (L + 1));
    std::vector<KmerCand> heap(KMER_MAX_PATHS + 2);
    KmerScratch sc;
    sc.km[0] = km0.data();
    sc.km[1] = km1.data();
    sc.bitmap = bitmap.data();
    sc.ops_best = ob.data();
    sc.ops_tmp = ot.data();
    sc.heap = heap.data();
    for (int x = 0; x < 2; ++x)
    {
        sc.seq[x] = bytes.data() + (size_t)x * (L + 1);
        sc.valid[x] = bytes.data() + (size_t)(2 + x) * (L + 1);
        sc.first[x] = bytes.data() + (size_t)(4 + x) * (L + 1);
    }
    sc.ops_cap = ops_cap;
    const KmerResult r = kmer_align_read(v, site, bases, L, 0, 1, sc);
    ops.assign(ob.begin(), ob.begin() + (r.status ? std::min(r.n_ops, ops_cap) : 0));
    rv_out.assign(sc.seq[1], sc.seq[1] + L);
    return r;
}

} // namespace

namespace
{
int g_geom_w = 32; // lanes per task used by the emulator (the kernels' default), see pgemu_set_geometry
}

extern "C" {

// W = 32, 16 or 8 lanes per task; rows per lane follow from the read length exactly like in the kernels' dispatch
void pgemu_set_geometry(int w) { g_geom_w = w; }

// Same calling convention as pgo_align_batch (oracle/pg_oracle.h), plus graph arguments.
// out6 = {graph_pos, score, unique, mapq, is_graph_reverse_strand, status}; returns 0 or negative.
int pgemu_align_batch(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom,
                      const int32_t* eto, int n_reads, const char* bases_blob, const int32_t* read_off,
                      const uint8_t* is_rev, unsigned flags, int32_t* out6, char* out_bases_blob, char* cigars,
                      int cigar_stride, int64_t* tiles_total)
{
    host::GraphStore gs;
    std::string err;
    if (gs.add(n_nodes, seq_blob, seq_off, n_edges, efrom, eto, err) < 0)
    {
        fprintf(stderr, "pgemu: %s\n", err.c_str());
        return -4;
    }
    int worst = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        const uint8_t* b = (const uint8_t*)bases_blob + read_off[i];
        const int L = read_off[i + 1] - read_off[i];
        if (L <= 0 || L > MAX_READ_LEN)
            return -3;
        Record rec;
        memset(&rec, 0, sizeof rec);
        std::vector<uint32_t> ops;
        int nt = 0;
        const SiteDev& sd0 = gs.sites[0];
        const uint8_t* gb = gs.bytes.data();
        const int32_t* gi = gs.ints.data();
        int rc;
        // reads that can leave gssw's 8-bit mode take a WIDE geometry (W = 32 only), like the kernels' dispatch
        if (L > BYTE_MAX_READ_LEN)
            rc = L <= 320 ? emu_align_one<10, 32>(sd0, gb, gi, b, L, flags, rec, ops, &nt)
                : (L <= 512 ? emu_align_one<16, 32>(sd0, gb, gi, b, L, flags, rec, ops, &nt)
                            : emu_align_one<32, 32>(sd0, gb, gi, b, L, flags, rec, ops, &nt));
        else if (g_geom_w == 32)
            rc = L <= 160 ? emu_align_one<5, 32>(sd0, gb, gi, b, L, flags, rec, ops, &nt)
                          : emu_align_one<8, 32>(sd0, gb, gi, b, L, flags, rec, ops, &nt);
        else if (g_geom_w == 16)
            rc = L <= 160 ? emu_align_one<10, 16>(sd0, gb, gi, b, L, flags, rec, ops, &nt)
                          : emu_align_one<16, 16>(sd0, gb, gi, b, L, flags, rec, ops, &nt);
        else
            rc = L <= 160 ? emu_align_one<20, 8>(sd0, gb, gi, b, L, flags, rec, ops, &nt)
                          : emu_align_one<32, 8>(sd0, gb, gi, b, L, flags, rec, ops, &nt);
        if (rc)
            worst = rc;
        if (tiles_total)
            *tiles_total += nt;
        int32_t* o = out6 + 6 * i;
        o[0] = rec.graph_pos;
        o[1] = rec.score;
        o[2] = rec.unique;
        o[3] = rec.unique ? 60 : 0;
        o[4] = ((is_rev ? is_rev[i] : 0) != 0) != (rec.chose_reverse != 0);
        o[5] = rec.status | ((int32_t)rec.query_clipped << 8);
        if (out_bases_blob)
            for (int j = 0; j < L; ++j)
                out_bases_blob[read_off[i] + j] = rec.chose_reverse ? (char)complement_base(b[L - 1 - j]) : (char)b[j];
        if (cigars)
        {
            std::string s = (flags & AF_CIGAR) ? host::format_cigar(rec, ops.data()) : std::string();
            size_t n = s.size() < (size_t)cigar_stride - 1 ? s.size() : (size_t)cigar_stride - 1;
            memcpy(cigars + (size_t)i * cigar_stride, s.data(), n);
            cigars[(size_t)i * cigar_stride + n] = 0;
        }
    }
    return worst;
}
// The exact-match stage (pg_path.cuh) on the host: same calling convention as pgo_path_align_batch (oracle/pg_oracle.h)
// plus graph arguments.  out8 = {mapped, graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, anchored}
int pgemu_path_align_batch(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom,
                           const int32_t* eto, int kmer_len, int n_reads, const char* bases_blob,
                           const int32_t* read_off, int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride,
                           int32_t* counters3)
{
    host::GraphStore gs;
    std::string err;
    if (gs.add(n_nodes, seq_blob, seq_off, n_edges, efrom, eto, err) < 0)
    {
        fprintf(stderr, "pgemu: %s\n", err.c_str());
        return -4;
    }
    host::PathIndexHost ix;
    host::build_path_index(gs, kmer_len, ix);
    const PathView v = make_path_view(ix.sites[0], gs.sites[0], ix.table.data(), ix.lists.data(), ix.succ.data(),
                                      gs.bytes.data(), gs.ints.data());
    counters3[0] = counters3[1] = counters3[2] = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        const uint8_t* b = (const uint8_t*)bases_blob + read_off[i];
        const int L = read_off[i + 1] - read_off[i];
        std::vector<uint8_t> q0((size_t)L + 1), q1((size_t)L + 1);
        path_strand_chars(b, L, 0, q0.data());
        path_strand_chars(b, L, 1, q1.data());
        PathResult rf, rr, r;
        path_scan_strand(v, q0.data(), L, 0, rf);
        path_scan_strand(v, q1.data(), L, 1, rr);
        path_combine(rf, rr, r);
        int32_t* o = out8 + 8 * i;
        memset(o, 0, 8 * sizeof(int32_t));
        ++counters3[0];
        counters3[1] += r.n_matches > 0;
        o[7] = r.n_matches > 0;
        std::string cg;
        if (r.n_full > 0)
        {
            Record rec;
            path_record(r, L, rec);
            std::vector<uint32_t> ops((size_t)r.first.n_nodes);
            path_emit(v, r.strand ? q1.data() : q0.data(), L, r, ops.data());
            cg = host::format_cigar(rec, ops.data());
            o[0] = 1;
            o[1] = rec.graph_pos;
            o[2] = rec.score;
            o[3] = rec.unique;
            o[4] = rec.unique ? 60 : 0;
            o[5] = rec.chose_reverse; // PathAligner.cpp:127-135: not combined with the read's own strand
            o[6] = (int32_t)cg.size();
            ++counters3[2];
        }
        if (out_bases_blob)
            for (int j = 0; j < L; ++j)
                out_bases_blob[read_off[i] + j] = (char)path_read_char(b, L, (r.n_full > 0) ? r.strand : 0, j);
        if (cigars)
        {
            size_t n = cg.size() < (size_t)cigar_stride - 1 ? cg.size() : (size_t)cigar_stride - 1;
            memcpy(cigars + (size_t)i * cigar_stride, cg.data(), n);
            cigars[(size_t)i * cigar_stride + n] = 0;
        }
    }
    return 0;
}

// rev_plan (pg_core.cuh) exhaustively: for every combination of forward-graph results (n_top 0..2 per half, score order),
// reversed-graph results and flags, revealing only the halves the plan asks for must give the decision of the full
// evaluation, in at most two rounds.  Returns the number of disagreements; *cases = combinations tried, *halves =
// reversed-graph halves asked for in total.
int pgemu_rev_plan_check(long* cases, long* halves)
{
    int bad = 0;
    *cases = *halves = 0;
    const unsigned flag_sets[] = { 0xFFFFFFFFu, 1u, 3u, 5u, 7u, 4u, 6u, 2u, 0u };
    for (unsigned flags : flag_sets)
        for (int f0 = 0; f0 <= 2; ++f0)
            for (int f1 = 0; f1 <= 2; ++f1)
                for (int sc = 0; sc < 3; ++sc) // S0 < S1, S0 == S1, S0 > S1
                    for (int r0 = 0; r0 <= 2; ++r0)
                        for (int r1 = 0; r1 <= 2; ++r1)
                        {
                            TaskOut fw, rv;
                            memset(&fw, 0, sizeof fw);
                            memset(&rv, 0, sizeof rv);
                            fw.n_top[0] = f0;
                            fw.n_top[1] = f1;
                            fw.score[0] = 10 + (sc == 2);
                            fw.score[1] = 10 + (sc == 0);
                            rv.n_top[0] = r0;
                            rv.n_top[1] = r1;
                            if (!(flags & AF_REVERSE_GRAPH)) // the kernels do not fill the reversed graph at all then
                                rv.n_top[0] = rv.n_top[1] = 0;
                            const Decision full = decide_strand(fw, rv, flags);
                            int known[2] = { -1, -1 }, rounds = 0;
                            for (int h = rev_plan(fw, known, flags); h >= 0 && rounds < 3; h = rev_plan(fw, known, flags))
                            {
                                known[h] = rv.n_top[h];
                                ++rounds;
                            }
                            const Decision lazy = decide_with(fw, known[0] >= 0 ? known[0] : 0, known[1] >= 0 ? known[1] : 0, flags);
                            bad += rounds > 2 || lazy.half != full.half || lazy.unique != full.unique || lazy.score != full.score;
                            ++*cases;
                            *halves += rounds;
                        }
    return bad;
}

// Sizing of the device-side index build (pg_host.hpp count_kmer_paths) against an actual enumeration of the site's k-mer
// paths with the device kernel's own depth-first walk restated on the host.  out3 = {n_paths by the DP, list bound by
// the DP, n_paths enumerated}; returns the number of node-list entries the enumeration needs.
long long pgemu_count_kmer_paths(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges,
                                 const int32_t* efrom, const int32_t* eto, int kmer_len, long long* out3)
{
    host::GraphStore gs;
    std::string err;
    if (gs.add(n_nodes, seq_blob, seq_off, n_edges, efrom, eto, err) < 0)
        return -4;
    std::vector<std::vector<int32_t>> sv;
    std::vector<int32_t> csr;
    host::build_successors(gs, 0, sv, csr);
    int64_t n_paths = 0, list_ints = 0;
    host::count_kmer_paths(gs, 0, kmer_len, csr.data(), csr.data() + n_nodes + 1, n_paths, list_ints);
    const int32_t* node_len = gs.ints.data() + gs.sites[0].tab_off[0] + n_nodes;
    long long enumerated = 0, lists = 0;
    std::vector<int> nodes((size_t)kmer_len + 2), endp((size_t)kmer_len + 3), ext((size_t)kmer_len + 3), nxt((size_t)kmer_len + 3);
    for (int v0 = 0; v0 < n_nodes; ++v0)
        for (int pos = 0; pos < node_len[v0]; ++pos)
        {
            int depth = 1;
            nodes[0] = v0;
            endp[1] = pos;
            ext[1] = kmer_len - 1;
            nxt[1] = 0;
            while (depth > 0)
            {
                const int last = nodes[(size_t)depth - 1];
                const int room = node_len[last] - endp[(size_t)depth] - 1;
                if (ext[(size_t)depth] > room)
                {
                    if (nxt[(size_t)depth] >= (int)sv[(size_t)last].size())
                    {
                        --depth;
                        continue;
                    }
                    nodes[(size_t)depth] = sv[(size_t)last][(size_t)nxt[(size_t)depth]++];
                    endp[(size_t)depth + 1] = 0;
                    ext[(size_t)depth + 1] = ext[(size_t)depth] - room - 1;
                    nxt[(size_t)depth + 1] = 0;
                    ++depth;
                    continue;
                }
                ++enumerated;
                lists += depth;
                --depth;
            }
        }
    out3[0] = n_paths;
    out3[1] = list_ints;
    out3[2] = enumerated;
    return lists;
}

// Align + count ONE site on the host with the device code of pg_core.cuh / pg_count.cuh.
// support: n_reads x 16 bytes (pg_read_support); path_words / ops_out: capacity `cap` words each (the op arena is
// returned too so that a test can look at the CIGARs); node_counts[n_nodes], edge_counts[n_edges] (4 x u32 each);
// family_words: entries {site, n, mask_lo, mask_hi, n x 4 counts}.  Returns 0, -5 on a capacity problem.
int pgemu_count_site(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom,
                     const int32_t* eto, const uint64_t* edge_labels, int n_reads, const char* bases_blob,
                     const int32_t* read_off, const uint8_t* is_rev, const int32_t* fragment, int remove_nonuniq,
                     double bad_align_frac, int use_support_filters, int family_slots, void* support,
                     uint32_t* path_words, uint32_t* ops_out, int cap, int32_t* used, uint32_t* node_counts,
                     uint32_t* edge_counts, uint32_t* family_words, int family_cap, int32_t* family_used)
{
    host::GraphStore gs;
    std::string err;
    if (gs.add(n_nodes, seq_blob, seq_off, n_edges, efrom, eto, err) < 0)
        return -4;
    for (int e = 0; e < n_edges; ++e)
        gs.in_label[(size_t)e] = edge_labels ? edge_labels[e] : 0ull;
    const SiteDev& sd0 = gs.sites[0];
    const uint8_t* gb = gs.bytes.data();
    const int32_t* gi = gs.ints.data();
    std::vector<Record> recs((size_t)n_reads);
    std::vector<uint32_t> arena;
    for (int i = 0; i < n_reads; ++i)
    {
        const uint8_t* b = (const uint8_t*)bases_blob + read_off[i];
        const int L = read_off[i + 1] - read_off[i];
        std::vector<uint32_t> ops;
        int nt = 0;
        const int rc = L <= 160 ? emu_align_one<5, 32>(sd0, gb, gi, b, L, 0xFFFFFFFFu, recs[(size_t)i], ops, &nt)
                                : emu_align_one<8, 32>(sd0, gb, gi, b, L, 0xFFFFFFFFu, recs[(size_t)i], ops, &nt);
        if (rc)
            return rc;
        recs[(size_t)i].cigar_off = (uint32_t)arena.size();
        arena.insert(arena.end(), ops.begin(), ops.begin() + recs[(size_t)i].cigar_len);
    }
    if ((int)arena.size() > cap)
        return -5;
    *used = (int32_t)arena.size();
    memcpy(ops_out, arena.data(), arena.size() * sizeof(uint32_t));

    CountParams prm;
    prm.remove_nonuniq = remove_nonuniq;
    prm.use_support_filters = use_support_filters;
    prm.bad_align_frac = bad_align_frac;
    prm.family_slots = family_slots > 0 ? family_slots : 256;
    host::CountHostTables ht;
    host::build_count_tables(gs, prm.family_slots, ht);
    CountTables t;
    t.sites = gs.sites.data();
    t.gints = gi;
    t.csite = ht.csite.data();
    t.csr_input = ht.csr_input.data();
    t.lab_edge = ht.lab_edge.data();
    t.lab_out = ht.lab_out.data();
    t.lab_in = ht.lab_in.data();
    ReadSupport* sup = static_cast<ReadSupport*>(support);
    for (int i = 0; i < n_reads; ++i)
        support_read(recs[(size_t)i], arena.data(), read_off[i + 1] - read_off[i], 0, is_rev && is_rev[i], t, prm, sup[i],
                     path_words);
    std::vector<int32_t> next;
    std::vector<uint8_t> head;
    if (!host::build_fragment_chains(fragment, nullptr, n_reads, next, head, err))
        return -1;
    std::vector<Count4> nc((size_t)n_nodes), ec((size_t)n_edges + 1), fam((size_t)ht.fam_rows);
    std::vector<unsigned long long> keys((size_t)ht.fam_keys + 1, 0ull);
    memset(nc.data(), 0, nc.size() * sizeof(Count4));
    memset(ec.data(), 0, ec.size() * sizeof(Count4));
    memset(fam.data(), 0, fam.size() * sizeof(Count4));
    for (int i = 0; i < n_reads; ++i)
        if (head[(size_t)i]
            && !count_fragment(i, 0, next.data(), sup, path_words, t, prm, nc.data(), ec.data(), keys.data(), fam.data()))
            return -5;
    memcpy(node_counts, nc.data(), (size_t)n_nodes * sizeof(Count4));
    memcpy(edge_counts, ec.data(), (size_t)n_edges * sizeof(Count4));
    const int n = 1 + n_nodes + n_edges;
    int w = 0;
    for (int slot = 0; slot < ht.csite[0].slots; ++slot)
    {
        if (!keys[(size_t)slot])
            continue;
        if (w + 4 + 4 * n > family_cap)
            return -5;
        family_words[w] = 0;
        family_words[w + 1] = (uint32_t)n;
        family_words[w + 2] = (uint32_t)keys[(size_t)slot];
        family_words[w + 3] = (uint32_t)(keys[(size_t)slot] >> 32);
        memcpy(family_words + w + 4, fam.data() + (size_t)slot * n, (size_t)n * sizeof(Count4));
        w += 4 + 4 * n;
    }
    *family_used = w;
    return 0;
}

// grm::KmerAligner<k> through the device source of pg_kmer.cuh.  out8 per read = {status (0 unmapped, 1 mapped, 2 not
// unique), graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, 0}; out_bases = the read's bases after
// the stage (reverse complement when the best candidate is on the reverse strand)
int pgemu_kmer_align_batch(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges, const int32_t* efrom,
                           const int32_t* eto, int n_paths, const int32_t* path_ptr, const int32_t* path_nodes, int k,
                           int n_reads, const char* bases_blob, const int32_t* read_off, const uint8_t* is_rev, int32_t* out8,
                           char* out_bases_blob, char* cigars, int cigar_stride)
{
    host::GraphStore gs;
    std::string err;
    if (gs.add(n_nodes, seq_blob, seq_off, n_edges, efrom, eto, err) < 0 || !gs.set_paths(0, n_paths, path_ptr, path_nodes, err))
        return -1;
    if (n_paths > KMER_MAX_PATHS)
        return -3;
    host::KmerIndexHost ix;
    host::build_kmer_index(gs, k, ix);
    KmerView v;
    v.sites = ix.sites.data();
    v.paths = ix.paths.data();
    v.seqs = ix.seqs.data();
    v.nodes = ix.nodes.data();
    v.kmers = ix.kmers.data();
    v.k = k;
    for (int i = 0; i < n_reads; ++i)
    {
        const uint8_t* b = (const uint8_t*)bases_blob + read_off[i];
        const int L = read_off[i + 1] - read_off[i];
        std::vector<uint32_t> ops;
        std::vector<uint8_t> rv;
        const KmerResult r = emu_kmer_one(v, 0, b, L, ix.max_path_nodes, ops, rv);
        int32_t* o = out8 + 8 * i;
        memset(o, 0, 8 * sizeof(int32_t));
        o[0] = r.status;
        if (out_bases_blob)
            memcpy(out_bases_blob + read_off[i], (r.status && r.rev) ? rv.data() : b, (size_t)L);
        if (cigars)
            cigars[(size_t)i * cigar_stride] = 0;
        if (!r.status)
            continue;
        Record rec;
        memset(&rec, 0, sizeof rec);
        rec.cigar_len = (uint32_t)ops.size();
        const std::string cs = host::format_cigar(rec, ops.data());
        o[1] = r.pos;
        o[2] = r.score;
        o[3] = r.status == 1;
        o[4] = r.status == 1 ? 60 : 0;
        o[5] = r.rev ? !(is_rev && is_rev[i]) : (is_rev && is_rev[i]);
        o[6] = (int32_t)cs.size();
        if (cigars && cigar_stride > 0)
        {
            const size_t m = std::min((size_t)cigar_stride - 1, cs.size());
            memcpy(cigars + (size_t)i * cigar_stride, cs.data(), m);
            cigars[(size_t)i * cigar_stride + m] = 0;
        }
    }
    return 0;
}
}
