// paragraph_b200 -- sm_100a kernels and the C-ABI (include/pg_align.h) on top of them.
//
//   pg_fill_kernel<R>  : one warp per (read, graph orientation) task; packed-int16 wavefront over all graph
//                        columns (pg_core.cuh); writes node maxima, and for forward-graph tasks the node last
//                        columns and a lane-state checkpoint every CK steps.  Replaces gssw_graph_fill +
//                        alignsEndAtMultNodes (gssw.c:3964-4028, GraphAligner.cpp:170-212).
//   pg_trace_kernel<R> : one warp per read; strand choice (GraphAligner.cpp:340-356), then the reference's
//                        traceback (gssw.c:1112-1818, 2621-3537) over tiles recomputed from the checkpoints.
//
// There is deliberately no CPU path in this file: without a CUDA device every entry point returns PG_E_CUDA.
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/pg_align.h"
#include "pg_core.cuh"
#include "pg_count.cuh"
#include "pg_kmer.cuh"
#include "pg_host.hpp"

using namespace pg;

static_assert(sizeof(pg_record) == sizeof(Record), "pg_record layout");

namespace
{

#ifndef PG_FILL_WARPS
#define PG_FILL_WARPS 4
#endif
#ifndef PG_TRACE_WARPS
#define PG_TRACE_WARPS 1
#endif
#ifndef PG_FILL_UNROLL
#define PG_FILL_UNROLL 4
#endif
#ifndef PG_SPEC_DEAD
#define PG_SPEC_DEAD 1 // speculative "no gap alive" blocks in the fill kernel.  Measured on the B200
                       // (profiles/r02a_*): fill 1.43 -> 1.27 ms on config 2, 33.1 -> 25.8 ms on config 5, bit-identical
                       // (result digests, the -m gpu suite, GPU fuzz).  0 = the plain kernel (A/B)
#endif
#ifndef PG_SPEC_PRUNE
#define PG_SPEC_PRUNE 0 // 1 (with PG_SPEC_DEAD): plus upper-bound pruning of gaps that cannot reach the best score so far
                        // (pg_core.cuh: gap_relevant; emulator-verified experiment)
#endif
#ifndef PG_FAST_UNROLL
#define PG_FAST_UNROLL 8
#endif
#ifndef PG_STATIC_BAR
#define PG_STATIC_BAR 1 // the staging mbarriers live in a static shared array: compute-sanitizer's synccheck only follows
                        // barriers it can see declared (in dynamic shared memory it reports "Missing init" on every wait
                        // and aborts the kernel; profiles/r02e_synccheck_variants.txt)
#endif
#ifndef PG_DEAD_BOUNDARY
#define PG_DEAD_BOUNDARY 0 // 1: node-boundary sub-blocks may run with the collapsed recurrence too (needs PG_LEAN_EVENTS).  Exact
                           // (GPU suite, fuzz) and measured slower on every shape, twice: off
#endif
#ifndef PG_DEADB_UNROLL
#define PG_DEADB_UNROLL 2
#endif
#ifndef PG_LEAN_EVENTS
#define PG_LEAN_EVENTS 1 // entry words + prefetched seeds at node boundaries (pg_core.cuh: node_event_pre); 0 = A/B
#endif
#ifndef PG_FAST_BLOCKS
#define PG_FAST_BLOCKS 1
#endif
constexpr int FILL_UNROLL = PG_FILL_UNROLL;
constexpr int DEADB_UNROLL = PG_DEADB_UNROLL; // unroll of the collapsed steps of a boundary sub-block
constexpr int FAST_UNROLL = PG_FAST_UNROLL;   // unroll of the boundary-free block of the fill kernel
constexpr bool FAST_BLOCKS = PG_FAST_BLOCKS != 0;
constexpr bool FILL_LAZY_F = PG_LAZY_F != 0;
constexpr int FILL_WARPS = PG_FILL_WARPS;   // warps per CTA, fill kernel
constexpr int TRACE_WARPS = PG_TRACE_WARPS; // warps per CTA, traceback kernel
constexpr unsigned FULL = 0xffffffffu;

struct FillArgs
{
    const SiteDev* sites;
    const uint8_t* gbytes;
    const int32_t* gints;
    const uint8_t* bases;
    const int32_t* read_off;
    const int32_t* read_site; // may be null
    int read0;                // first read of this chunk
    int n_tasks;              // 2 * reads in chunk
    unsigned flags;
    uint32_t* last; // [read in chunk][stride_last]
    uint32_t* ckpt; // [read in chunk][stride_ckpt]
    size_t stride_last, stride_ckpt;
    int n_nodes_cap; // seed / info table capacity per task (nodes)
    int code_smem_bytes; // per-task shared-memory room for the staged column codes (0 = read them from global/L1)
    int tab_ints_cap;    // per-task shared-memory room (words) for the staged node tables (STAGED only)
    uint32_t* tabG;  // seed + info tables in HBM when they do not fit shared memory (graphs with very many nodes), else null
    size_t stride_tab;
    TaskOut* tout; // [2 * n_reads] (global task index)
    int smem_words_per_task;
    // exact-match stage in front (pg_path.cuh): the reads still to align, compacted; null = all reads of the chunk
    const int32_t* todo;
    const int32_t* n_todo;
    // task list of the kernel: MODE_BOTH = (read, orientation) pairs, 2 per read (what GraphAligner::alignRead does);
    // MODE_FWD = the forward-graph task of every read; MODE_PAIRS = reversed-graph tasks whose two packed halves are the
    // still-needed reversed-graph fills (rev_plan, pg_core.cuh) of two reads of one site: rtasks[i] = {read << 1 | half,
    // read << 1 | half or -1}; their only result is n_top -> rv_ntop[read * 2 + half]
    int mode;
    const int2* rtasks;
    const int32_t* n_rtasks;
    int32_t* rv_ntop;
    int strided; // the grid is smaller than the task list: every warp strides over it (persistent CTAs)
};
enum { MODE_BOTH = 0, MODE_FWD = 1, MODE_PAIRS = 2 };

// ---- TMA (bulk async copy) staging of a task's column codes into shared memory ---------------------------------
// Canonical mbarrier protocol (CUDA programming guide, "Using TMA to transfer one-dimensional arrays"), at group
// scope: the elected lane initialises the barrier for `nlanes` arrivals and fences the init towards the async
// proxy; after a __syncwarp it posts the expected byte count (its own arrival) and issues cp.async.bulk
// (SASS: UBLKCP.S.G); the other lanes arrive; everybody waits for the phase (a persistent warp re-uses its barrier: the phase parity alternates per task).  Source and size are 16-byte aligned by
// construction (pg_host.hpp).
__device__ __forceinline__ void tma_barrier_init(uint64_t* bar, int nlanes, bool elected)
{
    if (elected)
    {
        const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar_s), "r"(nlanes) : "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}
__device__ __forceinline__ void tma_stage_codes(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                                bool elected)
{
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
    if (elected)
    {
        const uint32_t dst_s = (uint32_t)__cvta_generic_to_shared(dst_smem);
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_s), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_s),
                     "l"(src_gmem), "r"(bytes), "r"(bar_s)
                     : "memory");
    }
    else
        asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar_s) : "memory");
}
__device__ __forceinline__ void tma_wait(uint64_t* bar, uint32_t parity)
{
    const uint32_t bar_s = (uint32_t)__cvta_generic_to_shared(bar);
    uint32_t done = 0;
    while (!done)
        asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                     : "=r"(done)
                     : "r"(bar_s), "r"(parity)
                     : "memory");
}

// Profile row accessor of the fill kernel: 32-bit shared-window address of (code row block, this lane); the R loads
// of a step are LDS with immediate offsets.
template <int W> struct ProfSmem
{
    uint32_t a;
    __device__ __forceinline__ uint32_t operator()(int r) const
    {
        uint32_t v;
        asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a + (uint32_t)(r * W * 4)));
        return v;
    }
};

// finalize_task (pg_core.cuh, the serial statement used by the emulator) as W-lane group reductions: every lane scans
// its own column of the [node][3][W] table, four min / max reductions per packed half give the global maximum, the
// first node holding it, whether a second node holds it, and the earliest cell (column, then lane) in the first.
template <int W> __device__ __forceinline__ uint32_t group_max2(uint32_t v) // packed maximum over the W lanes of a task
{
#pragma unroll
    for (int d = W / 2; d >= 1; d >>= 1)
        v = max2(v, __shfl_xor_sync(FULL, v, d));
    return v;
}
template <int W> __device__ __forceinline__ int group_min(int v)
{
#pragma unroll
    for (int d = W / 2; d > 0; d >>= 1)
        v = min(v, __shfl_xor_sync(FULL, v, d, W));
    return v;
}
template <int R, int W> __device__ __forceinline__ void finalize_task_group(const uint32_t* tab, int n_nodes, int gl, TaskOut& o)
{
    constexpr int INF = 0x7fffffff, V = Sizes<R, W>::SEEDV, RW = Sizes<R, W>::ROWW;
    // this lane's info words (node maximum, first step per half) of node n; rows are lane-major, so the three words of a
    // lane come with one 128-bit load when they share an aligned quad (R = 5: words 5..7) instead of strided 32-bit ones
    const uint32_t* row0 = tab + (size_t)gl * RW;
    auto info3 = [&](int n, uint32_t& m, uint32_t& f0, uint32_t& f1) {
        const uint32_t* r = row0 + (size_t)n * W * RW;
        if ((V & 3) == 1) // the quad that starts at V - 1 holds all three
        {
            const uint4 q = *reinterpret_cast<const uint4*>(r + V - 1);
            m = q.y;
            f0 = q.z;
            f1 = q.w;
        }
        else
        {
            m = r[V];
            f0 = r[V + 1];
            f1 = r[V + 2];
        }
    };
#pragma unroll
    for (int h = 0; h < 2; ++h)
    {
        int S = 0;
        for (int n = 0; n < n_nodes; ++n)
        {
            uint32_t m, f0, f1;
            info3(n, m, f0, f1);
            S = max(S, half16(m, h) + MBIAS);
        }
        S = -group_min<W>(-S);
        int first = INF, second = INF, key = INF;
        for (int n = n_nodes - 1; n >= 0; --n)
        {
            uint32_t m, f0, f1;
            info3(n, m, f0, f1);
            if (half16(m, h) + MBIAS == S)
                first = n;
        }
        const int mnode = group_min<W>(first);
        if (mnode != INF)
        {
            for (int n = n_nodes - 1; n >= mnode; --n)
            {
                uint32_t m, f0, f1;
                info3(n, m, f0, f1);
                if (half16(m, h) + MBIAS != S)
                    continue;
                if (n > mnode)
                    second = n;
                else
                    key = (((int)(h ? f1 : f0) - gl) << 5) | gl; // (column, lane): W <= 32
            }
        }
        second = group_min<W>(second);
        key = group_min<W>(key);
        o.score[h] = S;
        o.n_top[h] = (mnode == INF) ? 0 : (second == INF ? 1 : 2);
        if (S == 0 || mnode == INF)
        {
            // every real cell is 0: gssw leaves ref_end = -1 -> empty CIGAR, position 0 (gssw.c:2728-2732)
            o.n_top[h] = min(n_nodes, 2);
            o.max_node[h] = -1;
            o.end_step[h] = 0;
            o.end_lane[h] = 0;
        }
        else
        {
            o.max_node[h] = mnode;
            o.end_lane[h] = key & 31;
            o.end_step[h] = (key >> 5) + (key & 31);
        }
    }
}

// STAGED: the task's column codes are TMA-staged into shared memory (graphs up to 16 KB); otherwise they are read
// through L1 from global memory.  A template parameter so that the hot loop's loads have a static address space.
// TABG: the seed / node-maximum tables live in HBM (graphs with so many nodes that they do not fit shared memory).
template <int R, int W, bool STAGED, bool TABG>
#if defined(PG_FILL_MINB) // A/B: promise the compiler this many CTAs per SM for the short-read instantiations (register cap)
__global__ void __launch_bounds__(FILL_WARPS * 32, (R <= 8 && W == 32) ? PG_FILL_MINB : 1) pg_fill_kernel(const FillArgs a)
#else
__global__ void __launch_bounds__(FILL_WARPS * 32) pg_fill_kernel(const FillArgs a)
#endif
{
    extern __shared__ uint32_t smem[];
    constexpr int NT = 32 / W; // tasks per warp: a group of W lanes per task
    constexpr bool WIDE = Sizes<R, W>::WIDE; // reads that can score past a byte: wider checkpoints, region maxima
    constexpr int CKW = Sizes<R, W>::CKW, ROWW = Sizes<R, W>::ROWW;
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / W, gl = lane % W;
    const int wpc = blockDim.x >> 5; // warps per CTA: 4, fewer when the node tables of a many-node graph need the room
    int n_tasks = a.n_tasks;
    if (a.mode == MODE_PAIRS)
        n_tasks = min(n_tasks, *a.n_rtasks);
    else if (a.todo) // tasks = the reads the exact-match stage left over, in the order it listed them
        n_tasks = min(n_tasks, (a.mode == MODE_FWD ? 1 : 2) * max(*a.n_todo - a.read0, 0));
    // One task per group of W lanes.  A grid smaller than the task list (a.strided: persistent CTAs, every warp on its
    // own -- no warp waits for the slowest of its CTA before the next task starts) strides over it.
    uint32_t tma_phase = 0;
    for (int wbase = (blockIdx.x * wpc + wic) * NT;; wbase += gridDim.x * wpc * NT)
    {
    const int ltask = wbase + grp;
    if (wbase >= n_tasks)
        return; // whole warp beyond the work list
    // the task: orientation o; packed half x = string (o, hx) of read rdx (MODE_PAIRS: possibly two different reads)
    const int tcl = ltask < n_tasks ? ltask : ltask - grp; // inactive tail groups shadow the warp's first task (they only idle along)
    int o, rd, rd1, h0 = 0, h1 = 1, slot;
    if (a.mode == MODE_PAIRS)
    {
        const int2 pr = a.rtasks[tcl];
        o = 1;
        rd = pr.x >> 1;
        h0 = pr.x & 1;
        rd1 = pr.y >= 0 ? (pr.y >> 1) : -1;
        h1 = pr.y >= 0 ? (pr.y & 1) : 1;
        slot = 0;
    }
    else
    {
        const int per = a.mode == MODE_FWD ? 1 : 2;
        o = a.mode == MODE_FWD ? 0 : (tcl & 1);
        slot = tcl / per; // index of the read within the chunk: its scratch slot
        rd = a.todo ? a.todo[a.read0 + slot] : a.read0 + slot;
        rd1 = rd;
    }
    const bool active = ltask < n_tasks && !(o == 1 && !(a.flags & AF_REVERSE_GRAPH));
    uint32_t* prof = smem + ((size_t)wic * NT + grp) * a.smem_words_per_task;
    // node table [n_nodes_cap][W][ROWW] (pg_core.cuh: Sizes): warp-private shared memory, or (fallback) HBM
    uint32_t* seedS = TABG ? a.tabG + (size_t)(ltask < n_tasks ? ltask : 0) * a.stride_tab : prof + NCODE * R * W;
    TaskOut* to = a.tout + (size_t)rd * 2 + o;
    if (a.mode == MODE_BOTH && !active && ltask < n_tasks && gl == 0)
    {
        TaskOut z;
        memset(&z, 0, sizeof z);
        *to = z;
    }
    const SiteDev sd = a.sites[a.read_site ? a.read_site[rd] : 0];
    // STAGED: the orientation's node tables (lengths, predecessor lists) are copied next to the seed tables, so that
    // the node-boundary code of the hot loop reads shared memory with 32-bit addresses instead of chasing HBM pointers
    int32_t* tabS = reinterpret_cast<int32_t*>(seedS + a.n_nodes_cap * ROWW * W);
    if (STAGED)
        for (int x = gl; x < sd.tab_ints; x += W)
            tabS[x] = a.gints[sd.tab_off[o] + x];
    const GraphView g = STAGED ? make_view_at(sd, a.gbytes + sd.codes_off[o], tabS) : make_view(sd, a.gbytes, a.gints, o);
    // ... followed by the entry word of every node (pg_core.cuh: entry_word), what a lane reads when it crosses into it
    uint32_t* evS = reinterpret_cast<uint32_t*>(tabS + sd.tab_ints);
    constexpr bool LEAN_EVENTS = STAGED && !TABG && !WIDE && PG_LEAN_EVENTS;
    if (LEAN_EVENTS)
    {
        __syncwarp();
        for (int m = 1 + gl; m < sd.n_nodes; m += W)
            evS[m] = entry_word(g, m);
    }
    const uint8_t* bases = a.bases + a.read_off[rd];
    const int L = a.read_off[rd + 1] - a.read_off[rd];
    // node sequences (column codes) of this task's orientation: TMA-staged into shared memory when they fit
    uint8_t* code_s = reinterpret_cast<uint8_t*>(prof + a.smem_words_per_task) - a.code_smem_bytes;
#if PG_STATIC_BAR
    __shared__ __align__(8) uint64_t static_bars[FILL_WARPS * 4]; // one per task slot of the CTA
    uint64_t* bar = static_bars + wic * NT + grp;
#else
    uint64_t* bar = reinterpret_cast<uint64_t*>(code_s) - 1;
#endif
    const uint32_t span = (uint32_t)code_span_bytes(g.G);
    const bool staged = STAGED && active; // the host only picks STAGED when every graph of the batch fits
    if (staged && wbase < (int)gridDim.x * wpc * NT) // first task of this warp
        tma_barrier_init(bar, W, gl == 0);
    __syncwarp();
    if (staged)
        tma_stage_codes(code_s, g.codes - SENT, span, bar, gl == 0); // in flight while the profile is built
    if (active)
        build_profile_pair<R, W>(prof, bases, L, o, h0, rd1 >= 0 ? a.bases + a.read_off[rd1] : nullptr,
                                 rd1 >= 0 ? a.read_off[rd1 + 1] - a.read_off[rd1] : 0, o, h1, gl);
    else
        for (int x = gl; x < NCODE * R * W; x += W)
            prof[x] = pk(NEG, NEG);
    __syncwarp();
    if (staged)
        tma_wait(bar, tma_phase);
    tma_phase ^= 1u;

    const bool save = active && (o == 0);
    TaskOut t;
    // pass 0: speculative blocks with the cheap bookkeeping; pass 1 (rare): everything exact -- see dead_range_score
    PG_NOUNROLL
    for (int pass = 0; pass < 2; ++pass)
    {
    const bool precise = pass > 0;
    Lane<R> s;
    lane_zero(s);
    LaneCtl c;
    ctl_at_step(c, g, 0, gl);
    if (WIDE)
        region_begin(c, g, L, gl);
    if (!active)
        c.colsLeft = COLS_INF;
    uint32_t* last = a.last + (size_t)slot * a.stride_last;
    uint32_t* ckpt = a.ckpt + (size_t)slot * a.stride_ckpt;
    const uint8_t* codes = (STAGED ? code_s + SENT : g.codes) - gl;
    const int my_nck = active ? num_ckpt(g.G, W) : 0;
    int nck = my_nck;
    if (NT > 1) // groups of one warp may belong to different sites: run to the longest, the others idle on sentinels
        for (int d = W; d < 32; d <<= 1)
            nck = max(nck, __shfl_xor_sync(FULL, nck, d));
    // this lane's column of the profile as a shared-window address: one IMAD per step selects the code's rows
    const uint32_t NO_F = pk(-1, -1); // "no insertion running into lane 0": any F <= 0 will do, negative lets lazy-F skip
    uint32_t lmask, nof0;
    asm("mov.u32 %0, %1;" : "=r"(lmask) : "r"(gl ? 1u : 0u));
    asm("mov.u32 %0, %1;" : "=r"(nof0) : "r"(gl ? 0u : NO_F));
    const ProfSmem<W> pf0 = { (uint32_t)__cvta_generic_to_shared(prof + gl) };
    const uint32_t zero = (uint32_t)a.n_tasks >> 31; // 0, in a register neither compiler stage folds (see lane_step_dead)
    // PG_SPEC_PRUNE: best t of this lane so far (a lower bound of it: folded in at sub-block starts), and the read rows
    // left below this lane's first row, per half (an absent half never counts)
    uint32_t Mall = zero;
    const int L1p = (a.mode == MODE_PAIRS) ? (rd1 >= 0 ? a.read_off[rd1 + 1] - a.read_off[rd1] : 0) : L;
    const uint32_t rem0 = pk(L - 1 - R * gl, L1p - 1 - R * gl);
    for (int cki = 0; cki < nck; ++cki)
    {
        const bool live = (NT == 1) || cki < my_nck;
        if (save && live)
            ckpt_store<R, W>(s, ckpt + (size_t)cki * CKW * W, gl);
        const int kbase = cki * CK;
        const uint8_t* cp = codes + kbase; // per-lane pointer, immediate offsets inside the unrolled body
        // Node boundaries are rare (a few per lane and task) but testing for them costs every step a counter, a
        // branch with its reconvergence point, a warp barrier and the register shuffling of the merge.  So: when no
        // lane of the warp meets a boundary within this block of CK steps (warp-uniform test), run the block
        // straight-line.
        // Speculation on "no gap alive" (pg_core.cuh: lane_step_dead; PG_SPEC_DEAD builds): the boundary test is made
        // per sub-block of SPEC_STEPS steps; a boundary-free sub-block runs with the collapsed recurrence when no lane
        // of the warp holds a positive E / F at its start, and is redone with the full step when a t > gap_open
        // appeared in it.
        if (PG_SPEC_DEAD && !WIDE && FAST_BLOCKS)
        {
            PG_NOUNROLL
            for (int sb = 0; sb < CK; sb += SPEC_STEPS)
            {
                const uint8_t* cpb = cp + sb;
                if (!__all_sync(FULL, c.colsLeft >= SPEC_STEPS))
                {
                    // the lanes that enter a merging node within this sub-block fold its older predecessors' last
                    // columns now, together (pg_core.cuh: seed_prefetch) -- their events, one lane at a time, stay short
                    SeedPre<R> pre;
                    if (LEAN_EVENTS)
                    {
                        __syncwarp();
                        seed_prefetch<R, W>(pre, c, g, evS, gl, seedS, SPEC_STEPS);
                    }
                    // A boundary sub-block may run with the collapsed recurrence as well (PG_DEAD_BOUNDARY): no lane holds a
                    // live gap and no prefetched seed brings one.  The lane's events run between the collapsed steps; each
                    // folds the maximum of the steps since the last one into the node it closes, and an event that left
                    // the lane with a live E (a second crossing inside the sub-block merges through the general path,
                    // whose seeds were not looked at beforehand) fails the attempt like a t > gap_open does.  What a
                    // failed attempt has to put back: Hp, the two shuffle values and the node bookkeeping -- the E of a
                    // valid attempt stay <= 0 whatever an event did to them, a failed attempt's E are set to 0 (they were
                    // <= 0 when it began), the rows it stored are stored again by the full steps.
                    if (PG_DEAD_BOUNDARY && LEAN_EVENTS && !PG_SPEC_PRUNE && !precise
                        && !__any_sync(FULL, gaps_alive(s) || seed_live(pre)))
                    {
                        bool brk = false;
                        DeadSave<R> keep;
                        dead_save(s, keep);
                        const int keepNode = c.node, keepLeft = c.colsLeft, keepF0 = c.first[0], keepF1 = c.first[1];
                        const uint32_t keepM = c.Mnode;
                        uint32_t Mt = zero, Mall = zero;
#pragma unroll DEADB_UNROLL
                        for (int kk = 0; kk < SPEC_STEPS; ++kk)
                        {
                            __syncwarp();
                            if (c.colsLeft == 0)
                            {
                                c.Mnode = max2(c.Mnode, add2(Mt, pk(-MBIAS, -MBIAS)));
                                Mall = max2(Mall, Mt);
                                Mt = zero;
                                node_event_pre<R, W>(s, c, g, evS, pre, gl, seedS);
                                brk = brk || e_alive(s);
                            }
                            else
                                --c.colsLeft;
                            uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                            rh *= lmask;
                            const int code = live ? cpb[kk] : 5;
                            const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
                            Mt = lane_step_dead<R>(s, rh, pf, zero, Mt);
                        }
                        Mall = max2(Mall, Mt);
                        if (!__any_sync(FULL, brk || dead_block_broken(Mall)))
                        {
                            c.Mnode = max2(c.Mnode, add2(Mt, pk(-MBIAS, -MBIAS)));
                            continue;
                        }
                        dead_restore(s, keep);
#pragma unroll
                        for (int r = 0; r < R; ++r) // (they were <= 0 when the attempt began: 0 is as good as what they were)
                            s.E[r] = zero;
                        c.node = keepNode;
                        c.colsLeft = keepLeft;
                        c.first[0] = keepF0;
                        c.first[1] = keepF1;
                        c.Mnode = keepM;
                    }
#pragma unroll FILL_UNROLL
                    for (int kk = 0; kk < SPEC_STEPS; ++kk) // (the boundary-aware steps below)
                    {
                        __syncwarp();
                        if (c.colsLeft == 0)
                        {
                            if (LEAN_EVENTS)
                                node_event_pre<R, W>(s, c, g, evS, pre, gl, seedS);
                            else
                                node_event<R, true, W>(s, c, g, gl, seedS, L);
                        }
                        else
                            --c.colsLeft;
                        uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                        uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
                        rh *= lmask;
                        rf = rf * lmask + nof0;
                        const int code = live ? cpb[kk] : 5;
                        const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
                        uint32_t tg[R];
                        const uint32_t m = lane_step_pf<R, false, FILL_LAZY_F>(s, rh, rf, pf, nullptr, nullptr, nullptr, tg);
                        track_max(c, m, kbase + sb + kk);
                    }
                    continue;
                }
                {
                    bool full = precise || __any_sync(FULL, gaps_alive(s));
                    uint32_t Sb = 0u;
                    bool haveSb = false;
                    if (PG_SPEC_PRUNE)
                    {
                        Mall = max2(Mall, add2(c.Mnode, pk(MBIAS, MBIAS)));
                        if (full) // live gaps: all of them unable to reach the best score so far?
                        {
                            Sb = group_max2<W>(Mall);
                            haveSb = true;
                            if (!__any_sync(FULL, lane_gaps_relevant<R>(s, rem0, sub2(rem0, pk(R, R)), Sb)))
                            {
                                lane_gaps_drop<R>(s);
                                full = false;
                            }
                        }
                    }
                    if (!full)
                    {
                        DeadSave<R> keep;
                        dead_save(s, keep);
                        uint32_t Mt = zero;
#if PG_SPEC_PRUNE // the pruning experiment lets blocks with t > gap_open stay "dead": exact bookkeeping there
                        const int keepF0 = c.first[0], keepF1 = c.first[1];
                        uint32_t Mn = track_t_begin(c);
#endif
#pragma unroll
                        for (int kk = 0; kk < SPEC_STEPS; ++kk)
                        {
                            uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                            rh *= lmask;
                            const int code = live ? cpb[kk] : 5;
                            const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
#if PG_SPEC_PRUNE
                            const uint32_t mt = lane_step_dead<R>(s, rh, pf, zero);
                            Mt = max2(Mt, mt);
                            track_t(c, Mn, mt, kbase + sb + kk);
#else
                            Mt = lane_step_dead<R>(s, rh, pf, zero, Mt); // the block's maximum, nothing else
#endif
                        }
                        full = __any_sync(FULL, dead_block_broken(Mt));
                        if (PG_SPEC_PRUNE && full) // the gaps it would have opened: all droppable?
                        {
                            if (!haveSb)
                                Sb = group_max2<W>(Mall);
                            full = __any_sync(FULL, dead_block_broken_pruned(Mt, rem0, Sb));
                        }
#if PG_SPEC_PRUNE
                        if (full)
                        {
                            dead_restore(s, keep);
                            c.first[0] = keepF0;
                            c.first[1] = keepF1;
                        }
                        else
                            track_t_end(c, Mn);
#else
                        if (full)
                            dead_restore(s, keep);
                        else // fold the block into the node maximum (kept with the offset MBIAS); no first-reached steps
                            c.Mnode = max2(c.Mnode, add2(Mt, pk(-MBIAS, -MBIAS)));
#endif
                    }
                    if (full)
                    {
#pragma unroll
                        for (int kk = 0; kk < SPEC_STEPS; ++kk)
                        {
                            uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                            uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
                            rh *= lmask;
                            rf = rf * lmask + nof0;
                            const int code = live ? cpb[kk] : 5;
                            const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
                            uint32_t tg[R];
                            const uint32_t m = lane_step_pf<R, false, FILL_LAZY_F>(s, rh, rf, pf, nullptr, nullptr, nullptr, tg);
                            track_max(c, m, kbase + sb + kk);
                        }
                    }
                    c.colsLeft -= SPEC_STEPS;
                }
            }
            continue;
        }
        if (FAST_BLOCKS && __all_sync(FULL, c.colsLeft >= CK))
        {
            {
#pragma unroll FAST_UNROLL
                for (int kk = 0; kk < CK; ++kk)
                {
                    uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                    uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
                    rh *= lmask; // lane 0 has no lane above it: H = 0, no insertion running in.  (Multiply-add by an opaque
                    rf = rf * lmask + nof0; // 0/1 instead of a select: runs on the FMA pipe, the ALU pipe is the busy one)
                    const int code = live ? cp[kk] : 5;
                    const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
                    uint32_t tg[R];
                    const uint32_t m = lane_step_pf<R, false, FILL_LAZY_F>(s, rh, rf, pf, nullptr, nullptr, nullptr, tg);
                    track_max(c, m, kbase + kk);
                    if (WIDE)
                        track_region<R>(c, m, tg, g, L, gl);
                }
            }
            c.colsLeft -= CK;
            continue;
        }
#pragma unroll FILL_UNROLL
        for (int kk = 0; kk < CK; ++kk)
        {
            const int k = kbase + kk;
            __syncwarp();
            if (c.colsLeft == 0) // rare, per lane: node boundary
                node_event<R, true, W>(s, c, g, gl, seedS, L);
            else
                --c.colsLeft;
            uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
            uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
            if (gl == 0)
            {
                rh = 0;
                rf = NO_F;
            }
            const int code = live ? cp[kk] : 5;
            const ProfSmem<W> pf = { pf0.a + (uint32_t)code * (uint32_t)(R * W * 4) };
            uint32_t tg[R];
            const uint32_t m = lane_step_pf<R, false, FILL_LAZY_F>(s, rh, rf, pf, nullptr, nullptr, nullptr, tg);
            track_max(c, m, k);
            if (WIDE)
                track_region<R>(c, m, tg, g, L, gl);
        }
    }
    __syncwarp();
    if (save) // node last columns for the traceback kernel: one coalesced copy of the node table
        for (int x = gl; x < g.n_nodes * ROWW * W; x += W)
            last[x] = seedS[x];
    if (!WIDE)
        finalize_task_group<R, W>(seedS, g.n_nodes, gl, t); // every lane of the warp takes part (group-wide reductions)
    // a forward-graph fill whose top score lies in the range the speculative blocks do not keep first-reached steps
    // for: once more, exactly (the traceback starts at that cell)
    bool again = false;
    if (PG_SPEC_DEAD && !PG_SPEC_PRUNE && !WIDE && FAST_BLOCKS && !precise)
        again = __any_sync(FULL, save && (dead_range_score(t.score[0]) || dead_range_score(t.score[1])));
    if (!again)
        break;
    __syncwarp();
    } // pass
    if (active && gl == 0)
    {
        if (WIDE) // long reads: the serial statement, with the 16-bit-mode uniqueness rule (n_top_rule)
            finalize_task<R, W>(seedS, g.n_nodes, t);
        if (a.mode == MODE_PAIRS) // all a reversed-graph fill is for: does the top score sit in more than one node?
        {
            a.rv_ntop[rd * 2 + h0] = t.n_top[0];
            if (rd1 >= 0)
                a.rv_ntop[rd1 * 2 + h1] = t.n_top[1];
        }
        else
            *to = t;
    }
    if (!a.strided)
        return;
    __syncwarp(); // the next task overwrites this one's profile, tables and staged codes
    if (STAGED)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
}

struct TraceArgs
{
    const SiteDev* sites;
    const uint8_t* gbytes;
    const int32_t* gints;
    const uint8_t* bases;
    const int32_t* read_off;
    const int32_t* read_site;
    int read0, n_reads; // chunk
    unsigned flags;
    const uint32_t* last;
    const uint32_t* ckpt;
    size_t stride_last, stride_ckpt;
    const TaskOut* tout;
    Record* records;
    uint32_t* arena;
    unsigned long long* cursor;
    unsigned long long arena_cap;
    int smem_bytes_per_task;
    int oplog_cap;
    const int32_t* todo; // see FillArgs
    const int32_t* n_todo;
    const uint8_t* prerev; // see PathArgs; null without the exact-match stage
    const int32_t* rv_ntop; // MODE_PAIRS: n_top of the reversed-graph fills that were needed (-1 = not needed); else null
    // the reads whose strand decision still waits for a second-round reversed-graph fill are traced by a second, tiny
    // launch: pending_mode 1 = skip the reads flagged in `pending`, 2 = only those (0 / null: every read)
    const uint8_t* pending;
    int pending_mode;
};

template <int R, int W> __global__ void __launch_bounds__(TRACE_WARPS * 32) pg_trace_kernel(const TraceArgs a)
{
    extern __shared__ uint32_t smem[];
    constexpr int NT = 32 / W; // reads per warp: a group of W lanes per read
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = lane / W, gl = lane % W;
    const unsigned gmask = (W == 32) ? FULL : (((1u << (W & 31)) - 1u) << (grp * W));
    const int lrd0 = (blockIdx.x * TRACE_WARPS + wic) * NT;
    int n_reads = a.n_reads;
    if (a.todo)
        n_reads = min(n_reads, max(*a.n_todo - a.read0, 0));
    if (lrd0 >= n_reads)
        return;
    bool active = lrd0 + grp < n_reads;
    const int lrd = active ? lrd0 + grp : lrd0; // idle tail groups shadow the warp's first read (no stores)
    const int rd = a.todo ? a.todo[a.read0 + lrd] : a.read0 + lrd;
    if (a.pending_mode)
    {
        const bool mine = (a.pending[rd] != 0) == (a.pending_mode == 2);
        if (!__any_sync(FULL, active && mine))
            return; // nothing for this warp in this launch
        active = active && mine;
    }
    uint8_t* wmem = reinterpret_cast<uint8_t*>(smem) + ((size_t)wic * NT + grp) * a.smem_bytes_per_task;
    uint32_t* prof = reinterpret_cast<uint32_t*>(wmem);
    uint32_t* oplog = prof + NCODE * R * W;
    uint32_t* tiles = oplog + a.oplog_cap;

    const SiteDev sd = a.sites[a.read_site ? a.read_site[rd] : 0];
    const GraphView g = make_view(sd, a.gbytes, a.gints, 0);
    const uint8_t* chars = a.gbytes + sd.chars_off;
    const uint8_t* bases = a.bases + a.read_off[rd];
    const int L = a.read_off[rd + 1] - a.read_off[rd];
    const uint32_t* last = a.last + (size_t)lrd * a.stride_last;
    const uint32_t* ckpt = a.ckpt + (size_t)lrd * a.stride_ckpt;
    build_profile<R, W>(prof, bases, L, 0, gl);

    const TaskOut fw = a.tout[(size_t)rd * 2];
    TaskOut rv;
    if (a.rv_ntop) // an unknown half is one the strand rule provably does not look at (rev_plan): any value will do
    {
        rv.n_top[0] = max(a.rv_ntop[rd * 2], 0);
        rv.n_top[1] = max(a.rv_ntop[rd * 2 + 1], 0);
    }
    else
        rv = a.tout[(size_t)rd * 2 + 1];
    const Decision d = decide_strand(fw, rv, a.flags);
    const int half = d.half;

    TileBuf<R> tb;
    tb.mem = tiles;
    tb.tile0 = tb.tile1 = -1;
    tb.blo0 = tb.blo1 = 0;
    tb.lru = 0;
    Walker w;
    memset(&w, 0, sizeof w);
    __syncwarp();
    const uint8_t* codes = g.codes - gl;
    bool done = !active;
    for (int guard = 0; guard < (1 << 20); ++guard)
    {
        // Walk phase: group-uniform (every lane of a group carries the same walker state; diagonal runs are probed
        // W cells at a time with a ballot, the other moves are executed redundantly by the group's lanes).  Groups
        // of one warp diverge here and reconverge for the recomputation below.
        if (!done)
            done = walk<R, W>(w, tb, g, chars, last, bases, L, half, fw, oplog, a.oplog_cap, gl, gmask);
        if (__all_sync(FULL, done))
            break;
        // Recompute phase (warp-uniform): every unfinished group rebuilds the tile it missed from its checkpoint
        int T = 0, slot = 0, blo = 0;
        if (!done)
        {
            T = w.need_step / TS;
            slot = tb.admit(T, w.need_row, blo);
        }
        __syncwarp();
        Lane<R> s;
        LaneCtl c;
        const int ck = T * TS / CK; // the checkpoint at or before the tile's first step
        if (!done)
        {
            ckpt_load<R, W>(s, ckpt + (size_t)ck * Sizes<R, W>::CKW * W, gl);
            ctl_at_step(c, g, ck * CK, gl);
            // the walk moves towards smaller steps: the tile it misses next is almost always T - 1.  Its checkpoint is
            // in HBM (the fill kernel wrote hundreds of MB since); start fetching it now, the load above was the stall
            // of this kernel that the other warps hide least (long scoreboard at ck_unpack)
            if (T > 0 && (T - 1) * TS / CK != ck)
            {
                const uint32_t* nx = ckpt + (size_t)((T - 1) * TS / CK) * Sizes<R, W>::CKW * W + gl;
#pragma unroll
                for (int x = 0; x < Sizes<R, W>::CKW; ++x)
                    asm volatile("prefetch.global.L1 [%0];" ::"l"(nx + x * W));
            }
        }
        else
        {
            lane_zero(s);
            c.node = 0;
            c.colsLeft = COLS_INF;
            c.Mnode = 0;
            c.first[0] = c.first[1] = 0;
        }
        uint32_t* dst = tiles + (size_t)slot * TileGeom<R>::SLOT_WORDS;
        const uint8_t* cp = codes + T * TS;
        // the steps between the checkpoint and the tile (CK > TS): the same recurrence, cells not kept
        int pre = done ? 0 : T * TS - ck * CK;
        if (NT > 1)
            for (int dd = W; dd < 32; dd <<= 1)
                pre = max(pre, __shfl_xor_sync(FULL, pre, dd));
        if (CK > TS)
        {
            const int mine = done ? 0 : T * TS - ck * CK; // groups with a shorter run-up idle on sentinel columns first
#pragma unroll 2
            for (int kk = -pre; kk < 0; ++kk)
            {
                const bool on = kk >= -mine;
                if (on)
                {
                    if (c.colsLeft == 0)
                        node_event<R, false, W>(s, c, g, gl, const_cast<uint32_t*>(last));
                    else
                        --c.colsLeft;
                }
                uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
                uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
                if (gl == 0)
                {
                    rh = 0;
                    rf = 0;
                }
                if (on)
                    lane_step<R, false, W>(s, rh, rf, prof, cp[kk], gl, nullptr, nullptr, nullptr);
            }
        }
        // The walk only ever moves to smaller steps (left, up and diagonal moves all do; so does the jump into a
        // predecessor's last column), and what it asks for is its current cell or a neighbour of it, at most two steps
        // before the current cell (pg_core.cuh: walk): of the tile it missed it can touch the steps up to need_step + 2,
        // no later one (tile_steps_needed; the emulator poisons the rest).  That saves 40 % of the recomputation.
        int kend = done ? 0 : tile_steps_needed(w.need_step, T);
        if (NT > 1)
            for (int dd = W; dd < 32; dd <<= 1)
                kend = max(kend, __shfl_xor_sync(FULL, kend, dd));
#pragma unroll 2
        for (int kk = 0; kk < kend; ++kk)
        {
            if (c.colsLeft == 0)
                node_event<R, false, W>(s, c, g, gl, const_cast<uint32_t*>(last));
            else
                --c.colsLeft;
            uint32_t rh = __shfl_up_sync(FULL, s.hbotLast, 1, W);
            uint32_t rf = __shfl_up_sync(FULL, s.foutLast, 1, W);
            if (gl == 0)
            {
                rh = 0;
                rf = 0;
            }
            uint32_t Hc[R], Ec[R], Fc[R];
            lane_step<R, true, W>(s, rh, rf, prof, done ? 5 : cp[kk], gl, Hc, Ec, Fc);
            if (!done)
                tile_store<R, Sizes<R, W>::WIDE>(dst + (size_t)kk * TileGeom<R>::BAND_ROWS, gl, blo, Hc, Ec, Fc, half);
        }
        __syncwarp();
    }
    __syncwarp();
    if (active && gl == 0)
    {
        Record rec;
        rec.graph_pos = w.position;
        rec.score = (int16_t)d.score;
        rec.unique = (uint8_t)d.unique;
        rec.chose_reverse = (uint8_t)half;
        rec.status = (uint8_t)(w.phase == 2 ? w.status : 1);
        const int flips = a.prerev ? a.prerev[rd] : 0; // reverse complements the earlier stages applied to the bases
        rec.mapped_by = (uint8_t)(flips == 0 ? STAGE_GSSW : (flips == 1 ? STAGE_GSSW_REV : STAGE_GSSW_REV2));
        rec.query_clipped = (uint16_t)w.clipped;
        rec.cigar_off = 0;
        rec.cigar_len = 0;
        if (a.flags & AF_CIGAR)
        {
            const int n = w.nops < a.oplog_cap ? w.nops : a.oplog_cap;
            const int m = emit_cigar(oplog, n, nullptr, 0);
            const unsigned long long off = atomicAdd(a.cursor, (unsigned long long)m);
            if (off + (unsigned long long)m <= a.arena_cap)
            {
                emit_cigar(oplog, n, a.arena + off, m);
                rec.cigar_off = (uint32_t)off;
                rec.cigar_len = (uint32_t)m;
            }
            else
                rec.status = 2;
        }
        a.records[rd] = rec;
    }
}

// ---- which reversed-graph fills are needed, and pairing them ----------------------------------------------------
// pg_plan_kernel: one thread per read of the chunk: rev_plan (pg_core.cuh) on the forward-graph result and on what is
// known of the reversed-graph halves; a needed half is appended to the request list, warp-aggregated so that requests
// stay in read order within a warp (neighbouring reads mostly share their site).
// pg_pair_kernel: one thread per two consecutive requests: same site -> one task carrying both halves, else two tasks.
struct PlanArgs
{
    const TaskOut* tout;
    const int32_t* rv_ntop;
    const int32_t* read_site;
    const int32_t* todo;
    const int32_t* n_todo;
    int read0, n_reads;
    unsigned flags;
    int32_t* req;
    int32_t* n_req;
    int2* rtasks;
    int32_t* n_rtasks;
    uint8_t* pending; // may be null; [read] 1 = this round asks for another reversed-graph fill of the read
};
__global__ void __launch_bounds__(128) pg_plan_kernel(const PlanArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    int n = a.n_reads;
    if (a.todo)
        n = min(n, max(*a.n_todo - a.read0, 0));
    int want = -1, rd = 0;
    if (i < n)
    {
        rd = a.todo ? a.todo[a.read0 + i] : a.read0 + i;
        const int known[2] = { a.rv_ntop[rd * 2], a.rv_ntop[rd * 2 + 1] };
        want = rev_plan(a.tout[(size_t)rd * 2], known, a.flags);
        if (a.pending)
            a.pending[rd] = want >= 0 ? 1 : 0;
    }
    const unsigned m = __ballot_sync(FULL, want >= 0);
    if (!m)
        return;
    const int lane = threadIdx.x & 31;
    int base = 0;
    if (lane == __ffs((int)m) - 1)
        base = atomicAdd(a.n_req, __popc(m));
    base = __shfl_sync(FULL, base, __ffs((int)m) - 1);
    if (want >= 0)
        a.req[base + __popc(m & ((1u << lane) - 1u))] = rd * 2 + want;
}
__global__ void __launch_bounds__(128) pg_pair_kernel(const PlanArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *a.n_req;
    if (2 * i >= n)
        return;
    const int x = a.req[2 * i], y = 2 * i + 1 < n ? a.req[2 * i + 1] : -1;
    const bool same = y >= 0 && (!a.read_site || a.read_site[x >> 1] == a.read_site[y >> 1]);
    if (same || y < 0)
        a.rtasks[atomicAdd(a.n_rtasks, 1)] = make_int2(x, y);
    else
    {
        const int at = atomicAdd(a.n_rtasks, 2);
        a.rtasks[at] = make_int2(x, -1);
        a.rtasks[at + 1] = make_int2(y, -1);
    }
}

// ---- exact-match stage (grm::PathAligner, pg_path.cuh) ----------------------------------------------------------
struct PathArgs
{
    const SiteDev* sites;
    const uint8_t* gbytes;
    const int32_t* gints;
    const PathSite* psites;
    const PathEntry* ptable;
    const int32_t* plists;
    const int32_t* psucc;
    uint8_t* bases; // writable: a second-chance read is reverse-complemented in place like PathAligner.cpp:124-128 does
    const int32_t* read_off;
    const int32_t* read_site; // may be null
    int n_reads;
    int no_gssw; // no DP stage behind this one: unmapped reads get their (unmapped) record here
    int second_chance; // the caller's filter has NonUniq: a non-unique exact match goes on to the DP (see pg_set_stages)
    uint8_t* prerev;   // [n_reads] 1 = bases were reverse-complemented here before the DP saw them
    int row_stride;    // bytes per thread of staged strand characters
    Record* records;
    uint32_t* arena;
    unsigned long long* cursor;
    unsigned long long arena_cap;
    int32_t* todo;   // reads the stage did not map, for the DP kernels
    int32_t* n_todo;
    unsigned long long* counters; // attempted, anchored, mapped (PathAligner.hh:66-68)
};

// Two threads per read, one per strand (the reference scans forward, then reverse: PathAligner.cpp:90-109).  Each thread
// stages its strand's characters in shared memory (row stride = 4 x odd bytes: the threads of a warp walk their rows in
// lock step without bank conflicts), scans, and leaves its result in shared memory; the forward thread combines the
// two, writes the record / op words of a mapped read or appends the read to the DP's to-do list.
constexpr int PATH_THREADS = 128;
__global__ void __launch_bounds__(PATH_THREADS) pg_path_kernel(const PathArgs a)
{
    extern __shared__ uint32_t smem[];
    PathResult* res = reinterpret_cast<PathResult*>(smem);
    uint8_t* rows = reinterpret_cast<uint8_t*>(res + PATH_THREADS);
    const int tid = threadIdx.x, strand = tid & 1;
    const int rd = blockIdx.x * (PATH_THREADS / 2) + (tid >> 1);
    const bool valid = rd < a.n_reads;
    uint8_t* q = rows + (size_t)tid * a.row_stride;
    int L = 0;
    PathView v;
    uint8_t* bases = nullptr;
    if (valid)
    {
        const int site = a.read_site ? a.read_site[rd] : 0;
        v = make_path_view(a.psites[site], a.sites[site], a.ptable, a.plists, a.psucc, a.gbytes, a.gints);
        bases = a.bases + a.read_off[rd];
        L = a.read_off[rd + 1] - a.read_off[rd];
        path_strand_chars(bases, L, strand, q);
        path_scan_strand(v, q, L, strand, res[tid]);
    }
    __syncthreads();
    if (!valid || strand != 0)
        return;
    PathResult r;
    path_combine(res[tid], res[tid + 1], r);
    if (r.n_matches > 0)
        atomicAdd(a.counters + 1, 1ull);
    a.prerev[rd] = 0;
    if (r.n_full > 1 && a.second_chance && !a.no_gssw)
    {
        // MAPPED by this stage but not unique: the NonUniq filter turns it into BAD_ALIGN and the gssw stage gets the
        // read (CompositeAligner.cpp:97-103, 146-170) -- with the bases PathAligner left behind
        atomicAdd(a.counters + 2, 1ull);
        if (r.strand)
        {
            const uint8_t* q1 = q + a.row_stride; // the reverse-strand thread's row = reverseComplement(bases)
            for (int x = 0; x < L; ++x)
                bases[x] = q1[x];
            a.prerev[rd] = 1;
        }
        a.todo[atomicAdd(a.n_todo, 1)] = rd;
        return;
    }
    if (r.n_full > 0)
    {
        Record rec;
        path_record(r, L, rec);
        const unsigned long long off = atomicAdd(a.cursor, (unsigned long long)rec.cigar_len);
        if (off + rec.cigar_len <= a.arena_cap)
        {
            path_emit(v, q + (r.strand ? a.row_stride : 0), L, r, a.arena + off);
            rec.cigar_off = (uint32_t)off;
        }
        else
        {
            rec.status = 2;
            rec.cigar_len = 0;
        }
        a.records[rd] = rec;
        atomicAdd(a.counters + 2, 1ull);
        return;
    }
    a.todo[atomicAdd(a.n_todo, 1)] = rd;
    if (a.no_gssw) // no later stage: the read stays UNMAPPED (CompositeAligner.cpp:78-176)
    {
        Record rec;
        memset(&rec, 0, sizeof rec);
        rec.status = (uint8_t)ST_UNMAPPED;
        a.records[rd] = rec;
    }
}

// ---- the same stage, one WARP per read (default) -------------------------------------------------------------------
// pg_path_kernel above gives every strand to one thread: 20 k threads for a 10 k-read batch, about one warp per SM
// sub-partition and 7 of 32 lanes active (profiles/r01_path_kernel_ncu.txt).  Here a warp owns the read: both strands'
// characters sit in shared memory, the k-mer hashes of all positions are computed and probed by the 32 lanes in
// parallel (hit[p] = table slot of a verified unique k-mer, -1 otherwise), and the reference's sequential scan --
// first anchor at or after pos, greedy extension, resume one past the match (PathAligner.cpp:93-108) -- runs
// warp-uniformly with 32 characters per compare (ballot).  Same results as path_scan_strand / path_extend
// (pg_path.cuh), which stay the statement of the logic that the CPU emulator checks against the oracle.
__device__ __forceinline__ int lead_ones(unsigned m) { return m == FULL ? 32 : __ffs((int)~m) - 1; }

// common prefix of q[qi ..) and s[si ..), at most n characters
__device__ __forceinline__ int warp_prefix(const uint8_t* q, int qi, const uint8_t* s, int si, int n, int lane)
{
    int done = 0;
    while (done < n)
    {
        const int x = done + lane;
        const int lead = lead_ones(__ballot_sync(FULL, x < n && q[qi + x] == s[si + x]));
        done += lead;
        if (lead < 32)
            break;
    }
    return n <= 0 ? 0 : (done < n ? done : n);
}
// common suffix of q[.. qi) and s[.. si) (both exclusive ends), at most n characters
__device__ __forceinline__ int warp_suffix(const uint8_t* q, int qi, const uint8_t* s, int si, int n, int lane)
{
    int done = 0;
    while (done < n)
    {
        const int x = done + lane;
        const int lead = lead_ones(__ballot_sync(FULL, x < n && q[qi - 1 - x] == s[si - 1 - x]));
        done += lead;
        if (lead < 32)
            break;
    }
    return n <= 0 ? 0 : (done < n ? done : n);
}

// path_extend (pg_path.cuh) with every character loop done by the warp; all lanes hold the same state
__device__ void warp_path_extend(const PathView& v, const PathEntry& e, const uint8_t* q, int L, int pos, PathMatch& m,
                                 uint32_t* ops, const PathMatch* first_pass, int lane)
{
    int pos_in_query = pos + v.k;
    int node = v.lists[e.nodes_off + e.n_nodes - 1];
    int pos_in_node = e.end_pos + 1;
    int n_after = 0;
    const int base = first_pass ? first_pass->n_before : 0;
    if (ops && lane == 0)
        for (int x = 0; x < e.n_nodes; ++x)
            ops[base + x] = (uint32_t)v.lists[e.nodes_off + x] << 16;
    while (true) // ---- end extension (extendPathEndMatching, PathOperations.cpp:117-189)
    {
        const int nl = v.node_len[node];
        const int adv = warp_prefix(q, pos_in_query, v.raw + v.node_start[node], pos_in_node,
                                    min(L - pos_in_query, nl - pos_in_node), lane);
        pos_in_query += adv;
        pos_in_node += adv;
        if (pos_in_node < nl)
            break;
        int num_longest = 0, longest = 0, best = 0, min_size = 0x7fffffff;
        for (int x = v.succ_ptr[node]; x < v.succ_ptr[node + 1]; ++x)
            min_size = min(min_size, v.node_len[v.succ_idx[x]]);
        for (int x = v.succ_ptr[node]; x < v.succ_ptr[node + 1]; ++x)
        {
            const int c = v.succ_idx[x];
            const int p = warp_prefix(q, pos_in_query, v.raw + v.node_start[c], 0, min(min_size, L - pos_in_query), lane);
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
        if (ops && lane == 0)
            ops[base + e.n_nodes + n_after] = (uint32_t)best << 16;
        ++n_after;
        pos_in_query += longest;
        pos_in_node = longest;
        node = best;
    }
    const int end_pos = pos_in_node - 1, q_end = pos_in_query;
    pos_in_query = pos;
    node = v.lists[e.nodes_off];
    pos_in_node = e.start_pos;
    int n_before = 0;
    while (true) // ---- start extension (extendPathStartMatching, PathOperations.cpp:191-266)
    {
        const int adv = warp_suffix(q, pos_in_query, v.raw + v.node_start[node], pos_in_node, min(pos_in_query, pos_in_node), lane);
        pos_in_query -= adv;
        pos_in_node -= adv;
        if (pos_in_node > 0)
            break;
        int num_longest = 0, longest = 0, best = 0, min_size = 0x7fffffff;
        for (int x = v.pred_ptr[node]; x < v.pred_ptr[node + 1]; ++x)
            min_size = min(min_size, v.node_len[v.pred_idx[x]]);
        for (int x = v.pred_ptr[node]; x < v.pred_ptr[node + 1]; ++x)
        {
            const int c = v.pred_idx[x];
            const int cl = v.node_len[c];
            const int p = warp_suffix(q, pos_in_query, v.raw + v.node_start[c], cl, min(min_size, pos_in_query), lane);
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
        ++n_before;
        if (ops && lane == 0)
            ops[base - n_before] = (uint32_t)best << 16;
        pos_in_query -= longest;
        node = best;
        pos_in_node = v.node_len[node] - longest;
    }
    m.qpos = pos_in_query;
    m.plen = q_end - pos_in_query;
    m.start_node = node;
    m.start_pos = pos_in_node;
    m.n_before = n_before;
    m.n_nodes = n_before + e.n_nodes + n_after;
    if (ops)
    {
        __syncwarp();
        for (int x = lane; x < m.n_nodes; x += 32)
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
}

constexpr int PATHW_WARPS = 4;
__global__ void __launch_bounds__(PATHW_WARPS * 32) pg_path_warp_kernel(const PathArgs a)
{
    extern __shared__ uint32_t smem[];
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int rd = blockIdx.x * PATHW_WARPS + wic;
    if (rd >= a.n_reads)
        return;
    // per warp: hit[max_len] (int32), then the two strands' characters (row_stride bytes each)
    uint8_t* wmem = reinterpret_cast<uint8_t*>(smem) + (size_t)wic * (size_t)(a.row_stride * 6);
    int32_t* hit = reinterpret_cast<int32_t*>(wmem);
    uint8_t* q0 = wmem + 4 * (size_t)a.row_stride;
    uint8_t* q1 = q0 + a.row_stride;
    const int site = a.read_site ? a.read_site[rd] : 0;
    const PathView v = make_path_view(a.psites[site], a.sites[site], a.ptable, a.plists, a.psucc, a.gbytes, a.gints);
    uint8_t* bases = a.bases + a.read_off[rd];
    const int L = a.read_off[rd + 1] - a.read_off[rd];
    const int k = v.k;
    for (int j = lane; j < L; j += 32)
    {
        const uint8_t c = bases[j];
        q0[j] = c;
        q1[L - 1 - j] = complement_base(c);
    }
    __syncwarp();
    PathResult r;
    r.n_matches = r.n_full = 0;
    r.strand = r.seed_pos = 0;
    r.seed_hash = 0;
    int seed_slot = -1;
    if (k > 0 && L >= k)
        for (int strand = 0; strand < 2; ++strand)
        {
            const uint8_t* q = strand ? q1 : q0;
            // every lane: rolling hash over its contiguous share of the positions, first probe of each: hit[p] = slot
            // of an entry with the k-mer's key (a CANDIDATE: its characters are compared when the scan gets there),
            // -1 when the probe sequence ends at an empty slot
            const int npos = L - k + 1, per = (npos + 31) / 32;
            const uint64_t top = path_hash_pow(k);
            {
                const int p0 = lane * per, p1 = min(npos, p0 + per);
                uint64_t h = 0;
                if (p0 < p1)
                    for (int j = 0; j < k; ++j)
                        h = path_hash_step(h, q[p0 + j]);
                for (int p = p0; p < p1; ++p)
                {
                    const uint64_t hk = h ? h : 1;
                    const uint32_t lo = (uint32_t)hk, hi = (uint32_t)(hk >> 32);
                    int32_t cand = -1;
                    for (uint32_t slot = (uint32_t)(hk ^ (hk >> 29)) & (uint32_t)v.mask;; slot = (slot + 1) & (uint32_t)v.mask)
                    {
                        const PathEntry& e = v.table[slot];
                        if ((e.key_lo | e.key_hi) == 0u)
                            break;
                        if (e.n_nodes > 0 && e.key_lo == lo && e.key_hi == hi)
                        {
                            cand = (int32_t)slot;
                            break;
                        }
                    }
                    hit[p] = cand;
                    if (p + 1 < p1)
                        h = path_hash_step(h - ((uint64_t)q[p] + 1u) * top, q[p + k]);
                }
            }
            __syncwarp();
            int pos = 0;
            while (pos + k <= L)
            {
                int found = -1;
                for (int b0 = pos; b0 + k <= L && found < 0; b0 += 32)
                {
                    const int p = b0 + lane;
                    const unsigned mk = __ballot_sync(FULL, p + k <= L && hit[p] >= 0);
                    if (mk)
                        found = b0 + __ffs((int)mk) - 1;
                }
                if (found < 0)
                    break;
                // verify the candidate's characters along its node list (32 at a time); on a hash collision go on
                // probing like path_lookup does
                int slot = hit[found];
                bool ok = false;
                while (true)
                {
                    const PathEntry& e = v.table[slot];
                    if ((e.key_lo | e.key_hi) == 0u)
                        break;
                    if (e.n_nodes > 0 && e.key_lo == v.table[hit[found]].key_lo && e.key_hi == v.table[hit[found]].key_hi)
                    {
                        int j = found;
                        ok = true;
                        for (int x = 0; x < e.n_nodes && ok; ++x)
                        {
                            const int nd = v.lists[e.nodes_off + x];
                            const int a0 = x == 0 ? e.start_pos : 0, b1 = x == e.n_nodes - 1 ? e.end_pos : v.node_len[nd] - 1;
                            const int n = b1 - a0 + 1;
                            ok = warp_prefix(q, j, v.raw + v.node_start[nd], a0, n, lane) == n;
                            j += n;
                        }
                        if (ok)
                            break;
                    }
                    slot = (slot + 1) & v.mask;
                }
                if (!ok)
                {
                    pos = found + 1;
                    continue;
                }
                __syncwarp();
                if (lane == 0)
                    hit[found] = slot;
                __syncwarp();
                PathMatch m;
                warp_path_extend(v, v.table[hit[found]], q, L, found, m, nullptr, nullptr, lane);
                ++r.n_matches;
                if (m.plen == L)
                {
                    if (r.n_full == 0)
                    {
                        r.strand = strand;
                        r.seed_pos = found;
                        r.first = m;
                        seed_slot = hit[found];
                    }
                    ++r.n_full;
                }
                pos = m.qpos + m.plen + 1; // PathAligner.cpp:106 and the loop's ++pos
            }
            __syncwarp();
        }
    if (lane == 0)
    {
        if (r.n_matches > 0)
            atomicAdd(a.counters + 1, 1ull);
        a.prerev[rd] = 0;
    }
    if (r.n_full > 1 && a.second_chance && !a.no_gssw)
    {
        // MAPPED by this stage but not unique: the NonUniq filter turns it into BAD_ALIGN and the gssw stage gets the
        // read (CompositeAligner.cpp:97-103, 146-170) -- with the bases PathAligner left behind
        if (r.strand)
            for (int x = lane; x < L; x += 32)
                bases[x] = q1[x];
        if (lane == 0)
        {
            atomicAdd(a.counters + 2, 1ull);
            a.prerev[rd] = (uint8_t)(r.strand ? 1 : 0);
            a.todo[atomicAdd(a.n_todo, 1)] = rd;
        }
        return;
    }
    if (r.n_full > 0)
    {
        Record rec;
        path_record(r, L, rec);
        unsigned long long off = 0;
        if (lane == 0)
            off = atomicAdd(a.cursor, (unsigned long long)rec.cigar_len);
        off = __shfl_sync(FULL, off, 0);
        if (off + rec.cigar_len <= a.arena_cap)
        {
            PathMatch m2;
            warp_path_extend(v, v.table[seed_slot], r.strand ? q1 : q0, L, r.seed_pos, m2, a.arena + off, &r.first, lane);
            rec.cigar_off = (uint32_t)off;
        }
        else
        {
            rec.status = 2;
            rec.cigar_len = 0;
        }
        if (lane == 0)
        {
            a.records[rd] = rec;
            atomicAdd(a.counters + 2, 1ull);
        }
        return;
    }
    if (lane == 0)
    {
        a.todo[atomicAdd(a.n_todo, 1)] = rd;
        if (a.no_gssw) // no later stage: the read stays UNMAPPED (CompositeAligner.cpp:78-176)
        {
            Record rec;
            memset(&rec, 0, sizeof rec);
            rec.status = (uint8_t)ST_UNMAPPED;
            a.records[rd] = rec;
        }
    }
}

// ---- the k-mer stage (grm::KmerAligner<K>, pg_kmer.cuh): one warp per read ------------------------------------------
// Work list: the reads the exact-match stage left over (todo_in), or all reads.  A uniquely mapped read gets its record
// and op words here; the others are listed in todo_out for the DP -- a read that mapped but not uniquely goes there
// with the bases KmerAligner leaves behind (reverse-complemented when its best candidate was on the reverse strand,
// KmerAligner.cpp:452-460) and one more flip counted in prerev.
struct KmerArgs
{
    KmerView view;
    uint8_t* bases;
    const int32_t* read_off;
    const int32_t* read_site;
    int n_reads;
    const int32_t* todo_in;
    const int32_t* n_todo_in;
    int32_t* todo_out;
    int32_t* n_todo_out;
    uint8_t* prerev;
    Record* records;
    uint32_t* arena;
    unsigned long long* cursor;
    unsigned long long arena_cap;
    unsigned long long* counters; // attempted, mapped (KmerAligner::attempted / mapped)
    int no_gssw;
    int max_len, ops_cap, smem_bytes_per_warp;
};
constexpr int KMER_WARPS = 4;
__global__ void __launch_bounds__(KMER_WARPS * 32) pg_kmer_kernel(const KmerArgs a)
{
    extern __shared__ uint32_t smem[];
    const int wic = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int w = blockIdx.x * KMER_WARPS + wic;
    const int n = a.todo_in ? min(a.n_reads, *a.n_todo_in) : a.n_reads;
    if (w >= n)
        return;
    const int rd = a.todo_in ? a.todo_in[w] : w;
    uint8_t* base = reinterpret_cast<uint8_t*>(smem) + (size_t)wic * a.smem_bytes_per_warp;
    const int ml = (a.max_len + 3) & ~3;
    KmerScratch sc;
    sc.km[0] = reinterpret_cast<uint32_t*>(base);
    sc.km[1] = sc.km[0] + ml;
    sc.bitmap = sc.km[1] + ml;
    sc.ops_best = sc.bitmap + KMER_WIN / 32;
    sc.ops_tmp = sc.ops_best + a.ops_cap;
    sc.heap = reinterpret_cast<KmerCand*>(sc.ops_tmp + a.ops_cap);
    uint8_t* bytes = reinterpret_cast<uint8_t*>(sc.heap + (KMER_MAX_PATHS + 2));
    sc.seq[0] = bytes;
    sc.seq[1] = bytes + ml;
    sc.valid[0] = bytes + 2 * ml;
    sc.valid[1] = bytes + 3 * ml;
    sc.first[0] = bytes + 4 * ml;
    sc.first[1] = bytes + 5 * ml;
    sc.ops_cap = a.ops_cap;
    uint8_t* bases = a.bases + a.read_off[rd];
    const int L = a.read_off[rd + 1] - a.read_off[rd];
    const int site = a.read_site ? a.read_site[rd] : 0;
    if (!a.todo_in && lane == 0)
        a.prerev[rd] = 0;
    const KmerResult r = kmer_align_read(a.view, site, bases, L, lane, 32, sc);
    if (lane == 0)
        atomicAdd(a.counters + 0, 1ull);
    const int flips = a.prerev[rd];
    if (r.status == 1 || (r.status == 2 && a.no_gssw))
    {
        Record rec;
        rec.graph_pos = r.pos;
        rec.score = (int16_t)r.score;
        rec.query_clipped = (uint16_t)r.clipped;
        rec.unique = (uint8_t)(r.status == 1);
        rec.chose_reverse = (uint8_t)r.rev;
        rec.status = 0;
        rec.mapped_by = (uint8_t)(flips ? STAGE_KMER_REV : STAGE_KMER);
        rec.cigar_off = 0;
        rec.cigar_len = (uint32_t)r.n_ops;
        unsigned long long off = 0;
        if (lane == 0)
            off = atomicAdd(a.cursor, (unsigned long long)r.n_ops);
        off = __shfl_sync(FULL, off, 0);
        __syncwarp(); // lane 0 wrote the op words (kmer_build_ops): make them visible to the lanes that copy them out
        if (r.n_ops <= a.ops_cap && off + (unsigned long long)r.n_ops <= a.arena_cap)
        {
            for (int x = lane; x < r.n_ops; x += 32)
                a.arena[off + x] = sc.ops_best[x];
            rec.cigar_off = (uint32_t)off;
        }
        else
        {
            rec.status = 2;
            rec.cigar_len = 0;
        }
        if (lane == 0)
        {
            a.records[rd] = rec;
            if (r.status == 1)
                atomicAdd(a.counters + 1, 1ull);
        }
        return;
    }
    if (a.no_gssw) // no later stage: the read stays UNMAPPED
    {
        if (lane == 0)
        {
            Record rec;
            memset(&rec, 0, sizeof rec);
            rec.status = (uint8_t)ST_UNMAPPED;
            a.records[rd] = rec;
        }
        return;
    }
    if (r.status == 2 && r.rev) // BAD_ALIGN on the reverse strand: the later stages see the reverse complement
    {
        for (int x = lane; x < L; x += 32)
            bases[x] = sc.seq[1][x];
        if (lane == 0)
            a.prerev[rd] = (uint8_t)(flips + 1);
    }
    if (lane == 0)
        a.todo_out[atomicAdd(a.n_todo_out, 1)] = rd;
}

// ---- building the exact-match index on the device --------------------------------------------------------------
// The host build (pg_host.hpp: build_path_index) costs ~25 us per site and thread; for a batch of 1 000 new sites that
// is as long as aligning their reads.  Here: one thread per start position of every site enumerates its k-mer paths
// (depth first over the successors, like extendPathEnd) and
//   pass 1 (pg_path_index_count): claims the table slot of each path's 64-bit hash (atomicCAS on the key) and adds the
//           path to the slot's occurrence count and to the slot's sum of a second, independent 64-bit hash;
//   pass 2 (pg_path_index_fill): a path whose slot counts exactly one occurrence is the unique path of its k-mer and
//           writes the entry (positions, node list); slots with more occurrences stay occupied but unusable.  Should
//           two DIFFERENT k-mers ever share a 64-bit hash, the second-hash sum of their slot differs from
//           count x second hash: the flag is raised and the host builds the index instead (exact in every case).
constexpr int PATH_KMAX_DEV = 64; // deepest stack of the device enumeration; longer k-mers use the host build
constexpr uint64_t PATH_HASH_B2 = 0xD6E8FEB86659FD93ull;

struct PathBuildArgs
{
    const SiteDev* sites;
    const uint8_t* gbytes;
    const int32_t* gints;
    const PathSite* psites;
    const int32_t* psucc;
    const int32_t* col_base; // [n_sites + 1] first start position of each site in the flat thread index
    int n_sites, k;
    PathEntry* table;
    uint32_t* cnt;           // per slot: occurrences
    unsigned long long* sum2; // per slot: sum of the second hash
    int32_t* lists;
    int32_t* list_cursor;    // per site
    const int32_t* list_cap; // per site
    int* flag;               // 1 = hash collision between different k-mers, 2 = a capacity was exceeded
};

template <bool FILL> __global__ void __launch_bounds__(128) pg_path_index_kernel(const PathBuildArgs a)
{
    const int gid = blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= a.col_base[a.n_sites])
        return;
    int lo = 0, hi = a.n_sites - 1; // site of this start position
    while (lo < hi)
    {
        const int mid = (lo + hi + 1) >> 1;
        if (a.col_base[mid] <= gid)
            lo = mid;
        else
            hi = mid - 1;
    }
    const int site = lo;
    const SiteDev sd = a.sites[site];
    const PathSite ps = a.psites[site];
    const int32_t* t = a.gints + sd.tab_off[0];
    const int32_t *node_start = t, *node_len = t + sd.n_nodes;
    const int32_t* succ_ptr = a.psucc + ps.succ_ptr_off;
    const int32_t* succ_idx = succ_ptr + sd.n_nodes + 1;
    const uint8_t* raw = a.gbytes + ps.raw_off;
    PathEntry* tab = a.table + ps.table_off;
    const uint32_t mask = (uint32_t)ps.table_mask;
    const int k = a.k;
    const int col = gid - a.col_base[site];
    int v0 = 0;
    while (v0 + 1 < sd.n_nodes && node_start[v0 + 1] <= col)
        ++v0;
    const int pos = col - node_start[v0];

    int nodes[PATH_KMAX_DEV + 1], endp[PATH_KMAX_DEV + 2], ext[PATH_KMAX_DEV + 2], nxt[PATH_KMAX_DEV + 2];
    int depth = 1;
    nodes[0] = v0;
    endp[1] = pos;
    ext[1] = k - 1;
    nxt[1] = 0;
    while (depth > 0)
    {
        const int last = nodes[depth - 1];
        const int room = node_len[last] - endp[depth] - 1;
        if (ext[depth] > room)
        {
            if (nxt[depth] >= succ_ptr[last + 1] - succ_ptr[last])
            {
                --depth;
                continue;
            }
            const int c = succ_idx[succ_ptr[last] + nxt[depth]++];
            nodes[depth] = c;
            endp[depth + 1] = 0;
            ext[depth + 1] = ext[depth] - room - 1;
            nxt[depth + 1] = 0;
            ++depth;
            continue;
        }
        // a complete k-mer path: nodes[0 .. depth), from pos in the first to end_pos in the last
        const int end_pos = endp[depth] + ext[depth];
        uint64_t h = 0, h2 = 0;
        for (int x = 0; x < depth; ++x)
        {
            const int nd = nodes[x];
            const int p0 = x == 0 ? pos : 0, p1 = x == depth - 1 ? end_pos : node_len[nd] - 1;
            const uint8_t* sq = raw + node_start[nd];
            for (int p = p0; p <= p1; ++p)
            {
                h = path_hash_step(h, sq[p]);
                h2 = h2 * PATH_HASH_B2 + (uint64_t)sq[p] + 1u;
            }
        }
        if (h == 0)
            h = 1;
        uint32_t slot = (uint32_t)(h ^ (h >> 29)) & mask;
        for (uint32_t probes = 0;; slot = (slot + 1) & mask, ++probes)
        {
            unsigned long long* key = reinterpret_cast<unsigned long long*>(&tab[slot]); // key_lo | key_hi << 32
            unsigned long long cur = FILL ? *key : atomicCAS(key, 0ull, (unsigned long long)h);
            if (!FILL && cur == 0ull)
                cur = h;
            if (cur == h)
                break;
            if (probes > mask) // table full: cannot happen with the host's sizing
            {
                atomicExch(a.flag, 2);
                return;
            }
        }
        const size_t gslot = (size_t)ps.table_off + slot;
        if (!FILL)
        {
            atomicAdd(a.cnt + gslot, 1u);
            atomicAdd(a.sum2 + gslot, (unsigned long long)h2);
        }
        else
        {
            const uint32_t n = a.cnt[gslot];
            if (n == 1)
            {
                const int off = atomicAdd(a.list_cursor + site, depth);
                if (off + depth > a.list_cap[site])
                    atomicExch(a.flag, 2);
                else
                {
                    for (int x = 0; x < depth; ++x)
                        a.lists[ps.lists_off + off + x] = nodes[x];
                    tab[slot].start_pos = pos;
                    tab[slot].end_pos = end_pos;
                    tab[slot].nodes_off = off;
                    tab[slot].n_nodes = depth;
                }
            }
            else
            {
                tab[slot].n_nodes = -1; // occupied, not a unique k-mer (several paths write the same value)
                if (a.sum2[gslot] != (unsigned long long)n * h2)
                    atomicExch(a.flag, 1); // different k-mers under one 64-bit hash: let the host do it exactly
            }
        }
        --depth;
    }
}

// ------------------------------------------------------------------------------------------------ host

template <typename T> struct DevBuf
{
    T* p = nullptr;
    size_t cap = 0; // elements
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMalloc(&p, n * sizeof(T));
        if (e == cudaSuccess)
            cap = n;
        return e;
    }
    void release()
    {
        if (p)
            cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    ~DevBuf() { release(); }
};
template <typename T> struct PinBuf
{
    T* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t n)
    {
        if (n <= cap)
            return cudaSuccess;
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
        cudaError_t e = cudaMallocHost(&p, n * sizeof(T));
        if (e == cudaSuccess)
            cap = n;
        return e;
    }
    void release()
    {
        if (p)
            cudaFreeHost(p);
        p = nullptr;
        cap = 0;
    }
};

} // namespace

// ---------------------------------------------------------------------------------------------
// counting stage kernels (logic in pg_count.cuh).  All three are thread-per-item passes over a few dozen bytes per
// read: HBM-latency bound, microseconds per 10k reads next to the milliseconds of the fill.
// ---------------------------------------------------------------------------------------------
struct CountArgs
{
    CountTables t;
    CountParams prm;
    const Record* records;
    const uint32_t* ops;
    const int32_t* read_off;
    const int32_t* read_site; // nullptr = site 0
    const uint8_t* is_rev;    // nullptr = all forward
    int n_reads;
    ReadSupport* sup;
    uint32_t* path;
    const int32_t* next;
    const uint8_t* head;
    Count4 *node_counts, *edge_counts, *fam_counts;
    unsigned long long* fam_keys;
    unsigned long long* cursor; // [0] family words written, [1] site + 1 whose family table overflowed
    uint32_t* fam_out;
    unsigned long long fam_out_cap;
    int n_sites;
};

__global__ void __launch_bounds__(128) pg_support_kernel(CountArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads)
        return;
    const Record rec = a.records[i];
    ReadSupport sup;
    support_read(rec, a.ops, a.read_off[i + 1] - a.read_off[i], a.read_site ? a.read_site[i] : 0,
                 a.is_rev ? a.is_rev[i] != 0 : false, a.t, a.prm, sup, a.path);
    a.sup[i] = sup;
}

__global__ void __launch_bounds__(128) pg_fragment_kernel(CountArgs a)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n_reads || !a.head[i])
        return;
    const int site = a.read_site ? a.read_site[i] : 0;
    if (!count_fragment(i, site, a.next, a.sup, a.path, a.t, a.prm, a.node_counts, a.edge_counts, a.fam_keys,
                        a.fam_counts))
        atomicMax(&a.cursor[1], (unsigned long long)site + 1ull);
}

// one warp per site: its lanes walk the site's slots; occupied slots are appended to the compact family list
__global__ void __launch_bounds__(128) pg_family_compact_kernel(CountArgs a)
{
    const int site = (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (site >= a.n_sites)
        return;
    const CountSite cs = a.t.csite[site];
    const SiteDev& sd = a.t.sites[site];
    const int n = 1 + sd.n_nodes + sd.n_edges;
    for (int slot = lane; slot < cs.slots; slot += 32)
    {
        const unsigned long long key = a.fam_keys[cs.key_base + slot];
        if (!key)
            continue;
        const unsigned long long off = atomicAdd(&a.cursor[0], 4ull + 4ull * (unsigned long long)n);
        if (off + 4ull + 4ull * (unsigned long long)n > a.fam_out_cap)
            continue;
        uint32_t* o = a.fam_out + off;
        o[0] = (uint32_t)site;
        o[1] = (uint32_t)n;
        o[2] = (uint32_t)key;
        o[3] = (uint32_t)(key >> 32);
        const uint32_t* src = reinterpret_cast<const uint32_t*>(a.fam_counts + cs.fam_base + (int64_t)slot * n);
        for (int x = 0; x < 4 * n; ++x)
            o[4 + x] = src[x];
    }
}

struct pg_ctx
{
    int device = 0;
    std::string err;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::vector<cudaEvent_t> evpool; // 4 events per chunk: before / after fill (main stream), before / after trace (aux stream)
    cudaStream_t aux_stream = nullptr; // odd chunks of a split batch run here, next to the even ones on the caller's stream
    int split = 2;                     // chunks a batch is cut into for that overlap (PG_SPLIT; 1 = off; 2 measured best
                                       // on the B200)
    int n_chunks_timed = 0;
    uint64_t launches = 0;
    float fill_ms = 0, trace_ms = 0;
    uint64_t scratch_limit = 64ull << 30; // pg_create lowers it to 40 % of the device's memory if that is less
    bool use_tma = true; // PG_NO_TMA=1 disables the shared-memory staging of column codes (A/B only)
    int geom_w = 32; // lanes per task; PG_GEOM_W=32|16|8 overrides (tuning / A-B measurements only)

    host::GraphStore graphs;
    bool graphs_dirty = true;
    PinBuf<uint8_t> h_graph;        // page-locked staging of the graph store: its upload does not hold up the host
    cudaEvent_t ev_graph = nullptr; // the last upload out of h_graph is through
    bool graph_staging_busy = false;
    DevBuf<SiteDev> d_sites;
    DevBuf<uint8_t> d_gbytes;
    DevBuf<int32_t> d_gints;

    // batch
    int n_reads = 0, max_len = 0;
    bool have_sites = false, uploaded = false, ran = false, staging_busy = false;
    bool imported = false; // the batch came from pg_batch_import: records and op words only, no read bases on the device
    size_t bases_bytes = 0;
    PinBuf<uint8_t> h_bases;
    PinBuf<int32_t> h_off, h_site;
    DevBuf<uint8_t> d_bases;
    DevBuf<int32_t> d_off, d_site;
    DevBuf<uint32_t> d_last, d_ckpt, d_arena, d_tab;
    DevBuf<TaskOut> d_tout;
    DevBuf<Record> d_records;
    DevBuf<unsigned long long> d_cursor;
    PinBuf<Record> h_records;
    PinBuf<uint32_t> h_arena;
    PinBuf<unsigned long long> h_cursor;
    unsigned long long arena_cap = 0;

    // pairing of the reversed-graph fills (rev_plan): on by default for the byte-packed geometries, PG_PAIR_REV=0 = off
    bool pair_rev = true;
    int stagger = 0; // PG_STAGGER: size of a batch's first chunk in per cent of the others (0 = all equal)
    int persist = 0; // PG_PERSIST: 1 = the paired reversed-graph launches, 2 = every fill launch runs as one wave of
                     // persistent CTAs striding over the task list
    int n_sms = 148;
    DevBuf<int32_t> d_rvntop, d_req, d_nreq; // d_nreq: [0] requests, [1] tasks
    DevBuf<uint8_t> d_pending;               // [read] 1 = waits for a second-round reversed-graph fill (pg_plan_kernel)
    cudaStream_t side_stream[2] = { nullptr, nullptr }; // second-round fills run here, next to the traceback of the settled reads
    std::vector<cudaEvent_t> evside;                    // 2 per chunk: round-1 plan done / second-round fill done
    int late_round = 1;                                 // PG_LATE_ROUND=0: second-round fills before the one traceback launch
    DevBuf<int2> d_rtasks;

    // exact-match stage (pg_path.cuh)
    int path_k = 0; // k-mer length; 0 = the stage is off
    bool gssw_on = true; // graphMatching of the cascade
    bool path_second_chance = false;
    bool path_scalar = false;
    DevBuf<uint8_t> d_prerev;
    // second chance: the exact-match kernels leave reverse-complemented bases behind for the DP (as PathAligner does
    // with the read); the uploaded bases are kept aside so that every pg_batch_run starts from what was uploaded
    DevBuf<uint8_t> d_bases_orig;
    bool have_bases_orig = false;
    bool path_dirty = true;
    DevBuf<PathSite> d_psites;
    DevBuf<PathEntry> d_ptable;
    DevBuf<int32_t> d_plists, d_psucc, d_todo, d_ntodo, d_pcolbase, d_plistcap, d_plistcur;
    DevBuf<uint32_t> d_pcnt;
    DevBuf<unsigned long long> d_psum2;
    bool path_host_index = false;      // PG_PATH_HOST_INDEX=1: build the index on the host (A/B, fallback testing)
    bool path_index_on_device = false; // how the current index was built
    DevBuf<unsigned long long> d_pcount;
    unsigned long long path_counters[3] = { 0, 0, 0 };
    unsigned long long path_index_us = 0; // host time of the last index build
    bool path_ran = false;
    float path_ms = 0;
    cudaEvent_t path_ev[2] = { nullptr, nullptr };

    // k-mer stage (pg_kmer.cuh)
    int kmer_k = 0; // 0 = the stage is off
    bool kmer_dirty = true, kmer_ran = false;
    int kmer_max_path_nodes = 0;
    DevBuf<KmerSiteDev> d_ksites;
    DevBuf<KmerPathDev> d_kpaths;
    DevBuf<uint8_t> d_kseqs;
    DevBuf<int32_t> d_knodes, d_todo2, d_ntodo2;
    DevBuf<KmerPos> d_kkmers;
    DevBuf<unsigned long long> d_kcount;
    unsigned long long kmer_counters[2] = { 0, 0 };
    float kmer_ms = 0;
    cudaEvent_t kmer_ev[2] = { nullptr, nullptr };
    // the reads the stages in front left for the DP: the list of the last stage that ran (null = all reads)
    const int32_t* todo_final = nullptr;
    const int32_t* ntodo_final = nullptr;

    // counting stage (pg_count.cuh)
    bool count_dirty = true;
    int count_slots = 0;
    uint64_t count_launches = 0;
    float count_ms = 0;
    DevBuf<CountSite> d_csite;
    DevBuf<int32_t> d_csr_input, d_next;
    DevBuf<uint64_t> d_lab_edge, d_lab_out, d_lab_in;
    DevBuf<uint8_t> d_head, d_isrev;
    DevBuf<ReadSupport> d_support;
    DevBuf<uint32_t> d_path, d_fam_out;
    DevBuf<Count4> d_node_counts, d_edge_counts, d_fam_counts;
    DevBuf<unsigned long long> d_fam_keys, d_count_cursor; // cursor[0] = family words, cursor[1] = overflow site + 1
    int64_t fam_rows = 0, fam_nkeys = 0;
    cudaEvent_t count_ev[2] = { nullptr, nullptr };
};

namespace
{

int fail(pg_ctx* c, int code, const std::string& msg)
{
    if (c)
        c->err = msg;
    return code;
}
#define PG_CUDA(c, call)                                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        cudaError_t e_ = (call);                                                                                       \
        if (e_ != cudaSuccess)                                                                                         \
            return fail(c, PG_E_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));                             \
    } while (0)

// true when p is page-locked host memory known to CUDA (cudaMallocHost / pg_host_alloc / cudaHostRegister)
bool is_pinned(const void* p)
{
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess)
    {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost;
}

// PG_DEBUG_HOST=1: where the host time of the entry points goes (microseconds between laps, on stderr)
struct HostLaps
{
    const bool on;
    const char* const who;
    std::chrono::steady_clock::time_point t;
    explicit HostLaps(const char* w) : on(getenv("PG_DEBUG_HOST") != nullptr), who(w)
    {
        if (on)
            t = std::chrono::steady_clock::now();
    }
    void lap(const char* what)
    {
        if (!on)
            return;
        const auto n = std::chrono::steady_clock::now();
        fprintf(stderr, "[pg host] %s: %s %.1f us\n", who, what, std::chrono::duration<double, std::micro>(n - t).count());
        t = n;
    }
};

int upload_graphs(pg_ctx* c)
{
    if (!c->graphs_dirty)
        return PG_OK;
    if (c->graphs.sites.empty())
        return fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    PG_CUDA(c, c->d_sites.reserve(c->graphs.sites.size()));
    PG_CUDA(c, c->d_gbytes.reserve(c->graphs.bytes.size()));
    PG_CUDA(c, c->d_gints.reserve(c->graphs.ints.size()));
    // through page-locked staging: the copies are asynchronous, the host goes on to launch the kernels behind them
    const size_t b_sites = c->graphs.sites.size() * sizeof(SiteDev), b_bytes = c->graphs.bytes.size(),
                 b_ints = c->graphs.ints.size() * sizeof(int32_t);
    const size_t o_bytes = (b_sites + 15) & ~(size_t)15, o_ints = (o_bytes + b_bytes + 15) & ~(size_t)15;
    if (!c->ev_graph)
        PG_CUDA(c, cudaEventCreateWithFlags(&c->ev_graph, cudaEventDisableTiming));
    if (c->graph_staging_busy) // (a graph registered while the previous one is still on its way: rare)
        PG_CUDA(c, cudaEventSynchronize(c->ev_graph));
    c->graph_staging_busy = false;
    PG_CUDA(c, c->h_graph.reserve(o_ints + b_ints + 16));
    memcpy(c->h_graph.p, c->graphs.sites.data(), b_sites);
    memcpy(c->h_graph.p + o_bytes, c->graphs.bytes.data(), b_bytes);
    memcpy(c->h_graph.p + o_ints, c->graphs.ints.data(), b_ints);
    PG_CUDA(c, cudaMemcpyAsync(c->d_sites.p, c->h_graph.p, b_sites, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(c->d_gbytes.p, c->h_graph.p + o_bytes, b_bytes, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaMemcpyAsync(c->d_gints.p, c->h_graph.p + o_ints, b_ints, cudaMemcpyHostToDevice, c->stream));
    PG_CUDA(c, cudaEventRecord(c->ev_graph, c->stream));
    c->graph_staging_busy = true;
    c->graphs_dirty = false;
    return PG_OK;
}

template <typename T> cudaError_t put(pg_ctx* c, DevBuf<T>& d, const std::vector<T>& h);

// index of the exact-match stage on the host (pg_host.hpp), uploaded
int upload_path_index_host(pg_ctx* c)
{
    host::PathIndexHost ix;
    host::build_path_index(c->graphs, c->path_k, ix);
    PG_CUDA(c, put(c, c->d_psites, ix.sites));
    PG_CUDA(c, put(c, c->d_ptable, ix.table));
    PG_CUDA(c, put(c, c->d_plists, ix.lists));
    PG_CUDA(c, put(c, c->d_psucc, ix.succ));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    return PG_OK;
}

// ... and on the device (pg_path_index_kernel): the host only sizes the tables.  Returns PG_OK with *done = false when
// the device build has to be redone on the host (hash collision between different k-mers / k too long).
int build_path_index_device(pg_ctx* c, bool* done)
{
    *done = false;
    const host::GraphStore& gs = c->graphs;
    const size_t ns = gs.sites.size();
    const int k = c->path_k;
    if (k > PATH_KMAX_DEV || c->path_host_index)
        return PG_OK;
    std::vector<PathSite> psites(ns);
    std::vector<int32_t> succ, col_base(ns + 1, 0), list_cap(ns, 0);
    size_t n_table = 0, n_lists = 0;
    const auto dbg0 = std::chrono::steady_clock::now();
    auto dbg = [&](const char* what) {
        if (getenv("PG_DEBUG_TIMING"))
            fprintf(stderr, "[pg] index build: %s at %.2f ms\n", what,
                    std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - dbg0).count());
    };
    for (size_t si = 0; si < ns; ++si)
    {
        PathSite& ps = psites[si];
        ps.k = k;
        ps.raw_off = gs.sites[si].raw_off;
        ps.succ_ptr_off = (int32_t)succ.size();
        std::vector<std::vector<int32_t>> sv;
        host::build_successors(gs, si, sv, succ);
        int64_t n_paths = 0, list_ints = 0;
        host::count_kmer_paths(gs, si, k, succ.data() + ps.succ_ptr_off, succ.data() + ps.succ_ptr_off + gs.sites[si].n_nodes + 1,
                               n_paths, list_ints);
        const size_t cap = host::path_table_cap((size_t)std::min<int64_t>(n_paths, (int64_t)1 << 40));
        if (n_table + cap > ((size_t)1 << 28) || n_lists + (size_t)list_ints > ((size_t)1 << 30))
            return fail(c, PG_E_GRAPH, "exact-match stage: the graphs have too many k-mer paths for the index ("
                            + std::to_string(n_paths) + " in site " + std::to_string(si) + ")");
        ps.table_off = (int32_t)n_table;
        ps.table_mask = (int32_t)(cap - 1);
        ps.lists_off = (int32_t)n_lists;
        list_cap[si] = (int32_t)list_ints;
        n_table += cap;
        n_lists += (size_t)list_ints;
        col_base[si + 1] = col_base[si] + gs.sites[si].G;
    }
    dbg("sized");
    PG_CUDA(c, put(c, c->d_psites, psites));
    PG_CUDA(c, put(c, c->d_psucc, succ));
    PG_CUDA(c, put(c, c->d_pcolbase, col_base));
    PG_CUDA(c, put(c, c->d_plistcap, list_cap));
    PG_CUDA(c, c->d_ptable.reserve(n_table));
    PG_CUDA(c, c->d_pcnt.reserve(n_table));
    PG_CUDA(c, c->d_psum2.reserve(n_table));
    PG_CUDA(c, c->d_plists.reserve(n_lists + 1));
    PG_CUDA(c, c->d_plistcur.reserve(ns + 1)); // + the flag
    dbg("allocated");
    PG_CUDA(c, cudaMemsetAsync(c->d_ptable.p, 0, n_table * sizeof(PathEntry), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_pcnt.p, 0, n_table * sizeof(uint32_t), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_psum2.p, 0, n_table * sizeof(unsigned long long), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_plistcur.p, 0, (ns + 1) * sizeof(int32_t), c->stream));
    PathBuildArgs ba;
    ba.sites = c->d_sites.p;
    ba.gbytes = c->d_gbytes.p;
    ba.gints = c->d_gints.p;
    ba.psites = c->d_psites.p;
    ba.psucc = c->d_psucc.p;
    ba.col_base = c->d_pcolbase.p;
    ba.n_sites = (int)ns;
    ba.k = k;
    ba.table = c->d_ptable.p;
    ba.cnt = c->d_pcnt.p;
    ba.sum2 = c->d_psum2.p;
    ba.lists = c->d_plists.p;
    ba.list_cursor = c->d_plistcur.p;
    ba.list_cap = c->d_plistcap.p;
    ba.flag = c->d_plistcur.p + ns;
    const int grid = (col_base[ns] + 127) / 128;
    pg_path_index_kernel<false><<<grid, 128, 0, c->stream>>>(ba);
    pg_path_index_kernel<true><<<grid, 128, 0, c->stream>>>(ba);
    PG_CUDA(c, cudaGetLastError());
    c->launches += 2;
    int flag = 0;
    PG_CUDA(c, cudaMemcpyAsync(&flag, ba.flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    dbg(flag ? "kernels done, FLAG raised" : "kernels done");
    *done = flag == 0;
    return PG_OK;
}

int upload_path_index(pg_ctx* c)
{
    if (!c->path_dirty)
        return PG_OK;
    const auto t0 = std::chrono::steady_clock::now();
    bool done = false;
    int rc = build_path_index_device(c, &done);
    if (rc != PG_OK)
        return rc;
    c->path_index_on_device = done;
    if (!done)
    {
        rc = upload_path_index_host(c);
        if (rc != PG_OK)
            return rc;
    }
    c->path_index_us = (unsigned long long)std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count();
    c->path_dirty = false;
    return PG_OK;
}

// CIGAR arena of a batch, in op words: a read emits at most one op per base plus two per node of its path (exact-match
// stage, DP), and never more than the traceback's op log holds (2 L + 16) -- the second bound keeps one many-node
// site of a multi-site batch from multiplying the arena of every read.
unsigned long long arena_words(const pg_ctx* c, int max_nodes)
{
    const unsigned long long per_read = (unsigned long long)std::min(c->max_len + 2 * max_nodes + 8, 2 * c->max_len + 16);
    return (unsigned long long)c->n_reads * per_read;
}

// The stage itself: one launch over all reads of the batch.  Mapped reads get their record and op words here; the
// others are listed in d_todo for the DP kernels.
int run_path_stage(pg_ctx* c)
{
    int rc = upload_path_index(c);
    if (rc != PG_OK)
        return rc;
    PG_CUDA(c, c->d_todo.reserve((size_t)c->n_reads));
    PG_CUDA(c, c->d_ntodo.reserve(1));
    PG_CUDA(c, c->d_pcount.reserve(3));
    PG_CUDA(c, cudaMemsetAsync(c->d_ntodo.p, 0, sizeof(int32_t), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_pcount.p, 0, 3 * sizeof(unsigned long long), c->stream));
    for (auto& ev : c->path_ev)
        if (!ev)
            PG_CUDA(c, cudaEventCreate(&ev));
    PathArgs pa;
    pa.sites = c->d_sites.p;
    pa.gbytes = c->d_gbytes.p;
    pa.gints = c->d_gints.p;
    pa.psites = c->d_psites.p;
    pa.ptable = c->d_ptable.p;
    pa.plists = c->d_plists.p;
    pa.psucc = c->d_psucc.p;
    pa.bases = c->d_bases.p;
    pa.read_off = c->d_off.p;
    pa.read_site = c->have_sites ? c->d_site.p : nullptr;
    pa.n_reads = c->n_reads;
    pa.no_gssw = (c->gssw_on || c->kmer_k > 0) ? 0 : 1; // a later stage takes the reads this one leaves
    pa.second_chance = c->path_second_chance ? 1 : 0;
    pa.prerev = c->d_prerev.p;
    pa.records = c->d_records.p;
    pa.arena = c->d_arena.p;
    pa.cursor = c->d_cursor.p;
    pa.arena_cap = c->arena_cap;
    pa.todo = c->d_todo.p;
    pa.n_todo = c->d_ntodo.p;
    pa.counters = c->d_pcount.p;
    int sw = (c->max_len + 3) / 4; // row stride in words, odd
    sw |= 1;
    pa.row_stride = 4 * sw;
    const size_t path_smem = (size_t)PATH_THREADS * (sizeof(PathResult) + (size_t)pa.row_stride);
    PG_CUDA(c, cudaFuncSetAttribute(pg_path_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)path_smem));
    const size_t pathw_smem = (size_t)PATHW_WARPS * 6 * (size_t)pa.row_stride;
    PG_CUDA(c, cudaFuncSetAttribute(pg_path_warp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pathw_smem));
    PG_CUDA(c, cudaEventRecord(c->path_ev[0], c->stream));
    if (c->path_scalar) // PG_PATH_SCALAR=1: the one-thread-per-strand kernel (A/B only)
        pg_path_kernel<<<(c->n_reads + PATH_THREADS / 2 - 1) / (PATH_THREADS / 2), PATH_THREADS, path_smem, c->stream>>>(pa);
    else
        pg_path_warp_kernel<<<(c->n_reads + PATHW_WARPS - 1) / PATHW_WARPS, PATHW_WARPS * 32, pathw_smem, c->stream>>>(pa);
    PG_CUDA(c, cudaGetLastError());
    ++c->launches;
    PG_CUDA(c, cudaEventRecord(c->path_ev[1], c->stream));
    c->path_ran = true;
    c->todo_final = c->d_todo.p;
    c->ntodo_final = c->d_ntodo.p;
    return PG_OK;
}

// what the stages in front of the DP share: records, the CIGAR arena with its cursor, the flip counts, and the
// uploaded bases kept aside (the stages hand reverse-complemented bases to the later ones; every pg_batch_run starts
// from what was uploaded)
int prepare_front_stages(pg_ctx* c)
{
    PG_CUDA(c, c->d_records.reserve((size_t)c->n_reads));
    PG_CUDA(c, c->d_cursor.reserve(1));
    PG_CUDA(c, c->d_prerev.reserve((size_t)c->n_reads));
    if (c->have_bases_orig) // an earlier run of this upload may have flipped reads
        PG_CUDA(c, cudaMemcpyAsync(c->d_bases.p, c->d_bases_orig.p, c->bases_bytes, cudaMemcpyDeviceToDevice, c->stream));
    else if (c->gssw_on && (c->kmer_k > 0 || c->path_second_chance))
    {
        PG_CUDA(c, c->d_bases_orig.reserve(c->bases_bytes + 16));
        PG_CUDA(c, cudaMemcpyAsync(c->d_bases_orig.p, c->d_bases.p, c->bases_bytes, cudaMemcpyDeviceToDevice, c->stream));
        c->have_bases_orig = true;
    }
    int max_path_nodes = c->graphs.max_nodes;
    for (auto const& sp : c->graphs.paths)
        for (auto const& pth : sp)
            max_path_nodes = std::max(max_path_nodes, (int)pth.size());
    c->arena_cap = arena_words(c, max_path_nodes);
    PG_CUDA(c, c->d_arena.reserve((size_t)c->arena_cap));
    PG_CUDA(c, cudaMemsetAsync(c->d_cursor.p, 0, sizeof(unsigned long long), c->stream));
    c->todo_final = nullptr;
    c->ntodo_final = nullptr;
    return PG_OK;
}

int upload_kmer_index(pg_ctx* c)
{
    if (!c->kmer_dirty)
        return PG_OK;
    host::KmerIndexHost ix;
    host::build_kmer_index(c->graphs, c->kmer_k, ix);
    if (ix.max_paths > KMER_MAX_PATHS)
        return fail(c, PG_E_GRAPH, "k-mer stage: a site has " + std::to_string(ix.max_paths) + " paths (at most "
                        + std::to_string(KMER_MAX_PATHS) + ")");
    c->kmer_max_path_nodes = ix.max_path_nodes;
    PG_CUDA(c, put(c, c->d_ksites, ix.sites));
    PG_CUDA(c, put(c, c->d_kpaths, ix.paths));
    PG_CUDA(c, put(c, c->d_kseqs, ix.seqs));
    PG_CUDA(c, put(c, c->d_knodes, ix.nodes));
    PG_CUDA(c, put(c, c->d_kkmers, ix.kmers));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->kmer_dirty = false;
    return PG_OK;
}

// grm::KmerAligner over the reads the exact-match stage left (or all reads): one launch
int run_kmer_stage(pg_ctx* c)
{
    int rc = upload_kmer_index(c);
    if (rc != PG_OK)
        return rc;
    PG_CUDA(c, c->d_todo2.reserve((size_t)c->n_reads));
    PG_CUDA(c, c->d_ntodo2.reserve(1));
    PG_CUDA(c, c->d_kcount.reserve(2));
    PG_CUDA(c, cudaMemsetAsync(c->d_ntodo2.p, 0, sizeof(int32_t), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_kcount.p, 0, 2 * sizeof(unsigned long long), c->stream));
    for (auto& ev : c->kmer_ev)
        if (!ev)
            PG_CUDA(c, cudaEventCreate(&ev));
    KmerArgs ka;
    ka.view.sites = c->d_ksites.p;
    ka.view.paths = c->d_kpaths.p;
    ka.view.seqs = c->d_kseqs.p;
    ka.view.nodes = c->d_knodes.p;
    ka.view.kmers = c->d_kkmers.p;
    ka.view.k = c->kmer_k;
    ka.bases = c->d_bases.p;
    ka.read_off = c->d_off.p;
    ka.read_site = c->have_sites ? c->d_site.p : nullptr;
    ka.n_reads = c->n_reads;
    ka.todo_in = c->todo_final;
    ka.n_todo_in = c->ntodo_final;
    ka.todo_out = c->d_todo2.p;
    ka.n_todo_out = c->d_ntodo2.p;
    ka.prerev = c->d_prerev.p;
    ka.records = c->d_records.p;
    ka.arena = c->d_arena.p;
    ka.cursor = c->d_cursor.p;
    ka.arena_cap = c->arena_cap;
    ka.counters = c->d_kcount.p;
    ka.no_gssw = c->gssw_on ? 0 : 1;
    ka.max_len = c->max_len;
    ka.ops_cap = c->max_len + 2 * c->kmer_max_path_nodes + 8;
    const int ml = (c->max_len + 3) & ~3;
    ka.smem_bytes_per_warp = (2 * ml + KMER_WIN / 32 + 2 * ka.ops_cap) * 4 + (KMER_MAX_PATHS + 2) * (int)sizeof(KmerCand) + 6 * ml;
    ka.smem_bytes_per_warp = (ka.smem_bytes_per_warp + 15) & ~15;
    const size_t smem = (size_t)KMER_WARPS * ka.smem_bytes_per_warp;
    PG_CUDA(c, cudaFuncSetAttribute(pg_kmer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    PG_CUDA(c, cudaEventRecord(c->kmer_ev[0], c->stream));
    pg_kmer_kernel<<<(c->n_reads + KMER_WARPS - 1) / KMER_WARPS, KMER_WARPS * 32, smem, c->stream>>>(ka);
    PG_CUDA(c, cudaGetLastError());
    ++c->launches;
    PG_CUDA(c, cudaEventRecord(c->kmer_ev[1], c->stream));
    c->kmer_ran = true;
    c->todo_final = c->d_todo2.p;
    c->ntodo_final = c->d_ntodo2.p;
    return PG_OK;
}

// host -> device copy of a small pageable vector (synchronous with respect to the host buffer after the sync below)
template <typename T> cudaError_t put(pg_ctx* c, DevBuf<T>& d, const std::vector<T>& h)
{
    cudaError_t e = d.reserve(h.size() ? h.size() : 1);
    if (e != cudaSuccess || h.empty())
        return e;
    return cudaMemcpyAsync(d.p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, c->stream);
}

// per-site tables of the counting stage (built by pg_host.hpp)
int upload_count_tables(pg_ctx* c, int slots)
{
    if (!c->count_dirty && c->count_slots == slots)
        return PG_OK;
    const host::GraphStore& gs = c->graphs;
    const size_t ns = gs.sites.size();
    if (gs.edge_base[ns] > 0x7FFFFFFF || gs.node_base[ns] > 0x7FFFFFFF)
        return fail(c, PG_E_GRAPH, "counting stage: more than 2^31 nodes or edges registered");
    host::CountHostTables t;
    host::build_count_tables(gs, slots, t);
    c->fam_rows = t.fam_rows;
    c->fam_nkeys = t.fam_keys;
    PG_CUDA(c, put(c, c->d_csite, t.csite));
    PG_CUDA(c, put(c, c->d_csr_input, t.csr_input));
    PG_CUDA(c, put(c, c->d_lab_edge, t.lab_edge));
    PG_CUDA(c, put(c, c->d_lab_out, t.lab_out));
    PG_CUDA(c, put(c, c->d_lab_in, t.lab_in));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->count_dirty = false;
    c->count_slots = slots;
    return PG_OK;
}

// device -> caller buffer (pinned: direct; pageable: synchronous copy)
cudaError_t get(pg_ctx* c, void* dst, const void* src, size_t bytes)
{
    if (!bytes)
        return cudaSuccess;
    return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, c->stream);
}

template <int R, int W> int run_chunks(pg_ctx* c, unsigned flags)
{
    constexpr int NT = 32 / W;
    HostLaps laps("run_chunks");
    const int max_nodes = c->graphs.max_nodes, max_G = c->graphs.max_G;
    const size_t s_last = host::last_words(max_nodes, R, W), s_ckpt = host::ckpt_words(max_G, R, W);
    // fill kernel shared memory per task: profile + seed/info tables; fewer warps per CTA for many-node graphs, and
    // beyond that the tables move to HBM
    const int tab_words = max_nodes * Sizes<R, W>::ROWW * W; // node table (pg_core.cuh: Sizes)
    // staged column codes (+ 8 bytes for the mbarrier in front, kept 16-byte aligned); graphs over 16 KB are read from L1/L2
    int code_bytes = (int)code_span_bytes(max_G) + 16;
    if (code_bytes > 16 * 1024 + 16 || !c->use_tma)
        code_bytes = 0;
    // staged with the codes (keeps 16-byte alignment), plus one entry word per node (pg_core.cuh: entry_word)
    const int tab_ints_cap = code_bytes ? ((c->graphs.max_tab_ints + max_nodes + 3) & ~3) : 0;
    int fill_words = NCODE * R * W + tab_words + tab_ints_cap + code_bytes / 4;
    // warps per CTA: the count that keeps most warps resident on an SM (shared memory is what limits residency here:
    // per task the profile + the seed / node-maximum tables of max_nodes nodes + the staged codes).  4 warps per CTA
    // for the usual 3-node graphs (41 KB per CTA, 5 CTAs); a batch with 9-node graphs (vcf2paragraph-shaped sites) needs
    // 20 KB per warp -- 4-warp CTAs would leave 2 CTAs = 8 warps on an SM, 1-warp CTAs fit 11.
    int fill_warps = FILL_WARPS;
    {
        const size_t sm_bytes = 227 * 1024, per_cta_extra = 1024; // usable shared memory per SM, per-CTA reservation
        size_t best = 0;
        for (int w = FILL_WARPS; w >= 1; w >>= 1)
        {
            const size_t cta = (size_t)w * NT * fill_words * 4;
            if (cta > 200 * 1024)
                continue;
            const size_t resident = std::min<size_t>(32, sm_bytes / (cta + per_cta_extra)) * (size_t)w;
            if (resident > best)
            {
                best = resident;
                fill_warps = w;
            }
        }
        if (const char* e = getenv("PG_FILL_WARPS_RT")) // A/B
            fill_warps = std::max(1, std::min(FILL_WARPS, atoi(e)));
        while (fill_warps > 1 && (size_t)fill_warps * NT * fill_words * 4 > 200 * 1024)
            fill_warps >>= 1;
    }
    bool tab_global = false;
    if ((size_t)fill_warps * NT * fill_words * 4 > 200 * 1024 || getenv("PG_FORCE_TABG"))
    {
        tab_global = true;
        code_bytes = 0; // the HBM-table variant reads the codes through L1
        fill_words = NCODE * R * W;
        fill_warps = FILL_WARPS;
    }
    const size_t per_read_bytes = (s_last + s_ckpt + (tab_global ? 2 * (size_t)tab_words : 0)) * sizeof(uint32_t);
    size_t chunk = (size_t)(c->scratch_limit / (per_read_bytes ? per_read_bytes : 1));
    if (chunk < 1)
        chunk = 1;
    if (chunk > (size_t)c->n_reads)
        chunk = (size_t)c->n_reads;
    // Overlap: the traceback kernel is latency-bound, the fill kernel ALU-bound -- cut the batch into a few chunks and
    // run the traceback of chunk i on a second stream while chunk i+1 is being filled (scratch double-buffered).
    const bool overlap = c->split > 1 && c->aux_stream && c->n_reads >= 1024;
    if (overlap)
    {
        chunk = std::min(chunk / 2 > 0 ? chunk / 2 : 1, ((size_t)c->n_reads + c->split - 1) / c->split);
        chunk = (chunk + 63) & ~(size_t)63; // whole CTAs
    }
    const size_t slots = overlap ? 2 : 1;
    PG_CUDA(c, c->d_last.reserve(slots * chunk * s_last));
    if (tab_global)
        PG_CUDA(c, c->d_tab.reserve(slots * chunk * 2 * (size_t)tab_words));
    PG_CUDA(c, c->d_ckpt.reserve(slots * chunk * s_ckpt));
    PG_CUDA(c, c->d_tout.reserve((size_t)c->n_reads * 2));
    PG_CUDA(c, c->d_records.reserve((size_t)c->n_reads));
    PG_CUDA(c, c->d_cursor.reserve(1));
    const int oplog_cap = 2 * c->max_len + 16;
    if (!(c->path_ran || c->kmer_ran)) // (else sized by prepare_front_stages, and partly filled)
    {
        c->arena_cap = arena_words(c, max_nodes);
        PG_CUDA(c, c->d_arena.reserve((size_t)c->arena_cap));
    }
    const bool after_path = c->path_ran || c->kmer_ran; // a stage in front already wrote records / op words
    if (!after_path)
        PG_CUDA(c, cudaMemsetAsync(c->d_cursor.p, 0, sizeof(unsigned long long), c->stream));

    const size_t fill_smem = (size_t)fill_warps * NT * fill_words * sizeof(uint32_t);
    const int trace_bytes = (NCODE * R * W + oplog_cap + 2 * TileGeom<R>::SLOT_WORDS) * 4; // per read
    const int trace_bytes_al = (trace_bytes + 15) & ~15;
    const size_t trace_smem = (size_t)TRACE_WARPS * NT * trace_bytes_al;
    if (fill_smem > 227 * 1024 || trace_smem > 227 * 1024)
        return fail(c, PG_E_GRAPH, "graph has too many nodes for the shared-memory seed table ("
                        + std::to_string(max_nodes) + " nodes)");
    PG_CUDA(c, cudaFuncSetAttribute(pg_fill_kernel<R, W, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
    PG_CUDA(c, cudaFuncSetAttribute(pg_fill_kernel<R, W, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
    PG_CUDA(c, cudaFuncSetAttribute(pg_fill_kernel<R, W, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fill_smem));
    PG_CUDA(c, cudaFuncSetAttribute(pg_trace_kernel<R, W>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trace_smem));

    laps.lap("buffers + function attributes");
    // chunk boundaries.  PG_STAGGER: a first chunk of half the size puts the two streams' sequences out of phase, so that the
    // (latency-bound) traceback of one chunk runs beside the (ALU-bound) fill of the next instead of beside the other traceback
    std::vector<size_t> bounds{ 0 };
    {
        size_t first = chunk;
        if (overlap && c->stagger && chunk >= 256)
            first = (chunk * (size_t)c->stagger / 100 + 63) & ~(size_t)63;
        for (size_t r = std::min(first, (size_t)c->n_reads); ; r = std::min(r + chunk, (size_t)c->n_reads))
        {
            bounds.push_back(r);
            if (r >= (size_t)c->n_reads)
                break;
        }
    }
    const size_t n_chunks = bounds.size() - 1;
    while (c->evpool.size() < 4 * n_chunks)
    {
        cudaEvent_t e;
        PG_CUDA(c, cudaEventCreate(&e));
        c->evpool.push_back(e);
    }
    while (c->evside.size() < 2 * n_chunks)
    {
        cudaEvent_t e;
        PG_CUDA(c, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        c->evside.push_back(e);
    }
    c->n_chunks_timed = (int)n_chunks;
    // Reversed-graph fills: only the halves the strand rule needs (rev_plan), two reads per task.  Not for the WIDE
    // geometries (their region maxima depend on the one read length of a task).
    const bool pairs = c->pair_rev && !Sizes<R, W>::WIDE && (flags & AF_REVERSE_GRAPH);
    if (pairs)
    {
        PG_CUDA(c, c->d_rvntop.reserve((size_t)c->n_reads * 2));
        PG_CUDA(c, c->d_req.reserve(2 * (size_t)c->n_reads));
        PG_CUDA(c, c->d_rtasks.reserve(2 * (size_t)c->n_reads));
        PG_CUDA(c, c->d_nreq.reserve(4));
        PG_CUDA(c, c->d_pending.reserve((size_t)c->n_reads));
        PG_CUDA(c, cudaMemsetAsync(c->d_rvntop.p, 0xFF, (size_t)c->n_reads * 2 * sizeof(int32_t), c->stream)); // -1 = unknown
    }
    // grid of one wave of resident CTAs (persistent launch), 0 = not wanted
    int wave = 0;
    if (c->persist > 0 && !tab_global)
    {
        int occ = 0;
        if (code_bytes)
            PG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pg_fill_kernel<R, W, true, false>, fill_warps * 32, fill_smem));
        else
            PG_CUDA(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, pg_fill_kernel<R, W, false, false>, fill_warps * 32, fill_smem));
        wave = occ * c->n_sms;
    }
    auto launch_fill = [&](FillArgs fa, int n_tasks, cudaStream_t st, bool few = false) {
        int fgrid = (n_tasks + fill_warps * NT - 1) / (fill_warps * NT);
        fa.strided = 0;
        if (few && !tab_global) // next to no tasks expected: a small grid of the non-staged instantiation strides over them
        {
            fa.strided = 1;
            fgrid = std::min(fgrid, 2 * c->n_sms);
            pg_fill_kernel<R, W, false, false><<<fgrid, fill_warps * 32, fill_smem, st>>>(fa);
            ++c->launches;
            return;
        }
        if (wave > 0 && fgrid > wave && (c->persist >= 2 || fa.mode == MODE_PAIRS))
        {
            fa.strided = 1;
            fgrid = wave;
        }
        if (tab_global)
            pg_fill_kernel<R, W, false, true><<<fgrid, fill_warps * 32, fill_smem, st>>>(fa);
        else if (code_bytes)
            pg_fill_kernel<R, W, true, false><<<fgrid, fill_warps * 32, fill_smem, st>>>(fa);
        else
            pg_fill_kernel<R, W, false, false><<<fgrid, fill_warps * 32, fill_smem, st>>>(fa);
        ++c->launches;
    };
    // overlap mode: chunk ci runs its whole sequence (fill, plan / pair / paired fills, traceback) on stream ci & 1, so
    // that the tail of one chunk's launches is filled by the other chunk's kernels; the auxiliary stream starts after
    // everything already queued on the caller's stream
    if (overlap)
    {
        PG_CUDA(c, cudaEventRecord(c->evpool[0], c->stream));
        PG_CUDA(c, cudaStreamWaitEvent(c->aux_stream, c->evpool[0], 0));
    }
    laps.lap("events + occupancy");
    size_t ci = 0;
    for (; ci < n_chunks; ++ci)
    {
        const size_t r0 = bounds[ci];
        const size_t slot = overlap ? (ci & 1) : 0;
        cudaStream_t cs = (overlap && (ci & 1)) ? c->aux_stream : c->stream; // chunk ci-2 used the same slot on the same stream
        nvtxRangePushA(ci & 1 ? "pg chunk (aux stream): fill + plan/pair + traceback" : "pg chunk (main stream): fill + plan/pair + traceback");
        PG_CUDA(c, cudaEventRecord(c->evpool[4 * ci], cs));
        const int nr = (int)(bounds[ci + 1] - r0);
        FillArgs fa;
        fa.sites = c->d_sites.p;
        fa.gbytes = c->d_gbytes.p;
        fa.gints = c->d_gints.p;
        fa.bases = c->d_bases.p;
        fa.read_off = c->d_off.p;
        fa.read_site = c->have_sites ? c->d_site.p : nullptr;
        fa.read0 = (int)r0;
        fa.n_tasks = pairs ? nr : 2 * nr;
        fa.mode = pairs ? MODE_FWD : MODE_BOTH;
        fa.rtasks = nullptr;
        fa.n_rtasks = nullptr;
        fa.rv_ntop = nullptr;
        fa.flags = flags;
        fa.last = c->d_last.p + slot * chunk * s_last;
        fa.ckpt = c->d_ckpt.p + slot * chunk * s_ckpt;
        fa.n_nodes_cap = max_nodes;
        fa.code_smem_bytes = code_bytes ? code_bytes - 16 : 0;
        fa.tab_ints_cap = code_bytes ? tab_ints_cap : 0;
        fa.tabG = tab_global ? c->d_tab.p + slot * chunk * 2 * (size_t)tab_words : nullptr;
        fa.stride_tab = (size_t)tab_words;
        fa.stride_last = s_last;
        fa.stride_ckpt = s_ckpt;
        fa.tout = c->d_tout.p;
        fa.smem_words_per_task = fill_words;
        fa.todo = after_path ? c->todo_final : nullptr;
        fa.n_todo = after_path ? c->ntodo_final : nullptr;
        launch_fill(fa, fa.n_tasks, cs);
        PG_CUDA(c, cudaGetLastError());
        bool late = false;
        if (pairs)
        {
            // two rounds: the half rev_plan asks for first, then -- for the few reads it did not settle -- the other one
            PlanArgs pa;
            pa.tout = c->d_tout.p;
            pa.rv_ntop = c->d_rvntop.p;
            pa.read_site = fa.read_site;
            pa.todo = fa.todo;
            pa.n_todo = fa.n_todo;
            pa.read0 = (int)r0;
            pa.n_reads = nr;
            pa.flags = flags;
            pa.req = c->d_req.p + slot * (size_t)c->n_reads;
            pa.n_req = c->d_nreq.p + 2 * slot;
            pa.rtasks = c->d_rtasks.p + slot * (size_t)c->n_reads;
            pa.n_rtasks = pa.n_req + 1;
            FillArgs fr = fa;
            fr.mode = MODE_PAIRS;
            fr.rtasks = pa.rtasks;
            fr.n_rtasks = pa.n_rtasks;
            fr.rv_ntop = c->d_rvntop.p;
            fr.todo = nullptr;
            fr.n_todo = nullptr;
            // The second round is next to empty (a handful of reads in 10 000) but costs a whole fill latency; with
            // `late` it runs on a side stream while the traceback of the settled reads is under way, and a second
            // traceback launch picks up the reads that waited for it (TraceArgs::pending_mode).
            // (worth it while that latency is a visible share of the chunk: measured -4 % on 2 x 5 000 reads, nothing on
            // 2 x 12 000, +1 % on 2 x 48 000 where the second traceback launch costs more than the wait it hides)
            late = c->late_round && c->side_stream[slot] != nullptr && (c->late_round > 1 || nr <= 32768);
            for (int round = 0; round < 2; ++round)
            {
                PG_CUDA(c, cudaMemsetAsync(pa.n_req, 0, 2 * sizeof(int32_t), cs));
                pa.pending = (late && round == 1) ? c->d_pending.p : nullptr;
                pg_plan_kernel<<<(nr + 127) / 128, 128, 0, cs>>>(pa);
                pg_pair_kernel<<<((nr + 1) / 2 + 127) / 128, 128, 0, cs>>>(pa);
                c->launches += 2;
                // at most one task per request: the grid covers the worst case (round 0 needs about nr / 2 tasks, round 1
                // next to none) and the CTAs beyond *n_rtasks leave at once
                fr.n_tasks = nr;
                cudaStream_t fs = cs;
                if (late && round == 1)
                {
                    fs = c->side_stream[slot];
                    PG_CUDA(c, cudaEventRecord(c->evside[2 * ci], cs));
                    PG_CUDA(c, cudaStreamWaitEvent(fs, c->evside[2 * ci], 0));
                }
                launch_fill(fr, nr, fs, round == 1);
                PG_CUDA(c, cudaGetLastError());
                if (fs != cs)
                    PG_CUDA(c, cudaEventRecord(c->evside[2 * ci + 1], fs));
            }
        }
        PG_CUDA(c, cudaEventRecord(c->evpool[4 * ci + 1], cs));
        cudaStream_t ts = cs;
        PG_CUDA(c, cudaEventRecord(c->evpool[4 * ci + 2], ts));

        TraceArgs ta;
        ta.sites = c->d_sites.p;
        ta.gbytes = c->d_gbytes.p;
        ta.gints = c->d_gints.p;
        ta.bases = c->d_bases.p;
        ta.read_off = c->d_off.p;
        ta.read_site = fa.read_site;
        ta.read0 = (int)r0;
        ta.n_reads = nr;
        ta.flags = flags;
        ta.last = fa.last;
        ta.ckpt = fa.ckpt;
        ta.stride_last = s_last;
        ta.stride_ckpt = s_ckpt;
        ta.tout = c->d_tout.p;
        ta.records = c->d_records.p;
        ta.arena = c->d_arena.p;
        ta.cursor = c->d_cursor.p;
        ta.arena_cap = c->arena_cap;
        ta.smem_bytes_per_task = trace_bytes_al;
        ta.oplog_cap = oplog_cap;
        ta.todo = fa.todo;
        ta.n_todo = fa.n_todo;
        ta.prerev = after_path ? c->d_prerev.p : nullptr;
        ta.rv_ntop = pairs ? c->d_rvntop.p : nullptr;
        const int tgrid = (nr + TRACE_WARPS * NT - 1) / (TRACE_WARPS * NT);
        ta.pending = late ? c->d_pending.p : nullptr;
        ta.pending_mode = late ? 1 : 0;
        pg_trace_kernel<R, W><<<tgrid, TRACE_WARPS * 32, trace_smem, ts>>>(ta);
        PG_CUDA(c, cudaGetLastError());
        ++c->launches;
        if (late)
        {
            PG_CUDA(c, cudaStreamWaitEvent(ts, c->evside[2 * ci + 1], 0));
            ta.pending_mode = 2;
            pg_trace_kernel<R, W><<<tgrid, TRACE_WARPS * 32, trace_smem, ts>>>(ta);
            PG_CUDA(c, cudaGetLastError());
            ++c->launches;
        }
        PG_CUDA(c, cudaEventRecord(c->evpool[4 * ci + 3], ts));
        nvtxRangePop();
    }
    if (overlap) // everything later on the caller's stream (download, counting stage) sees both streams finished
        for (size_t k = n_chunks >= 2 ? n_chunks - 2 : 0; k < n_chunks; ++k)
            PG_CUDA(c, cudaStreamWaitEvent(c->stream, c->evpool[4 * k + 3], 0));
    laps.lap("launches");
    return PG_OK;
}

} // namespace

extern "C" {

#define PG_STR_(x) #x
#define PG_STR(x) PG_STR_(x)
const char* pg_version(void)
{
#if PG_SPEC_DEAD
    return "paragraph_b200 0.4 sm_100a CK=" PG_STR(PG_CK) " int16x2-wavefront W=16/32/8 reads<=1024 spec-dead-blocks lean-node-events";
#else
    return "paragraph_b200 0.4 sm_100a CK=" PG_STR(PG_CK) " int16x2-wavefront W=16/32/8 reads<=1024";
#endif
}

int pg_create(int device, pg_ctx** out)
{
    if (!out)
        return PG_E_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n)
    {
        fprintf(stderr, "paragraph_b200: no usable CUDA device %d (%s) -- there is no CPU fallback\n", device,
                e != cudaSuccess ? cudaGetErrorString(e) : "bad ordinal");
        return PG_E_CUDA;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    {
        fprintf(stderr, "paragraph_b200: device %d is not sm_100+ -- kernels are built for sm_100a only\n", device);
        return PG_E_CUDA;
    }
    pg_ctx* c = new pg_ctx;
    c->device = device;
    if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess)
    {
        delete c;
        return PG_E_CUDA;
    }
    c->stream = c->own_stream;
    {
        // PG_PRIO=1: the auxiliary stream (odd chunks) at the lowest priority, 2: at the highest (A/B)
        int least = 0, greatest = 0;
        const char* e = getenv("PG_PRIO");
        const int prio = e ? atoi(e) : 0;
        cudaError_t rc;
        if (prio && cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess)
            rc = cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio == 1 ? least : greatest);
        else
            rc = cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking);
        if (rc != cudaSuccess)
            c->aux_stream = nullptr;
    }
    for (auto& st : c->side_stream)
        if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess)
            st = nullptr;
    if (const char* e = getenv("PG_LATE_ROUND"))
        c->late_round = atoi(e);
    if (const char* e = getenv("PG_SPLIT"))
        c->split = std::max(1, atoi(e));
    if (const char* e = getenv("PG_STAGGER"))
        c->stagger = std::max(0, std::min(100, atoi(e)));
    if (const char* e = getenv("PG_NO_TMA"))
        c->use_tma = atoi(e) == 0;
    if (const char* e = getenv("PG_PAIR_REV"))
        c->pair_rev = atoi(e) != 0;
    if (const char* e = getenv("PG_PERSIST"))
        c->persist = atoi(e);
    cudaDeviceGetAttribute(&c->n_sms, cudaDevAttrMultiProcessorCount, device);
    if (const char* e = getenv("PG_PATH_HOST_INDEX"))
        c->path_host_index = atoi(e) != 0;
    if (const char* e = getenv("PG_PATH_SCALAR"))
        c->path_scalar = atoi(e) != 0;
    {
        size_t free_b = 0, total_b = 0;
        if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess && total_b > 0)
            c->scratch_limit = std::min<uint64_t>(c->scratch_limit, (uint64_t)(total_b / 10 * 4));
    }
    if (const char* e = getenv("PG_SCRATCH_GB")) // A/B: the default of pg_set_scratch_limit
        c->scratch_limit = (uint64_t)std::max(1, atoi(e)) << 30;
    if (const char* e = getenv("PG_GEOM_W"))
    {
        const int w = atoi(e);
        if (w == 32 || w == 16 || w == 8)
            c->geom_w = w;
    }
    *out = c;
    return PG_OK;
}

void pg_destroy(pg_ctx* c)
{
    if (!c)
        return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->d_sites.release();
    c->d_gbytes.release();
    c->d_gints.release();
    c->d_bases.release();
    c->d_off.release();
    c->d_site.release();
    c->d_last.release();
    c->d_tab.release();
    c->d_ckpt.release();
    c->d_arena.release();
    c->d_tout.release();
    c->d_records.release();
    c->d_cursor.release();
    c->d_psites.release();
    c->d_ptable.release();
    c->d_plists.release();
    c->d_psucc.release();
    c->d_todo.release();
    c->d_ntodo.release();
    c->d_pcount.release();
    c->d_prerev.release();
    c->d_bases_orig.release();
    c->d_ksites.release();
    c->d_kpaths.release();
    c->d_kseqs.release();
    c->d_knodes.release();
    c->d_kkmers.release();
    c->d_todo2.release();
    c->d_ntodo2.release();
    c->d_kcount.release();
    for (auto& ev : c->kmer_ev)
        if (ev)
            cudaEventDestroy(ev);
    c->d_rvntop.release();
    c->d_req.release();
    c->d_nreq.release();
    c->d_pending.release();
    c->d_rtasks.release();
    c->d_pcolbase.release();
    c->d_plistcap.release();
    c->d_plistcur.release();
    c->d_pcnt.release();
    c->d_psum2.release();
    for (auto& ev : c->path_ev)
        if (ev)
            cudaEventDestroy(ev);
    c->h_bases.release();
    c->h_graph.release();
    if (c->ev_graph)
        cudaEventDestroy(c->ev_graph);
    c->h_off.release();
    c->h_site.release();
    c->h_records.release();
    c->h_arena.release();
    c->h_cursor.release();
    for (auto& ev : c->evpool)
        cudaEventDestroy(ev);
    for (auto& ev : c->count_ev)
        if (ev)
            cudaEventDestroy(ev);
    if (c->aux_stream)
        cudaStreamDestroy(c->aux_stream);
    for (auto st : c->side_stream)
        if (st)
            cudaStreamDestroy(st);
    for (auto& ev : c->evside)
        cudaEventDestroy(ev);
    if (c->own_stream)
        cudaStreamDestroy(c->own_stream);
    delete c;
}

const char* pg_last_error(const pg_ctx* c) { return c ? c->err.c_str() : "null context"; }

int pg_set_stream(pg_ctx* c, void* s)
{
    if (!c)
        return PG_E_ARG;
    cudaStream_t next = s ? (cudaStream_t)s : c->own_stream;
    if (next != c->stream && c->graph_staging_busy) // work on the new stream must see the graphs copied on the old one
    {
        PG_CUDA(c, cudaSetDevice(c->device));
        PG_CUDA(c, cudaStreamWaitEvent(next, c->ev_graph, 0));
    }
    c->stream = next;
    return PG_OK;
}

int pg_set_scratch_limit(pg_ctx* c, uint64_t bytes)
{
    if (!c || bytes < (1u << 20))
        return PG_E_ARG;
    c->scratch_limit = bytes;
    return PG_OK;
}

int pg_add_graph(pg_ctx* c, int32_t n_nodes, const char* blob, const int32_t* off, int32_t n_edges, const int32_t* ef,
                 const int32_t* et, int32_t* site_id)
{
    if (!c)
        return PG_E_ARG;
    std::string err;
    int id = c->graphs.add(n_nodes, blob, off, n_edges, ef, et, err);
    if (id < 0)
        return fail(c, PG_E_GRAPH, err);
    c->graphs_dirty = true;
    c->count_dirty = true;
    c->path_dirty = true;
    c->kmer_dirty = true;
    if (site_id)
        *site_id = id;
    return PG_OK;
}

int pg_add_graphs(pg_ctx* c, int32_t n_sites, const int32_t* node_ptr, const char* blob, const int32_t* off,
                  const int32_t* edge_ptr, const int32_t* ef, const int32_t* et, int32_t* first_site_id)
{
    if (!c || n_sites < 0 || (n_sites > 0 && (!node_ptr || !blob || !off || !edge_ptr)))
        return fail(c, PG_E_ARG, "pg_add_graphs: bad arguments");
    const host::GraphStore::Mark mark = c->graphs.mark();
    const int first = (int)c->graphs.sites.size();
    std::string err;
    for (int32_t s = 0; s < n_sites; ++s)
    {
        const int32_t n0 = node_ptr[s], e0 = edge_ptr[s];
        const int id = c->graphs.add(node_ptr[s + 1] - n0, blob, off + n0, edge_ptr[s + 1] - e0, ef ? ef + e0 : nullptr,
                                     et ? et + e0 : nullptr, err);
        if (id < 0)
        {
            c->graphs.rollback(mark);
            return fail(c, PG_E_GRAPH, "site " + std::to_string(s) + " of the batch: " + err);
        }
    }
    if (n_sites > 0)
        c->graphs_dirty = c->count_dirty = c->path_dirty = c->kmer_dirty = true;
    if (first_site_id)
        *first_site_id = first;
    return PG_OK;
}

int pg_clear_graphs(pg_ctx* c)
{
    if (!c)
        return PG_E_ARG;
    c->graphs.clear();
    c->graphs_dirty = true;
    c->count_dirty = true;
    c->path_dirty = true;
    c->kmer_dirty = true;
    // the site ids of an uploaded / imported batch name graphs that are gone now: the batch goes with them
    c->uploaded = c->ran = false;
    return PG_OK;
}

int pg_batch_upload(pg_ctx* c, int32_t n_reads, const char* bases, const int32_t* off, const int32_t* site)
{
    struct Range
    {
        Range() { nvtxRangePushA("pg_batch_upload"); }
        ~Range() { nvtxRangePop(); }
    } range;
    HostLaps laps("pg_batch_upload");
    if (!c || n_reads < 0 || (n_reads > 0 && (!bases || !off)))
        return fail(c, PG_E_ARG, "pg_batch_upload: bad arguments");
    if (n_reads == 0) // an empty batch is legal (grm::alignReads on an empty read vector does nothing)
    {
        c->n_reads = 0;
        c->uploaded = true;
        c->imported = false;
        c->ran = false;
        return PG_OK;
    }
    PG_CUDA(c, cudaSetDevice(c->device));
    const int nsites = (int)c->graphs.sites.size();
    if (nsites == 0)
        return fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    int maxl = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        const int l = off[i + 1] - off[i];
        if (l <= 0 || l > PG_MAX_READ_LEN)
            return fail(c, PG_E_READ_LEN, "read " + std::to_string(i) + " has length " + std::to_string(l)
                            + " (supported: 1.." + std::to_string(PG_MAX_READ_LEN) + ")");
        if (l > maxl)
            maxl = l;
        if (site && (site[i] < 0 || site[i] >= nsites))
            return fail(c, PG_E_ARG, "read " + std::to_string(i) + " names unknown site " + std::to_string(site[i]));
    }
    const size_t nb = (size_t)(off[n_reads] - off[0]);
    PG_CUDA(c, c->d_bases.reserve(nb + 16));
    PG_CUDA(c, c->d_off.reserve((size_t)n_reads + 1));
    if (c->staging_busy) // a previous upload's copies may still read the staging buffers
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->staging_busy = false;
    // Caller buffers that are already page-locked (pg_host_alloc) are copied from directly; pageable ones go
    // through the context's pinned staging buffers.
    const bool direct = off[0] == 0 && is_pinned(bases) && is_pinned(off);
    if (direct)
    {
        PG_CUDA(c, cudaMemcpyAsync(c->d_bases.p, bases, nb, cudaMemcpyHostToDevice, c->stream));
        PG_CUDA(c, cudaMemcpyAsync(c->d_off.p, off, ((size_t)n_reads + 1) * sizeof(int32_t), cudaMemcpyHostToDevice,
                                   c->stream));
    }
    else
    {
        PG_CUDA(c, c->h_bases.reserve(nb + 16));
        PG_CUDA(c, c->h_off.reserve((size_t)n_reads + 1));
        memcpy(c->h_bases.p, bases + off[0], nb);
        for (int i = 0; i <= n_reads; ++i)
            c->h_off.p[i] = off[i] - off[0];
        PG_CUDA(c, cudaMemcpyAsync(c->d_bases.p, c->h_bases.p, nb, cudaMemcpyHostToDevice, c->stream));
        PG_CUDA(c, cudaMemcpyAsync(c->d_off.p, c->h_off.p, ((size_t)n_reads + 1) * sizeof(int32_t),
                                   cudaMemcpyHostToDevice, c->stream));
        c->staging_busy = true;
    }
    c->have_sites = site != nullptr;
    if (site)
    {
        PG_CUDA(c, c->h_site.reserve((size_t)n_reads));
        PG_CUDA(c, c->d_site.reserve((size_t)n_reads));
        memcpy(c->h_site.p, site, (size_t)n_reads * sizeof(int32_t));
        PG_CUDA(c, cudaMemcpyAsync(c->d_site.p, c->h_site.p, (size_t)n_reads * sizeof(int32_t), cudaMemcpyHostToDevice,
                                   c->stream));
        c->staging_busy = true;
    }
    c->n_reads = n_reads;
    c->max_len = maxl;
    c->bases_bytes = nb;
    c->have_bases_orig = false;
    c->imported = false;
    c->uploaded = true;
    c->ran = false;
    laps.lap("all");
    return PG_OK;
}

int pg_batch_run(pg_ctx* c, uint32_t flags)
{
    if (!c)
        return PG_E_ARG;
    if (!c->uploaded)
        return fail(c, PG_E_STATE, "pg_batch_run before pg_batch_upload");
    if (c->imported)
        return fail(c, PG_E_STATE, "pg_batch_run on an imported batch: pg_batch_import brings alignments, not reads");
    if (c->n_reads == 0)
    {
        c->ran = true;
        c->n_chunks_timed = 0;
        return PG_OK;
    }
    HostLaps laps("pg_batch_run");
    PG_CUDA(c, cudaSetDevice(c->device));
    int rc = upload_graphs(c);
    if (rc != PG_OK)
        return rc;
    laps.lap("upload_graphs");
    c->path_ran = c->kmer_ran = false;
    if (c->path_k > 0 || c->kmer_k > 0) // grm::CompositeAligner's stages in front of gssw (CompositeAligner.cpp:82-126)
    {
        rc = prepare_front_stages(c);
        if (rc != PG_OK)
            return rc;
        if (c->path_k > 0) // exact matches first
        {
            rc = run_path_stage(c);
            if (rc != PG_OK)
                return rc;
        }
        if (c->kmer_k > 0) // then gapless alignment to the graph's paths for the rest
        {
            rc = run_kmer_stage(c);
            if (rc != PG_OK)
                return rc;
        }
        if (!c->gssw_on)
        {
            c->n_chunks_timed = 0;
            c->ran = true;
            return PG_OK;
        }
    }
    else if (!c->gssw_on)
        return fail(c, PG_E_STATE, "pg_batch_run: no alignment stage enabled (pg_set_stages)");
    // geometry: W lanes per task, R rows per lane (W * R >= read length); see DESIGN.md "geometry"
    // reads over BYTE_MAX_READ_LEN can score past a byte (gssw's 16-bit mode): WIDE geometries, W = 32 only
    if (c->max_len > BYTE_MAX_READ_LEN)
        rc = c->max_len <= 320 ? run_chunks<10, 32>(c, flags)
            : (c->max_len <= 512 ? run_chunks<16, 32>(c, flags) : run_chunks<32, 32>(c, flags));
    else if (c->geom_w == 32)
        rc = c->max_len <= 160 ? run_chunks<5, 32>(c, flags) : run_chunks<8, 32>(c, flags);
    else if (c->geom_w == 8 && c->max_len <= 160)
        rc = run_chunks<20, 8>(c, flags);
    else
        rc = c->max_len <= 160 ? run_chunks<10, 16>(c, flags) : run_chunks<16, 16>(c, flags);
    laps.lap("stages in front + run_chunks");
    if (rc == PG_OK)
        c->ran = true;
    return rc;
}

int pg_batch_download(pg_ctx* c, pg_record* records, uint32_t* ops, uint64_t cap, uint64_t* used)
{
    struct Range
    {
        Range() { nvtxRangePushA("pg_batch_download"); }
        ~Range() { nvtxRangePop(); }
    } range;
    if (!c || (!records && c->n_reads > 0))
        return fail(c, PG_E_ARG, "pg_batch_download: bad arguments");
    if (!c->ran)
        return fail(c, PG_E_STATE, "pg_batch_download before pg_batch_run");
    if (c->n_reads == 0)
    {
        if (used)
            *used = 0;
        return PG_OK;
    }
    PG_CUDA(c, cudaSetDevice(c->device));
    const bool rec_direct = is_pinned(records);
    Record* hrec = reinterpret_cast<Record*>(records);
    if (!rec_direct)
    {
        PG_CUDA(c, c->h_records.reserve((size_t)c->n_reads));
        hrec = c->h_records.p;
    }
    PG_CUDA(c, c->h_cursor.reserve(1));
    PG_CUDA(c, cudaMemcpyAsync(hrec, c->d_records.p, (size_t)c->n_reads * sizeof(Record), cudaMemcpyDeviceToHost,
                               c->stream));
    PG_CUDA(c, cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                               c->stream));
    // the op words: how many there are is only known after the cursor has arrived, but a batch of short reads uses a few
    // words per read -- copy a guess of 8 per read in the same trip and fetch the rest (rare) afterwards
    const bool ops_direct = ops && is_pinned(ops);
    unsigned long long guess = 0;
    if (ops && cap > 0)
    {
        guess = std::min<unsigned long long>(std::min<unsigned long long>(cap, c->arena_cap), 8ull * (unsigned long long)c->n_reads);
        uint32_t* dst = ops;
        if (!ops_direct)
        {
            PG_CUDA(c, c->h_arena.reserve((size_t)std::max<unsigned long long>(guess, 1)));
            dst = c->h_arena.p;
        }
        if (guess)
            PG_CUDA(c, cudaMemcpyAsync(dst, c->d_arena.p, (size_t)guess * sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    }
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->staging_busy = false;
    if (getenv("PG_DEBUG_TIMELINE") && c->n_chunks_timed > 0) // where each chunk's phases sat on the device clock (ms from
    {                                                          // the first chunk's start): the two-stream overlap, in numbers
        for (int ci = 0; ci < c->n_chunks_timed; ++ci)
        {
            float t[4] = { 0, 0, 0, 0 };
            for (int x = 0; x < 4; ++x)
                cudaEventElapsedTime(&t[x], c->evpool[0], c->evpool[4 * ci + x]);
            fprintf(stderr, "[pg timeline] chunk %d on %s stream: fill %.3f .. %.3f ms, traceback %.3f .. %.3f ms\n", ci,
                    (c->split > 1 && c->aux_stream && (ci & 1)) ? "aux " : "main", t[0], t[1], t[2], t[3]);
        }
    }
    c->fill_ms = c->trace_ms = 0;
    for (int ci = 0; ci < c->n_chunks_timed; ++ci) // summed over the chunks of the batch
    {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, c->evpool[4 * ci], c->evpool[4 * ci + 1]);
        cudaEventElapsedTime(&b, c->evpool[4 * ci + 2], c->evpool[4 * ci + 3]);
        c->fill_ms += a;
        c->trace_ms += b;
    }
    unsigned long long n = *c->h_cursor.p;
    if (n > c->arena_cap)
        n = c->arena_cap;
    if (used)
        *used = n;
    if (!rec_direct)
        memcpy(records, hrec, (size_t)c->n_reads * sizeof(Record));
    for (int i = 0; i < c->n_reads; ++i)
        if (hrec[i].status == 2)
            return fail(c, PG_E_CAPACITY, "device cigar arena overflow at read " + std::to_string(i));
    if (n)
    {
        if (!ops || cap < n)
            return fail(c, PG_E_CAPACITY, "cigar arena too small: need " + std::to_string(n) + " ops");
        if (ops_direct)
        {
            if (n > guess) // the rest
            {
                PG_CUDA(c, cudaMemcpyAsync(ops + guess, c->d_arena.p + guess, (size_t)(n - guess) * sizeof(uint32_t),
                                           cudaMemcpyDeviceToHost, c->stream));
                PG_CUDA(c, cudaStreamSynchronize(c->stream));
            }
        }
        else
        {
            const unsigned long long have = std::min(n, guess);
            memcpy(ops, c->h_arena.p, (size_t)have * sizeof(uint32_t));
            if (n > guess)
            {
                PG_CUDA(c, c->h_arena.reserve((size_t)(n - guess)));
                PG_CUDA(c, cudaMemcpyAsync(c->h_arena.p, c->d_arena.p + guess, (size_t)(n - guess) * sizeof(uint32_t),
                                           cudaMemcpyDeviceToHost, c->stream));
                PG_CUDA(c, cudaStreamSynchronize(c->stream));
                memcpy(ops + guess, c->h_arena.p, (size_t)(n - guess) * sizeof(uint32_t));
            }
        }
    }
    return PG_OK;
}

int pg_align_batch(pg_ctx* c, int32_t n_reads, const char* bases, const int32_t* off, const int32_t* site,
                   uint32_t flags, pg_record* records, uint32_t* ops, uint64_t cap, uint64_t* used)
{
    int rc = pg_batch_upload(c, n_reads, bases, off, site);
    if (rc != PG_OK)
        return rc;
    rc = pg_batch_run(c, flags);
    if (rc != PG_OK)
        return rc;
    return pg_batch_download(c, records, ops, cap, used);
}

int pg_format_cigar(const pg_record* rec, const uint32_t* ops, char* out, int cap)
{
    if (!rec || (!ops && rec->cigar_len))
        return PG_E_ARG;
    Record r;
    memcpy(&r, rec, sizeof r);
    const std::string s = host::format_cigar(r, ops);
    if (out && cap > 0)
    {
        const size_t n = s.size() < (size_t)cap - 1 ? s.size() : (size_t)cap - 1;
        memcpy(out, s.data(), n);
        out[n] = 0;
    }
    return (int)s.size();
}

int pg_host_alloc(uint64_t bytes, void** out)
{
    if (!out || !bytes)
        return PG_E_ARG;
    *out = nullptr;
    return cudaMallocHost(out, (size_t)bytes) == cudaSuccess ? PG_OK : PG_E_CUDA;
}

void pg_host_free(void* p)
{
    if (p)
        cudaFreeHost(p);
}

int pg_batch_import(pg_ctx* c, int32_t n_reads, const int32_t* read_len, const int32_t* site, const pg_record* records,
                    const uint32_t* ops, uint64_t n_ops)
{
    if (!c || n_reads < 0 || (n_reads > 0 && (!read_len || !records)) || (n_ops > 0 && !ops))
        return fail(c, PG_E_ARG, "pg_batch_import: bad arguments");
    c->uploaded = false;
    c->ran = false;
    c->n_reads = 0;
    c->imported = true;
    c->have_bases_orig = false;
    if (n_reads == 0)
    {
        c->uploaded = c->ran = true;
        c->n_chunks_timed = 0;
        return PG_OK;
    }
    const int nsites = (int)c->graphs.sites.size();
    if (nsites == 0)
        return fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    PG_CUDA(c, cudaSetDevice(c->device));
    if (c->staging_busy)
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
    c->staging_busy = false;
    PG_CUDA(c, c->h_off.reserve((size_t)n_reads + 1));
    int maxl = 0;
    c->h_off.p[0] = 0;
    for (int i = 0; i < n_reads; ++i)
    {
        if (read_len[i] <= 0 || (int64_t)c->h_off.p[i] + read_len[i] > 0x7FFFFFFF)
            return fail(c, PG_E_READ_LEN, "pg_batch_import: bad length of read " + std::to_string(i));
        if (site && (site[i] < 0 || site[i] >= nsites))
            return fail(c, PG_E_ARG, "read " + std::to_string(i) + " names unknown site " + std::to_string(site[i]));
        if ((uint64_t)records[i].cigar_off + records[i].cigar_len > n_ops)
            return fail(c, PG_E_ARG, "pg_batch_import: record " + std::to_string(i) + " points outside cigar_ops");
        c->h_off.p[i + 1] = c->h_off.p[i] + read_len[i];
        maxl = std::max(maxl, (int)read_len[i]);
    }
    PG_CUDA(c, c->d_off.reserve((size_t)n_reads + 1));
    PG_CUDA(c, c->d_records.reserve((size_t)n_reads));
    PG_CUDA(c, c->d_arena.reserve((size_t)n_ops + 1));
    PG_CUDA(c, c->d_cursor.reserve(1));
    PG_CUDA(c, c->h_cursor.reserve(1));
    PG_CUDA(c, cudaMemcpyAsync(c->d_off.p, c->h_off.p, ((size_t)n_reads + 1) * sizeof(int32_t), cudaMemcpyHostToDevice,
                               c->stream));
    PG_CUDA(c, cudaMemcpyAsync(c->d_records.p, records, (size_t)n_reads * sizeof(Record), cudaMemcpyHostToDevice, c->stream));
    if (n_ops)
        PG_CUDA(c, cudaMemcpyAsync(c->d_arena.p, ops, (size_t)n_ops * sizeof(uint32_t), cudaMemcpyHostToDevice, c->stream));
    *c->h_cursor.p = n_ops;
    PG_CUDA(c, cudaMemcpyAsync(c->d_cursor.p, c->h_cursor.p, sizeof(unsigned long long), cudaMemcpyHostToDevice, c->stream));
    c->have_sites = site != nullptr;
    if (site)
    {
        PG_CUDA(c, c->h_site.reserve((size_t)n_reads));
        PG_CUDA(c, c->d_site.reserve((size_t)n_reads));
        memcpy(c->h_site.p, site, (size_t)n_reads * sizeof(int32_t));
        PG_CUDA(c, cudaMemcpyAsync(c->d_site.p, c->h_site.p, (size_t)n_reads * sizeof(int32_t), cudaMemcpyHostToDevice,
                                   c->stream));
    }
    PG_CUDA(c, cudaStreamSynchronize(c->stream)); // caller buffers may be pageable and reused right away
    c->arena_cap = n_ops;
    c->n_reads = n_reads;
    c->max_len = maxl;
    c->uploaded = true;
    c->ran = true;
    c->n_chunks_timed = 0;
    return PG_OK;
}

int pg_set_edge_labels(pg_ctx* c, int32_t site, const uint64_t* masks)
{
    if (!c)
        return PG_E_ARG;
    if (site < 0 || (size_t)site >= c->graphs.sites.size())
        return fail(c, PG_E_ARG, "pg_set_edge_labels: unknown site " + std::to_string(site));
    const int64_t eb = c->graphs.edge_base[(size_t)site];
    const int n = c->graphs.sites[(size_t)site].n_edges;
    for (int e = 0; e < n; ++e)
        c->graphs.in_label[(size_t)(eb + e)] = masks ? masks[e] : 0ull;
    c->count_dirty = true;
    return PG_OK;
}

int pg_batch_count(pg_ctx* c, const int32_t* fragment, const uint8_t* is_reverse_strand, const pg_count_params* params,
                   pg_read_support* support, uint32_t* path_words, uint64_t path_cap, uint64_t* path_used,
                   pg_count4* node_counts, uint64_t node_cap, pg_count4* edge_counts, uint64_t edge_cap,
                   uint32_t* family_words, uint64_t family_cap, uint64_t* family_used)
{
    static_assert(sizeof(pg_read_support) == sizeof(ReadSupport) && sizeof(pg_count4) == sizeof(Count4), "ABI structs");
    if (!c || !params)
        return fail(c, PG_E_ARG, "pg_batch_count: bad arguments");
    if (!c->ran)
        return fail(c, PG_E_STATE, "pg_batch_count before pg_batch_run");
    const host::GraphStore& gs = c->graphs;
    const size_t ns = gs.sites.size();
    if (ns == 0)
        return fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    const uint64_t total_nodes = (uint64_t)gs.node_base[ns], total_edges = (uint64_t)gs.edge_base[ns];
    if ((node_counts && node_cap < total_nodes) || (edge_counts && edge_cap < total_edges))
        return fail(c, PG_E_CAPACITY, "pg_batch_count: need " + std::to_string(total_nodes) + " node rows and "
                        + std::to_string(total_edges) + " edge rows");
    CountParams prm;
    prm.remove_nonuniq = params->remove_nonuniq;
    prm.use_support_filters = params->use_support_filters;
    prm.bad_align_frac = params->bad_align_frac;
    prm.family_slots = params->family_slots > 0 ? params->family_slots : 256;
    if (path_used)
        *path_used = 0;
    if (family_used)
        *family_used = 0;
    if (node_counts)
        memset(node_counts, 0, (size_t)total_nodes * sizeof(pg_count4));
    if (edge_counts)
        memset(edge_counts, 0, (size_t)total_edges * sizeof(pg_count4));
    if (c->n_reads == 0)
        return PG_OK;
    PG_CUDA(c, cudaSetDevice(c->device));
    const int n = c->n_reads;

    // fragments: chain the reads of a fragment in input order; the first one accumulates
    std::vector<int32_t> next;
    std::vector<uint8_t> head;
    {
        std::string err;
        if (!host::build_fragment_chains(fragment, c->have_sites ? c->h_site.p : nullptr, n, next, head, err))
            return fail(c, PG_E_ARG, "pg_batch_count: " + err);
    }
    int rc = upload_graphs(c);
    if (rc == PG_OK)
        rc = upload_count_tables(c, prm.family_slots);
    if (rc != PG_OK)
        return rc;

    // the op count of the batch bounds the path words
    PG_CUDA(c, c->h_cursor.reserve(1));
    PG_CUDA(c, cudaMemcpyAsync(c->h_cursor.p, c->d_cursor.p, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    const unsigned long long n_ops = *c->h_cursor.p;
    if (n_ops > c->arena_cap)
        return fail(c, PG_E_CAPACITY, "device cigar arena overflowed in the last run");
    if (path_words && path_cap < n_ops)
        return fail(c, PG_E_CAPACITY, "pg_batch_count: path_words needs " + std::to_string(n_ops) + " words");

    PG_CUDA(c, put(c, c->d_next, next));
    PG_CUDA(c, put(c, c->d_head, head));
    if (is_reverse_strand)
    {
        PG_CUDA(c, c->d_isrev.reserve((size_t)n));
        PG_CUDA(c, cudaMemcpyAsync(c->d_isrev.p, is_reverse_strand, (size_t)n, cudaMemcpyHostToDevice, c->stream));
    }
    PG_CUDA(c, c->d_support.reserve((size_t)n));
    PG_CUDA(c, c->d_path.reserve((size_t)n_ops + 1));
    PG_CUDA(c, c->d_node_counts.reserve((size_t)total_nodes + 1));
    PG_CUDA(c, c->d_edge_counts.reserve((size_t)total_edges + 1));
    PG_CUDA(c, c->d_fam_counts.reserve((size_t)c->fam_rows + 1));
    PG_CUDA(c, c->d_fam_keys.reserve((size_t)c->fam_nkeys + 1));
    PG_CUDA(c, c->d_count_cursor.reserve(2));
    const unsigned long long fam_out_cap = 4ull * (unsigned long long)c->fam_rows + 4ull * (unsigned long long)c->fam_nkeys;
    PG_CUDA(c, c->d_fam_out.reserve((size_t)fam_out_cap + 1));
    if (!c->count_ev[0])
    {
        PG_CUDA(c, cudaEventCreate(&c->count_ev[0]));
        PG_CUDA(c, cudaEventCreate(&c->count_ev[1]));
    }
    PG_CUDA(c, cudaEventRecord(c->count_ev[0], c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_node_counts.p, 0, ((size_t)total_nodes + 1) * sizeof(Count4), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_edge_counts.p, 0, ((size_t)total_edges + 1) * sizeof(Count4), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_fam_counts.p, 0, ((size_t)c->fam_rows + 1) * sizeof(Count4), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_fam_keys.p, 0, ((size_t)c->fam_nkeys + 1) * sizeof(unsigned long long), c->stream));
    PG_CUDA(c, cudaMemsetAsync(c->d_count_cursor.p, 0, 2 * sizeof(unsigned long long), c->stream));

    CountArgs a;
    a.t.sites = c->d_sites.p;
    a.t.gints = c->d_gints.p;
    a.t.csite = c->d_csite.p;
    a.t.csr_input = c->d_csr_input.p;
    a.t.lab_edge = c->d_lab_edge.p;
    a.t.lab_out = c->d_lab_out.p;
    a.t.lab_in = c->d_lab_in.p;
    a.prm = prm;
    a.records = c->d_records.p;
    a.ops = c->d_arena.p;
    a.read_off = c->d_off.p;
    a.read_site = c->have_sites ? c->d_site.p : nullptr;
    a.is_rev = is_reverse_strand ? c->d_isrev.p : nullptr;
    a.n_reads = n;
    a.sup = c->d_support.p;
    a.path = c->d_path.p;
    a.next = c->d_next.p;
    a.head = c->d_head.p;
    a.node_counts = c->d_node_counts.p;
    a.edge_counts = c->d_edge_counts.p;
    a.fam_counts = c->d_fam_counts.p;
    a.fam_keys = c->d_fam_keys.p;
    a.cursor = c->d_count_cursor.p;
    a.fam_out = c->d_fam_out.p;
    a.fam_out_cap = fam_out_cap;
    a.n_sites = (int)ns;
    const int grid = (n + 127) / 128;
    pg_support_kernel<<<grid, 128, 0, c->stream>>>(a);
    PG_CUDA(c, cudaGetLastError());
    pg_fragment_kernel<<<grid, 128, 0, c->stream>>>(a);
    PG_CUDA(c, cudaGetLastError());
    pg_family_compact_kernel<<<(unsigned)((ns * 32 + 127) / 128), 128, 0, c->stream>>>(a);
    PG_CUDA(c, cudaGetLastError());
    c->launches += 3;
    c->count_launches += 3;
    PG_CUDA(c, cudaEventRecord(c->count_ev[1], c->stream));

    unsigned long long cur[2] = { 0, 0 };
    PG_CUDA(c, cudaMemcpyAsync(cur, c->d_count_cursor.p, sizeof cur, cudaMemcpyDeviceToHost, c->stream));
    if (support)
        PG_CUDA(c, get(c, support, c->d_support.p, (size_t)n * sizeof(ReadSupport)));
    if (path_words)
        PG_CUDA(c, get(c, path_words, c->d_path.p, (size_t)n_ops * sizeof(uint32_t)));
    if (node_counts)
        PG_CUDA(c, get(c, node_counts, c->d_node_counts.p, (size_t)total_nodes * sizeof(Count4)));
    if (edge_counts)
        PG_CUDA(c, get(c, edge_counts, c->d_edge_counts.p, (size_t)total_edges * sizeof(Count4)));
    PG_CUDA(c, cudaStreamSynchronize(c->stream));
    cudaEventElapsedTime(&c->count_ms, c->count_ev[0], c->count_ev[1]);
    if (path_used)
        *path_used = n_ops;
    if (cur[1])
        return fail(c, PG_E_CAPACITY, "pg_batch_count: site " + std::to_string(cur[1] - 1) + " has more than "
                        + std::to_string(prm.family_slots) + " distinct path-family sets (raise family_slots)");
    if (family_used)
        *family_used = cur[0];
    if (cur[0])
    {
        if (!family_words || family_cap < cur[0])
            return fail(c, PG_E_CAPACITY, "pg_batch_count: family_words needs " + std::to_string(cur[0]) + " words");
        PG_CUDA(c, get(c, family_words, c->d_fam_out.p, (size_t)cur[0] * sizeof(uint32_t)));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    return PG_OK;
}

int pg_count_stats(const pg_ctx* c, uint64_t* launches, float* ms)
{
    if (!c)
        return PG_E_ARG;
    if (launches)
        *launches = c->count_launches;
    if (ms)
        *ms = c->count_ms;
    return PG_OK;
}

int pg_set_stages(pg_ctx* c, int32_t path_kmer_len, int32_t graph_matching, int32_t nonuniq_second_chance)
{
    if (!c || path_kmer_len < 0 || path_kmer_len > 4096)
        return fail(c, PG_E_ARG, "pg_set_stages: bad k-mer length");
    if (path_kmer_len == 0 && !graph_matching && c->kmer_k == 0)
        return fail(c, PG_E_ARG, "pg_set_stages: no alignment stage enabled");
    if (path_kmer_len != c->path_k)
        c->path_dirty = true;
    c->path_k = path_kmer_len;
    c->gssw_on = graph_matching != 0;
    c->path_second_chance = nonuniq_second_chance != 0;
    return PG_OK;
}

int pg_path_stats(pg_ctx* c, uint64_t* counters4, float* path_ms)
{
    if (!c)
        return PG_E_ARG;
    if (c->path_ran)
    {
        PG_CUDA(c, cudaSetDevice(c->device));
        PG_CUDA(c, cudaMemcpyAsync(c->path_counters, c->d_pcount.p, 3 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        c->path_counters[0] = (unsigned long long)c->n_reads; // attempted: every read of the batch
        PG_CUDA(c, cudaEventElapsedTime(&c->path_ms, c->path_ev[0], c->path_ev[1]));
    }
    else
    {
        c->path_counters[0] = c->path_counters[1] = c->path_counters[2] = 0;
        c->path_ms = 0;
    }
    if (counters4)
    {
        for (int i = 0; i < 3; ++i)
            counters4[i] = c->path_counters[i];
        counters4[3] = c->path_index_us;
    }
    if (path_ms)
        *path_ms = c->path_ms;
    return PG_OK;
}

int pg_set_paths(pg_ctx* c, int32_t site, int32_t n_paths, const int32_t* path_ptr, const int32_t* path_nodes)
{
    if (!c)
        return PG_E_ARG;
    std::string err;
    if (!c->graphs.set_paths(site, n_paths, path_ptr, path_nodes, err))
        return fail(c, PG_E_GRAPH, err);
    c->kmer_dirty = true;
    return PG_OK;
}

int pg_set_kmer_stage(pg_ctx* c, int32_t kmer_len)
{
    if (!c || kmer_len < 0 || kmer_len == 1 || kmer_len > 16)
        return fail(c, PG_E_ARG, "pg_set_kmer_stage: k-mer length must be 0 (off) or 2..16");
    if (kmer_len != c->kmer_k)
        c->kmer_dirty = true;
    c->kmer_k = kmer_len;
    return PG_OK;
}

int pg_kmer_stats(pg_ctx* c, uint64_t* counters2, float* kmer_ms)
{
    if (!c)
        return PG_E_ARG;
    if (c->kmer_ran)
    {
        PG_CUDA(c, cudaSetDevice(c->device));
        PG_CUDA(c, cudaMemcpyAsync(c->kmer_counters, c->d_kcount.p, 2 * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                                   c->stream));
        PG_CUDA(c, cudaStreamSynchronize(c->stream));
        cudaEventElapsedTime(&c->kmer_ms, c->kmer_ev[0], c->kmer_ev[1]);
    }
    else
    {
        c->kmer_counters[0] = c->kmer_counters[1] = 0;
        c->kmer_ms = 0;
    }
    if (counters2)
    {
        counters2[0] = c->kmer_counters[0];
        counters2[1] = c->kmer_counters[1];
    }
    if (kmer_ms)
        *kmer_ms = c->kmer_ms;
    return PG_OK;
}

int pg_stats(const pg_ctx* c, uint64_t* launches, float* fill_ms, float* trace_ms)
{
    if (!c)
        return PG_E_ARG;
    if (launches)
        *launches = c->launches;
    if (fill_ms)
        *fill_ms = c->fill_ms;
    if (trace_ms)
        *trace_ms = c->trace_ms;
    return PG_OK;
}
}
