/* paragraph_b200 -- C-ABI of the B200 read->graph alignment engine.
 *
 * Drop-in boundary: this library replaces what paragraph's grm::GraphAligner does per read
 *   (src/c++/lib/grm/GraphAligner.cpp:277-285 setGraph, :308-404 alignRead = 4 x gssw_graph_fill +
 *    gssw_graph_trace_back + alignsEndAtMultNodes + strand choice; external/gssw/gssw.h:235-715),
 * batch-first, behind grm::alignReads (src/c++/include/grm/Align.hh:49-52).  Plain C types, caller-owned
 * host buffers, no exceptions; every call returns a status and pg_last_error() explains failures.
 * The host-side C++ mirror of grm::GraphAligner / CompositeAligner / alignReads that sits on top of this
 * ABI is paragraph_b200/csrc/host/pg_grm.hh; INTEGRATION.md shows the binding a maintainer adds.
 *
 * There is no CPU fallback: every entry point fails with PG_E_CUDA when no sm_100 device is usable.
 */
#ifndef PG_ALIGN_H
#define PG_ALIGN_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PG_OK 0
#define PG_E_ARG (-1)      /* bad argument (message in pg_last_error) */
#define PG_E_CUDA (-2)     /* CUDA runtime / device error */
#define PG_E_READ_LEN (-3) /* a read is empty or longer than PG_MAX_READ_LEN */
#define PG_E_GRAPH (-4)    /* graph not topologically ordered / empty node / too large */
#define PG_E_CAPACITY (-5) /* caller's cigar arena too small */
#define PG_E_STATE (-6)    /* call order (e.g. run before upload) */

/* Reads longer than this leave gssw's 8-bit mode (external/gssw/gssw.c:380, :4001-4013) whose 16-bit
 * fallback has different tie behaviour; they are rejected rather than answered differently. */
#define PG_MAX_READ_LEN 250

/* GraphAligner alignment flags (src/c++/include/grm/GraphAligner.hh:64-67) */
#define PG_AF_CIGAR 0x01u
#define PG_AF_BOTH_STRANDS 0x02u
#define PG_AF_REVERSE_GRAPH 0x04u
#define PG_AF_ALL 0xFFFFFFFFu

typedef struct pg_ctx pg_ctx;

/* One record per read, in input order.  Replaces the fields GraphAligner::alignRead writes into
 * common::Read (src/c++/include/common/Read.hh:97-108; GraphAligner.cpp:358-401):
 *   graph_pos, graph_alignment_score, is_graph_alignment_unique (graph_mapq = unique ? 60 : 0),
 *   chose_reverse (=> bases := reverseComplement(bases), quals reversed, and
 *   is_graph_reverse_strand = is_reverse_strand != chose_reverse), graph_cigar = ops[cigar_off .. +cigar_len). */
typedef struct pg_record
{
    int32_t graph_pos;
    int32_t score;
    uint8_t unique;
    uint8_t chose_reverse;
    uint8_t status; /* 0 ok; 1 traceback dead end (the reference would assert/spin); 2 op log overflow */
    uint8_t query_clipped; /* soft-clipped query bases = sum of the S ops; readfilters::BadAlign
                              (src/c++/lib/paragraph/readfilters/BadAlign.hh:62-73) filters a read when
                              read_len - query_clipped < round(bad_align_frac * read_len) -- no CIGAR decode needed */
    uint32_t cigar_off;
    uint32_t cigar_len;
} pg_record;

/* A graph CIGAR op: node id << 16 | run length << 3 | op, ops in path order, run-length merged per node
 * exactly like gssw_cigar_push_back/_front (gssw.c:3679-3700).  extractCigar (GraphAligner.cpp:88-108)
 * prints them as "<node>[<len><op>...]...". */
#define PG_OP_M 0
#define PG_OP_X 1
#define PG_OP_N 2
#define PG_OP_I 3
#define PG_OP_D 4
#define PG_OP_S 5
#define PG_OP_NONE 7 /* length 0: the path touches this node without an op (prints as "id[]"), see pg_core.cuh */
#define PG_CIGAR_NODE(w) ((uint32_t)(w) >> 16)
#define PG_CIGAR_LEN(w) (((uint32_t)(w) >> 3) & 0x1FFFu)
#define PG_CIGAR_OP(w) ((uint32_t)(w) & 7u)

/* ---- context ------------------------------------------------------------------------------- */
/* One context per host thread / CUDA stream (grm::alignReads builds one aligner per chunk,
 * src/c++/lib/grm/Align.cpp:107-110).  `device` is a CUDA ordinal. */
int pg_create(int device, pg_ctx** out);
void pg_destroy(pg_ctx* ctx);
const char* pg_last_error(const pg_ctx* ctx);
/* Launch on this cudaStream_t (e.g. torch's current stream); NULL = the context's own stream. */
int pg_set_stream(pg_ctx* ctx, void* cuda_stream);
/* Upper bound for per-batch device scratch (checkpoints); batches are processed in chunks under it. */
int pg_set_scratch_limit(pg_ctx* ctx, uint64_t bytes);

/* ---- graphs (sites) -------------------------------------------------------------------------- */
/* Replaces GraphAligner::setGraph -> GraphAlignerImpl::initializeGraph for the graph and its reverse
 * (GraphAligner.cpp:110-167, 277-285).  Node ids must be topologically ordered (efrom < eto), sequences
 * non-empty; sequences are upper-cased; predecessor order is ascending id.  Returns the site id. */
int pg_add_graph(pg_ctx* ctx, int32_t n_nodes, const char* seq_blob, const int32_t* seq_off, int32_t n_edges,
                 const int32_t* efrom, const int32_t* eto, int32_t* site_id);
int pg_clear_graphs(pg_ctx* ctx);

/* ---- alignment -------------------------------------------------------------------------------- */
/* Replaces the loop `for read: GraphAligner::alignRead(read, flags)` (Align.cpp:72-84 with only the gssw
 * stage enabled).  read_site may be NULL (all reads on site 0).  Host buffers in, host buffers out;
 * records[n_reads], cigar_ops[cigar_cap] are written; *cigar_used = ops written. */
int pg_align_batch(pg_ctx* ctx, int32_t n_reads, const char* bases_blob, const int32_t* read_off,
                   const int32_t* read_site, uint32_t flags, pg_record* records, uint32_t* cigar_ops,
                   uint64_t cigar_cap, uint64_t* cigar_used);

/* The same in three stages, so that a caller (bench.py) can time the kernels with inputs resident in HBM:
 * upload = H2D of the reads, run = the kernels only (asynchronous on the stream), download = D2H + sync. */
int pg_batch_upload(pg_ctx* ctx, int32_t n_reads, const char* bases_blob, const int32_t* read_off,
                    const int32_t* read_site);
int pg_batch_run(pg_ctx* ctx, uint32_t flags);
int pg_batch_download(pg_ctx* ctx, pg_record* records, uint32_t* cigar_ops, uint64_t cigar_cap, uint64_t* cigar_used);

/* ---- page-locked host buffers (optional) -------------------------------------------------------- */
/* Reads / records / CIGAR buffers allocated here (or otherwise page-locked and known to CUDA) are copied to and from
 * the device directly; pageable buffers are staged through the context's own pinned buffers (one extra memcpy). */
int pg_host_alloc(uint64_t bytes, void** out);
void pg_host_free(void* p);

/* ---- helpers ----------------------------------------------------------------------------------- */
/* "<node>[<len><op>...]..." into out (NUL terminated); returns the string length (may exceed cap). */
int pg_format_cigar(const pg_record* rec, const uint32_t* cigar_ops, char* out, int cap);
/* Kernels launched by this context so far, and the last batch's per-kernel device time in ms
 * (fill, traceback) measured with CUDA events on the launching stream. */
int pg_stats(const pg_ctx* ctx, uint64_t* kernel_launches, float* last_fill_ms, float* last_trace_ms);
/* Library / build description, e.g. "paragraph_b200 sm_100a CK=16". */
const char* pg_version(void);

#ifdef __cplusplus
}
#endif
#endif
