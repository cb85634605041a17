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

/* Longest read (32 rows per lane x 32 lanes).  Reads whose score reaches 251 take gssw's 16-bit mode in the
 * reference (external/gssw/gssw.c:380, 527-786, 4001-4013); results stay bit-identical, including what
 * GraphAligner's uniqueness scan makes of a 16-bit matrix (src/c++/lib/grm/GraphAligner.cpp:177-186). */
#define PG_MAX_READ_LEN 1024

/* pg_record::mapped_by: the stage of grm::CompositeAligner that mapped the read */
#define PG_STAGE_GSSW_ID 0
#define PG_STAGE_PATH_ID 1
#define PG_STAGE_GSSW_REV_ID 2 /* gssw, after ONE earlier stage had reverse-complemented the bases (second chance) */
#define PG_STAGE_KMER_ID 3     /* grm::KmerAligner (pg_set_kmer_stage) */
#define PG_STAGE_KMER_REV_ID 4 /* KmerAligner, after the exact-match stage had reverse-complemented the bases */
#define PG_STAGE_GSSW_REV2_ID 5 /* gssw, after both earlier stages had reverse-complemented the bases */

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
    int16_t score;          /* gssw_graph_mapping::score is an int16_t as well */
    uint16_t query_clipped; /* soft-clipped query bases = sum of the S ops; readfilters::BadAlign
                               (src/c++/lib/paragraph/readfilters/BadAlign.hh:62-73) filters a read when
                               read_len - query_clipped < round(bad_align_frac * read_len) -- no CIGAR decode needed */
    uint8_t unique;
    uint8_t chose_reverse;
    uint8_t status; /* 0 ok; 1 traceback dead end (the reference would assert/spin); 2 op log overflow;
                       3 unmapped: no enabled stage mapped the read (graph_mapping_status stays UNMAPPED) */
    uint8_t mapped_by; /* stage of grm::CompositeAligner that mapped the read (PG_STAGE_*_ID).  The stages set the
                          strand differently: gssw -> is_graph_reverse_strand = is_reverse_strand != chose_reverse,
                          bases reverse-complemented and quals reversed if chose_reverse (GraphAligner.cpp:358-378);
                          PathAligner -> is_graph_reverse_strand = chose_reverse, bases reverse-complemented, quals
                          untouched (PathAligner.cpp:124-135); KmerAligner -> is_graph_reverse_strand =
                          is_reverse_strand != chose_reverse, bases reverse-complemented, quals untouched
                          (KmerAligner.cpp:452-466); the _REV / _REV2 ids -> the same, applied to bases that one / two
                          earlier stages had already reverse-complemented (graphtools::reverseComplement each time) */
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
/* Upper bound for per-batch device scratch (checkpoints); batches are processed in chunks under it.
   Default: 64 GiB, or 40 % of the device memory if that is less. */
int pg_set_scratch_limit(pg_ctx* ctx, uint64_t bytes);

/* ---- graphs (sites) -------------------------------------------------------------------------- */
/* Replaces GraphAligner::setGraph -> GraphAlignerImpl::initializeGraph for the graph and its reverse
 * (GraphAligner.cpp:110-167, 277-285).  Node ids must be topologically ordered (efrom < eto), sequences
 * non-empty; sequences are upper-cased; predecessor order is ascending id.  Returns the site id. */
int pg_add_graph(pg_ctx* ctx, int32_t n_nodes, const char* seq_blob, const int32_t* seq_off, int32_t n_edges,
                 const int32_t* efrom, const int32_t* eto, int32_t* site_id);
/* Many sites in one call -- what grmpy's workflow hands out one (sample, graph) pair at a time
 * (src/c++/lib/grmpy/Workflow.cpp:108-146; every alignSingleSample loads its graph, AlignSamples.cpp:115-130).
 * Graph s has the nodes [node_ptr[s], node_ptr[s+1]) of seq_off (n_nodes_total + 1 offsets into seq_blob) and the
 * edges [edge_ptr[s], edge_ptr[s+1]) of efrom / eto (node ids local to the site).  Same rules and errors as
 * pg_add_graph; the sites get consecutive ids starting at *first_site_id.  Nothing is registered if one fails. */
int pg_add_graphs(pg_ctx* ctx, int32_t n_sites, const int32_t* node_ptr, const char* seq_blob, const int32_t* seq_off,
                  const int32_t* edge_ptr, const int32_t* efrom, const int32_t* eto, int32_t* first_site_id);
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
/* ------------------------------------------------------------------------------------------------------------
 * Counting stage: what paragraph::alignAndDisambiguate does with the aligned reads of a site
 * after grm::alignReads -- the read filter chain, disambiguateReads and countReads -- computed on the device from
 * the op words of the last pg_batch_run, so that a caller that only needs counts (grmpy: AlignSamples.cpp:124-127)
 * never downloads or parses a CIGAR.  Replaces
 *   createReadFilter: NonUniq, BadAlign                src/c++/lib/paragraph/ReadFilter.cpp:73-90
 *   disambiguateReads + node/edge support filters      src/c++/lib/paragraph/Disambiguation.cpp:82-142, 212-296
 *   readsToFragments / Fragment::addRead counters      src/c++/lib/common/Fragment.cpp:33-67, 141-182
 *   countNodes / countEdges / countPathFamilies        src/c++/lib/paragraph/ReadCounting.cpp:52-127
 * ------------------------------------------------------------------------------------------------------------ */

/* Path-family labels of a site's edges (the "sequences" of the graph JSON's edges; Graph::addLabelToEdge,
 * GraphInput.cpp:137-147): one 64-bit mask per edge in the order the edges were given to pg_add_graph, bit k =
 * label k (the caller keeps the k -> name table; at most 64 labels per site).  NULL clears the site's labels. */
int pg_set_edge_labels(pg_ctx* ctx, int32_t site, const uint64_t* edge_label_mask);

#define PG_V_MAPPED 0    /* passed the filters: takes part in the counts (Read::MAPPED) */
#define PG_V_NONUNIQ 1   /* removed by readfilters::NonUniq ("nonuniq") */
#define PG_V_BAD_ALIGN 2 /* removed by readfilters::BadAlign ("bad_align") */
#define PG_V_INVALID 3   /* decodeGraphAlignment would throw on this CIGAR (empty alignment of a score-0 read that
                            NonUniq did not remove, or gssw's 'U' quirk): the reference aborts the site; here the read
                            is reported and left out of the counts */

#define PG_SUP_NODE_MASK 0xFFFFu
#define PG_SUP_NODE 0x40000000u /* path word: the read supports this node (graph_nodes_supported) */
#define PG_SUP_EDGE 0x80000000u /* path word: the read supports the edge previous path node -> this node */

typedef struct pg_read_support
{
    uint64_t sequences;    /* bit k set: label k is in graph_sequences_supported */
    uint32_t path_off;     /* first path word of the read ( = its record's cigar_off) */
    uint16_t path_len;     /* number of path nodes; 0 unless verdict == PG_V_MAPPED */
    uint8_t verdict;       /* PG_V_* */
    uint8_t graph_reverse; /* is_graph_reverse_strand = is_reverse_strand != chose_reverse (GraphAligner.cpp:358-359) */
} pg_read_support;

typedef struct pg_count4
{
    uint32_t fragments, reads, fwd, rev; /* JSON "<name>", "<name>:READS", ":FWD", ":REV" (ReadCounting.cpp:52-69) */
} pg_count4;

typedef struct pg_count_params
{
    int32_t remove_nonuniq;      /* Parameters::remove_nonuniq_reads (default 1) */
    int32_t use_support_filters; /* 1: the node/edge filters alignAndDisambiguate passes to disambiguateReads;
                                    0: null filters, as the reference's unit tests call disambiguateReads */
    double bad_align_frac;       /* Parameters::bad_align_frac (default 0.8) */
    int32_t family_slots;        /* most distinct label sets per site the device table holds (a site with L labels
                                    gets min(family_slots, 2^L) slots); 0 = 256 */
    int32_t reserved;
} pg_count_params;

/* Instead of pg_batch_upload + pg_batch_run: load alignments made elsewhere (the other stages of the reference's
 * CompositeAligner cascade -- PathAligner / KmerAligner / KlibAligner -- or a previous run) as the context's current
 * batch, so that pg_batch_count can filter, disambiguate and count them.  records[i].cigar_off / cigar_len index
 * cigar_ops (n_ops words, encoding above); read_len[i] is the read's length (common::Read::bases().size());
 * read_site as in pg_batch_upload. */
int pg_batch_import(pg_ctx* ctx, int32_t n_reads, const int32_t* read_len, const int32_t* read_site,
                    const pg_record* records, const uint32_t* cigar_ops, uint64_t n_ops);

/* Run the counting stage on the batch last executed by pg_batch_run / pg_align_batch on this context.
 *   fragment          [n_reads] fragment of each read: reads of one fragment (mates) share a value >= 0; a fragment
 *                     must not span sites.  NULL = every read is its own fragment.
 *   is_reverse_strand [n_reads] the read's BAM strand (common::Read::is_reverse_strand) or NULL (all forward)
 *   support           [n_reads] or NULL
 *   path_words        [path_cap] or NULL: path words of read i at path_words[support[i].path_off ..]; path_cap must
 *                     be >= the op count pg_batch_download reports; *path_used receives it
 *   node_counts       [sum of n_nodes over all registered sites], site-major in pg_add_graph order, node order
 *   edge_counts       [sum of n_edges ...], edges in the order given to pg_add_graph
 *   family_words      read_counts_by_sequence: one entry per (site, distinct non-empty label set) in no particular
 *                     order: {site, n, mask_lo, mask_hi, n x pg_count4} with n = 1 + n_nodes + n_edges and the rows
 *                     "total", nodes, edges (DETAILED_READ_COUNTS); *family_used = words written
 * PG_E_CAPACITY: an output buffer is too small (sizes in pg_last_error) or a site has more than family_slots label
 * sets. */
int pg_batch_count(pg_ctx* ctx, const int32_t* fragment, const uint8_t* is_reverse_strand,
                   const pg_count_params* params, pg_read_support* support, uint32_t* path_words, uint64_t path_cap,
                   uint64_t* path_used, pg_count4* node_counts, uint64_t node_cap, pg_count4* edge_counts,
                   uint64_t edge_cap, uint32_t* family_words, uint64_t family_cap, uint64_t* family_used);

/* Counting-stage kernels launched by this context so far and the device time in ms of the last pg_batch_count
 * (memsets + the three kernels, CUDA events on the launching stream). */
int pg_count_stats(const pg_ctx* ctx, uint64_t* kernel_launches, float* last_count_ms);

/* ---- stages of the cascade --------------------------------------------------------------------------------
 * grm::CompositeAligner(pathMatching, graphMatching, klib, kmer) (src/c++/include/grm/CompositeAligner.hh:44-67,
 * lib/grm/CompositeAligner.cpp:78-176) tries its enabled stages in order; the first that maps a read wins.
 *   path_kmer_len > 0 : exact-match stage first = grm::PathAligner(kmer_size) (include/grm/PathAligner.hh:39-78,
 *                       lib/grm/PathAligner.cpp:75-164) with the graphtools::KmerIndex it builds in setGraph (32 in
 *                       the reference, `paragraph` switches it on by default: src/c++/main/paragraph.cpp:60).  A read
 *                       with a full-length exact match on either strand gets score = read length, graph_pos = offset
 *                       in the first node, ops "<overlap>M" per node, unique = 0 iff a second full-length match
 *                       exists, mapped_by = PG_STAGE_PATH_ID.  The index of every registered graph is (re)built at
 *                       the next pg_batch_run.
 *   graph_matching    : the gssw stage (the DP kernels) for the reads still unmapped; when 0 they stay unmapped
 *                       (pg_record::status 3).
 *   nonuniq_second_chance : the cascade runs the caller's read filter right after the exact-match stage and hands a
 *                       rejected read to the later stages (CompositeAligner.cpp:97-103).  With paragraph's default
 *                       filter chain that happens exactly to NON-UNIQUE exact matches (NonUniq, ReadFilter.cpp:79-83;
 *                       BadAlign never rejects an unclipped match).  1 = do that on the device: such a read goes to
 *                       the DP with the bases PathAligner left behind (reverse-complemented if its first match was on
 *                       the reverse strand; mapped_by = PG_STAGE_GSSW_REV_ID then).  0 = report the non-unique exact
 *                       match itself (what the cascade does with a null filter); a host adapter with an arbitrary
 *                       filter callback re-submits rejected reads itself (pg_grm.hh).
 * Default: path_kmer_len = 0, graph_matching = 1 (what grmpy runs: src/c++/main/grmpy.cpp:69-72).  The KmerAligner
 * stage is pg_set_kmer_stage below; the KlibAligner stage is not built. */
int pg_set_stages(pg_ctx* ctx, int32_t path_kmer_len, int32_t graph_matching, int32_t nonuniq_second_chance);
/* counters4 = {attempted, anchored, mapped} of the last batch (PathAligner::attempted/anchored/mapped,
 * PathAligner.hh:66-68) + the host time in microseconds of the last index build; path_ms = the stage's device time. */
int pg_path_stats(pg_ctx* ctx, uint64_t* counters4, float* path_ms);

/* ---- k-mer stage: grm::KmerAligner<K> (src/c++/lib/grm/KmerAligner.cpp; CompositeAligner.cpp:105-126, between the
 * exact-match stage and gssw; off by default in both CLIs, main/grmpy.cpp:72) ------------------------------------------
 * Gapless alignment of the read, on both strands, to the sequence of one of the site's PATHS: offsets proposed by shared
 * k-mers, at most two mismatches; score = matching bases, soft clips where the path is 'N' (N-filled source / sink),
 * unique unless an equally good candidate gives a different (position, CIGAR) -- then the read goes on to gssw with the
 * bases KmerAligner left behind, like the reference's BAD_ALIGN (pickBest, KmerAligner.cpp:479-517).  Bit-exact incl.
 * the order libstdc++'s heap keeps equal candidates in (oracle/pg_oracle_kmer.c). */
/* The paths of a site's graph JSON (grm::pathsFromJson, GraphInput.cpp:168-197): path p = the nodes
 * path_nodes[path_ptr[p] .. path_ptr[p+1]), whole nodes, consecutive ones joined by an edge.  At most 62 per site. */
int pg_set_paths(pg_ctx* ctx, int32_t site, int32_t n_paths, const int32_t* path_ptr, const int32_t* path_nodes);
/* kmer_len: 0 = off (default), else 2..16 (the reference instantiates 16; its unit test 10). */
int pg_set_kmer_stage(pg_ctx* ctx, int32_t kmer_len);
/* counters2 = {attempted, mapped} of the last batch (KmerAligner::attempted / mapped); kmer_ms = device time. */
int pg_kmer_stats(pg_ctx* ctx, uint64_t* counters2, float* kmer_ms);

/* Kernels launched by this context so far, and the last batch's per-kernel device time in ms
 * (fill, traceback) measured with CUDA events on the launching stream. */
int pg_stats(const pg_ctx* ctx, uint64_t* kernel_launches, float* last_fill_ms, float* last_trace_ms);
/* Library / build description, e.g. "paragraph_b200 sm_100a CK=16". */
const char* pg_version(void);

#ifdef __cplusplus
}
#endif
#endif
