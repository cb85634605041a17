/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of paragraph's read->graph
 * alignment path.  Not part of the product: only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load this library.
 *
 * Parity status: PINNED.  pg_oracle.c is checked (tests/test_oracle.py) against
 *   - the reference's own golden vectors (src/c++/test/test_paragraph_parts.cpp:113-144),
 *   - oracle/_ref/libpgref.so = the unmodified reference sources compiled here
 *     (cell-by-cell mH/mE/mF, best cell, CIGAR, uniqueness, strand) on seeded fuzz inputs,
 *   - the committed fixtures under tests/golden/ generated from that library.
 */
#ifndef PG_ORACLE_H
#define PG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGO_AF_CIGAR 1u
#define PGO_AF_BOTH_STRANDS 2u
#define PGO_AF_REVERSE_GRAPH 4u
#define PGO_AF_ALL 0xFFFFFFFFu

#define PGO_OK 0
#define PGO_E_BYTE_OVERFLOW (-2) /* no longer returned: the 16-bit fallback (gssw.c:4001-4013) is restated */
#define PGO_E_ARG (-1)

typedef struct pgo_graph pgo_graph;

/* Nodes must be in topological order (edge from < to), as graphtools::Graph enforces
 * (graph-tools src/graphcore/Graph.cpp:113-116).  Sequences are upper-cased on load
 * (src/c++/lib/grm/GraphAligner.cpp:126,138); predecessor lists are ascending ids (:147-157). */
pgo_graph* pgo_graph_create(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges,
                            const int32_t* efrom, const int32_t* eto);
void pgo_graph_destroy(pgo_graph* g);

/* GraphAligner::alignRead (GraphAligner.cpp:308-404).
 * out6 = {graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen};
 * out_bases (len bytes, may be NULL) receives the possibly reverse-complemented bases. */
int pgo_align_read(const pgo_graph* g, const char* bases, int len, int is_reverse_strand, unsigned flags,
                   int32_t* out6, char* out_bases, char* cigar, int cigar_cap);

int pgo_align_batch(const pgo_graph* g, int n_reads, const char* bases_blob, const int32_t* read_off,
                    const uint8_t* is_rev, unsigned flags, int32_t* out6, char* out_bases_blob, char* cigars,
                    int cigar_stride);

/* One gssw_graph_fill + gssw_graph_trace_back on the forward (reversed_graph=0) or reversed
 * graph.  `read` must already be upper-case.  node_stats[4*n] = {score1, ref_end1, read_end1, is_byte};
 * mats (may be NULL) = per node mH,mE,mF (len*L bytes each; filled in byte mode only); res3 = {max_node, position, score};
 * multi (may be NULL) receives alignsEndAtMultNodes (GraphAligner.cpp:170-212). */
int pgo_fill_trace(const pgo_graph* g, int reversed_graph, const char* read, int L, int32_t* node_stats,
                   uint8_t* mats, int32_t* res3, int32_t* multi, char* cigar, int cigar_cap);

/* Same with 16-bit matrices: valid in both modes.  node_stats[4*n+3] = is_byte (0 after the reference's fallback to
 * 16-bit mode, which happens when a score reaches 251: gssw.c:380, 467, 4001-4013). */
int pgo_fill_trace16(const pgo_graph* g, int reversed_graph, const char* read, int L, int32_t* node_stats,
                     uint16_t* mats, int32_t* res3, int32_t* multi, char* cigar, int cigar_cap);

/* readfilters::BadAlign on a graph CIGAR string (BadAlign.hh:62-73); NonUniq is simply !unique (NonUniq.hh:48-52) */
int pgo_bad_align(const char* cigar, double bad_align_frac, int* clipped_out);

/* 0 = faithful restatement (default); 1/2 = plain-recurrence model variants used by tests/ to
 * prove the CUDA path's simplifications decision-equivalent (see pg_oracle.c). Not thread-safe. */
void pgo_set_fill_variant(int v);

/* ---- read filtering + disambiguation + read counting of one site (oracle/pg_oracle_counts.c) ---- */
enum { PGO_V_MAPPED = 0, PGO_V_NONUNIQ = 1, PGO_V_BAD_ALIGN = 2, PGO_V_INVALID = 3 };
#define PGO_SUP_NODE_MASK 0xFFFFu
#define PGO_SUP_NODE 0x40000000u /* path word: the read supports this node */
#define PGO_SUP_EDGE 0x80000000u /* path word: the read supports the edge (previous path node -> this node) */
typedef struct pgo_read_support
{
    uint64_t sequences; /* bit k: path family (edge label) k is in graph_sequences_supported */
    uint32_t path_off;  /* first path word of this read */
    uint16_t path_len;  /* path nodes (0 unless MAPPED) */
    uint8_t verdict;    /* PGO_V_* */
    uint8_t graph_reverse;
} pgo_read_support;
typedef struct pgo_count4
{
    uint32_t fragments, reads, fwd, rev; /* "<name>", ":READS", ":FWD", ":REV" of ReadCounting.cpp:52-69 */
} pgo_count4;

/* Filter chain (NonUniq if remove_nonuniq, then BadAlign), disambiguateReads with the node/edge support filters of
 * alignAndDisambiguate (use_support_filters=1) or with null filters (0, as the reference's unit tests call it),
 * fragments by `fragment` id (>= 0; NULL = every read its own fragment) and the three count tables.
 * node_len[n_nodes]; edge_labels[n_edges] (bit k = label k) or NULL; cigars = n_reads strings, cigar_stride apart.
 * node_counts[n_nodes], edge_counts[n_edges]; family_words receives one entry per distinct non-empty label set,
 * in order of first appearance: {mask_lo, mask_hi, (1 + n_nodes + n_edges) x pgo_count4 = total, nodes, edges}.
 * Returns 0, -1 path_words too small, -2 negative fragment id, -3 family_words too small. */
int pgo_count_site(
    int n_nodes, const int32_t* node_len, int n_edges, const int32_t* efrom, const int32_t* eto,
    const uint64_t* edge_labels, int n_reads, const int32_t* read_len, const int32_t* graph_pos, const uint8_t* unique,
    const char* cigars, int cigar_stride, const uint8_t* is_graph_reverse, const int32_t* fragment, int remove_nonuniq,
    double bad_align_frac, int use_support_filters, pgo_read_support* support, uint32_t* path_words, int path_cap,
    int* path_used, pgo_count4* node_counts, pgo_count4* edge_counts, uint32_t* family_words, int family_cap,
    int* family_used);

/* ---- grm::PathAligner, the exact-match stage in front of the DP (oracle/pg_oracle_path.c) ---- */
struct pgo_path_index;
struct pgo_path_index* pgo_path_index_create(int n_nodes, const char* seq_blob, const int32_t* seq_off, int n_edges,
                                             const int32_t* efrom, const int32_t* eto, int kmer_len);
void pgo_path_index_destroy(struct pgo_path_index* ix);
/* out8 = {mapped, graph_pos, score, unique, mapq, is_graph_reverse_strand, cigar_strlen, anchored} */
int pgo_path_align_read(const struct pgo_path_index* ix, const char* bases, int len, int32_t* out8, char* out_bases,
                        char* cigar, int cigar_cap);
/* counters3 = {attempted, anchored, mapped} (PathAligner.hh:66-68) */
int pgo_path_align_batch(const struct pgo_path_index* ix, int n_reads, const char* bases_blob, const int32_t* read_off,
                         int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride, int32_t* counters3);

/* ---- grm::KmerAligner<K> (oracle/pg_oracle_kmer.c): gapless alignment to the graph's paths, seeded by k-mers ---- */
struct pgo_kmer_index;
struct pgo_kmer_index* pgo_kmer_index_create(int n_nodes, const char* blob, const int32_t* off, int n_paths,
                                             const int32_t* path_ptr, const int32_t* path_nodes, int k);
void pgo_kmer_index_destroy(struct pgo_kmer_index* ix);
/* out8 = {status (0 UNMAPPED, 1 MAPPED, 2 BAD_ALIGN = not unique), graph_pos, score, unique, mapq,
 *         is_graph_reverse_strand, cigar_strlen, 0} */
int pgo_kmer_align_read(const struct pgo_kmer_index* ix, const char* bases, int L, int is_reverse_strand, int32_t* out8,
                        char* out_bases, char* cigar, int cigar_cap);
int pgo_kmer_align_batch(const struct pgo_kmer_index* ix, int n_reads, const char* blob, const int32_t* off,
                         const uint8_t* is_rev, int32_t* out8, char* out_bases_blob, char* cigars, int cigar_stride);

#ifdef __cplusplus
}
#endif
#endif
