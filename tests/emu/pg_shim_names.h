/* TEST INFRASTRUCTURE ONLY.  Renames the C-ABI entry points of include/pg_align.h to pgshim_* so that the host mirror
 * (paragraph_b200/csrc/host/pg_grm.hh) can be compiled against tests/emu/pg_abi_shim.cpp -- the lane emulator behind
 * the same signatures -- on a machine without a GPU.  Because of the renaming the shim library exports no pg_* symbol:
 * it cannot be loaded in place of libpgalign.so (PG_LIB), and nothing in the product includes this file. */
#ifndef PG_SHIM_NAMES_H
#define PG_SHIM_NAMES_H
#define pg_create pgshim_create
#define pg_destroy pgshim_destroy
#define pg_last_error pgshim_last_error
#define pg_set_stream pgshim_set_stream
#define pg_set_scratch_limit pgshim_set_scratch_limit
#define pg_add_graph pgshim_add_graph
#define pg_add_graphs pgshim_add_graphs
#define pg_clear_graphs pgshim_clear_graphs
#define pg_align_batch pgshim_align_batch
#define pg_batch_upload pgshim_batch_upload
#define pg_batch_run pgshim_batch_run
#define pg_batch_download pgshim_batch_download
#define pg_host_alloc pgshim_host_alloc
#define pg_host_free pgshim_host_free
#define pg_format_cigar pgshim_format_cigar
#define pg_set_edge_labels pgshim_set_edge_labels
#define pg_batch_import pgshim_batch_import
#define pg_batch_count pgshim_batch_count
#define pg_count_stats pgshim_count_stats
#define pg_set_stages pgshim_set_stages
#define pg_path_stats pgshim_path_stats
#define pg_set_paths pgshim_set_paths
#define pg_set_kmer_stage pgshim_set_kmer_stage
#define pg_kmer_stats pgshim_kmer_stats
#define pg_stats pgshim_stats
#define pg_version pgshim_version
#endif
