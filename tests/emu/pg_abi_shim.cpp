// TEST INFRASTRUCTURE ONLY -- the lane emulator (pg_emu.cpp: the device source of pg_core.cuh / pg_path.cuh /
// pg_count.cuh compiled for the host) behind the signatures of include/pg_align.h, so that the C++ host mirror of the
// reference interface (paragraph_b200/csrc/host/pg_grm.hh) and its test program (tests/cpp/test_grm_mirror.cpp) can
// be exercised on a machine without a GPU.  The entry points are compiled under the names pgshim_* (pg_shim_names.h):
// this library cannot stand in for libpgalign.so, and nothing in the product loads it.  Covered: what the mirror calls
// (contexts, graphs, pg_align_batch incl. the stages of pg_set_stages, pg_format_cigar, edge labels, pg_batch_count);
// streams, pinned buffers and timings are meaningless here and return neutral values.
#include "pg_shim_names.h"

#include "../../include/pg_align.h"

#include "pg_emu.cpp"

#include <cstdlib>

struct pg_ctx
{
    std::string err;
    host::GraphStore graphs;
    int path_k = 0, kmer_k = 0;
    bool gssw_on = true, second = false;
    uint64_t kmer_counters[2] = { 0, 0 };
    // the last batch
    bool ran = false, have_sites = false;
    int n_reads = 0;
    std::vector<Record> recs;
    std::vector<uint32_t> arena;
    std::vector<int32_t> read_off, site;
    uint64_t path_counters[3] = { 0, 0, 0 };
    // pg_batch_upload / pg_batch_run: the staged input of the three-step form
    std::vector<char> up_bases;
    std::vector<int32_t> up_off, up_site;
    int up_n = -1;
    bool up_have_site = false, up_ran = false;
    uint32_t up_flags = 0;
};

namespace
{
int shim_fail(pg_ctx* c, int code, const std::string& msg)
{
    if (c)
        c->err = msg;
    return code;
}

// the kernels' geometry dispatch at W = 32 (pg_kernels.cu: pg_batch_run)
int shim_dp(const SiteDev& sd, const uint8_t* gb, const int32_t* gi, const uint8_t* b, int L, unsigned flags, Record& rec,
            std::vector<uint32_t>& ops)
{
    int nt = 0;
    if (L > BYTE_MAX_READ_LEN)
        return L <= 320 ? emu_align_one<10, 32>(sd, gb, gi, b, L, flags, rec, ops, &nt)
            : (L <= 512 ? emu_align_one<16, 32>(sd, gb, gi, b, L, flags, rec, ops, &nt)
                        : emu_align_one<32, 32>(sd, gb, gi, b, L, flags, rec, ops, &nt));
    return L <= 160 ? emu_align_one<5, 32>(sd, gb, gi, b, L, flags, rec, ops, &nt)
                    : emu_align_one<8, 32>(sd, gb, gi, b, L, flags, rec, ops, &nt);
}
} // namespace

extern "C" {

const char* pg_version(void) { return "paragraph_b200 ABI shim over the lane emulator (tests only)"; }

int pg_create(int, pg_ctx** out)
{
    if (!out)
        return PG_E_ARG;
    *out = new pg_ctx();
    return PG_OK;
}
void pg_destroy(pg_ctx* c) { delete c; }
const char* pg_last_error(const pg_ctx* c) { return c ? c->err.c_str() : "null context"; }
int pg_set_stream(pg_ctx* c, void*) { return c ? PG_OK : PG_E_ARG; }
int pg_set_scratch_limit(pg_ctx* c, uint64_t) { return c ? PG_OK : PG_E_ARG; }

int pg_add_graph(pg_ctx* c, int32_t n_nodes, const char* blob, const int32_t* off, int32_t n_edges, const int32_t* ef,
                 const int32_t* et, int32_t* site_id)
{
    if (!c)
        return PG_E_ARG;
    std::string err;
    const int s = c->graphs.add(n_nodes, blob, off, n_edges, ef, et, err);
    if (s < 0)
        return shim_fail(c, PG_E_GRAPH, err);
    if (site_id)
        *site_id = s;
    return PG_OK;
}
int pg_add_graphs(pg_ctx* c, int32_t n_sites, const int32_t* node_ptr, const char* blob, const int32_t* off,
                  const int32_t* edge_ptr, const int32_t* ef, const int32_t* et, int32_t* first_site_id)
{
    if (!c || n_sites < 0 || (n_sites > 0 && (!node_ptr || !blob || !off || !edge_ptr)))
        return PG_E_ARG;
    const pg::host::GraphStore::Mark mark = c->graphs.mark();
    const int first = (int)c->graphs.sites.size();
    std::string err;
    for (int32_t s = 0; s < n_sites; ++s)
        if (c->graphs.add(node_ptr[s + 1] - node_ptr[s], blob, off + node_ptr[s], edge_ptr[s + 1] - edge_ptr[s],
                          ef ? ef + edge_ptr[s] : nullptr, et ? et + edge_ptr[s] : nullptr, err) < 0)
        {
            c->graphs.rollback(mark);
            return shim_fail(c, PG_E_GRAPH, "site " + std::to_string(s) + " of the batch: " + err);
        }
    if (first_site_id)
        *first_site_id = first;
    return PG_OK;
}
int pg_clear_graphs(pg_ctx* c)
{
    if (!c)
        return PG_E_ARG;
    c->graphs.clear();
    c->ran = false;
    return PG_OK;
}

int pg_set_stages(pg_ctx* c, int32_t path_kmer_len, int32_t graph_matching, int32_t nonuniq_second_chance)
{
    if (!c || path_kmer_len < 0 || path_kmer_len > 4096)
        return shim_fail(c, PG_E_ARG, "pg_set_stages: bad k-mer length");
    if (path_kmer_len == 0 && !graph_matching && c->kmer_k == 0)
        return shim_fail(c, PG_E_ARG, "pg_set_stages: no alignment stage enabled");
    c->path_k = path_kmer_len;
    c->gssw_on = graph_matching != 0;
    c->second = nonuniq_second_chance != 0;
    return PG_OK;
}
int pg_set_paths(pg_ctx* c, int32_t site, int32_t n_paths, const int32_t* path_ptr, const int32_t* path_nodes)
{
    if (!c)
        return PG_E_ARG;
    std::string err;
    if (!c->graphs.set_paths(site, n_paths, path_ptr, path_nodes, err))
        return shim_fail(c, PG_E_GRAPH, err);
    return PG_OK;
}
int pg_set_kmer_stage(pg_ctx* c, int32_t kmer_len)
{
    if (!c || kmer_len < 0 || kmer_len == 1 || kmer_len > 16)
        return shim_fail(c, PG_E_ARG, "pg_set_kmer_stage: k-mer length must be 0 (off) or 2..16");
    c->kmer_k = kmer_len;
    return PG_OK;
}
int pg_kmer_stats(pg_ctx* c, uint64_t* counters2, float* kmer_ms)
{
    if (!c)
        return PG_E_ARG;
    if (counters2)
    {
        counters2[0] = c->kmer_counters[0];
        counters2[1] = c->kmer_counters[1];
    }
    if (kmer_ms)
        *kmer_ms = 0.f;
    return PG_OK;
}
int pg_path_stats(pg_ctx* c, uint64_t* counters4, float* path_ms)
{
    if (!c)
        return PG_E_ARG;
    if (counters4)
    {
        counters4[0] = c->path_counters[0];
        counters4[1] = c->path_counters[1];
        counters4[2] = c->path_counters[2];
        counters4[3] = 0;
    }
    if (path_ms)
        *path_ms = 0.f;
    return PG_OK;
}

// the cascade of pg_batch_run (pg_kernels.cu: run_path_stage + run_chunks), read by read
int pg_align_batch(pg_ctx* c, int32_t n_reads, const char* bases, const int32_t* off, const int32_t* site, uint32_t flags,
                   pg_record* records, uint32_t* ops_out, uint64_t cap, uint64_t* used)
{
    static_assert(sizeof(pg_record) == sizeof(Record), "pg_record layout");
    if (!c || n_reads < 0 || (n_reads > 0 && (!bases || !off || !records)))
        return shim_fail(c, PG_E_ARG, "pg_align_batch: bad arguments");
    const host::GraphStore& gs = c->graphs;
    if (gs.sites.empty())
        return shim_fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    for (int i = 0; i < n_reads; ++i)
    {
        const int L = off[i + 1] - off[i];
        if (L <= 0 || L > MAX_READ_LEN)
            return shim_fail(c, PG_E_READ_LEN, "read " + std::to_string(i) + " is empty or too long");
        if (site && (site[i] < 0 || (size_t)site[i] >= gs.sites.size()))
            return shim_fail(c, PG_E_ARG, "read " + std::to_string(i) + ": unknown site");
    }
    host::PathIndexHost ix;
    if (c->path_k > 0)
        host::build_path_index(gs, c->path_k, ix);
    host::KmerIndexHost kix;
    KmerView kv;
    memset(&kv, 0, sizeof kv);
    if (c->kmer_k > 0)
    {
        host::build_kmer_index(gs, c->kmer_k, kix);
        if (kix.max_paths > KMER_MAX_PATHS)
            return shim_fail(c, PG_E_GRAPH, "k-mer stage: too many paths on a site");
        kv.sites = kix.sites.data();
        kv.paths = kix.paths.data();
        kv.seqs = kix.seqs.data();
        kv.nodes = kix.nodes.data();
        kv.kmers = kix.kmers.data();
        kv.k = c->kmer_k;
    }
    c->kmer_counters[0] = c->kmer_counters[1] = 0;
    c->n_reads = n_reads;
    c->have_sites = site != nullptr;
    c->read_off.assign(off, off + n_reads + 1);
    c->site.assign((size_t)n_reads, 0);
    c->recs.assign((size_t)n_reads, Record());
    c->arena.clear();
    c->path_counters[0] = c->path_counters[1] = c->path_counters[2] = 0;
    const uint8_t* gb = gs.bytes.data();
    const int32_t* gi = gs.ints.data();
    for (int i = 0; i < n_reads; ++i)
    {
        const int s = site ? site[i] : 0;
        c->site[(size_t)i] = s;
        const SiteDev& sd = gs.sites[(size_t)s];
        const uint8_t* b = (const uint8_t*)bases + off[i];
        const int L = off[i + 1] - off[i];
        Record rec;
        memset(&rec, 0, sizeof rec);
        std::vector<uint32_t> ops;
        std::vector<uint8_t> q0((size_t)L + 1), q1((size_t)L + 1);
        bool to_dp = true;
        int flips = 0; // reverse complements the stages in front applied to the bases (the later stages see them)
        std::vector<uint8_t> cur(b, b + L), tmp;
        if (c->path_k > 0)
        {
            const PathView v = make_path_view(ix.sites[(size_t)s], sd, ix.table.data(), ix.lists.data(), ix.succ.data(), gb, gi);
            path_strand_chars(b, L, 0, q0.data());
            path_strand_chars(b, L, 1, q1.data());
            PathResult rf, rr, r;
            path_scan_strand(v, q0.data(), L, 0, rf);
            path_scan_strand(v, q1.data(), L, 1, rr);
            path_combine(rf, rr, r);
            ++c->path_counters[0];
            c->path_counters[1] += r.n_matches > 0;
            c->path_counters[2] += r.n_full > 0;
            if (r.n_full > 1 && c->second && (c->gssw_on || c->kmer_k > 0))
            {
                if (r.strand != 0) // the later stages get the bases PathAligner left behind
                {
                    cur.assign(q1.begin(), q1.begin() + L);
                    ++flips;
                }
            }
            else if (r.n_full > 0)
            {
                path_record(r, L, rec);
                ops.resize((size_t)r.first.n_nodes);
                path_emit(v, r.strand ? q1.data() : q0.data(), L, r, ops.data());
                to_dp = false;
            }
        }
        if (to_dp && c->kmer_k > 0) // pg_kmer_kernel
        {
            const KmerResult kr = emu_kmer_one(kv, s, cur.data(), L, kix.max_path_nodes, ops, tmp);
            ++c->kmer_counters[0];
            if (kr.status == 1 || (kr.status == 2 && !c->gssw_on))
            {
                rec.graph_pos = kr.pos;
                rec.score = (int16_t)kr.score;
                rec.query_clipped = (uint16_t)kr.clipped;
                rec.unique = (uint8_t)(kr.status == 1);
                rec.chose_reverse = (uint8_t)kr.rev;
                rec.mapped_by = (uint8_t)(flips ? STAGE_KMER_REV : STAGE_KMER);
                rec.cigar_len = (uint32_t)ops.size();
                c->kmer_counters[1] += kr.status == 1;
                to_dp = false;
            }
            else
            {
                ops.clear();
                if (kr.status == 2 && kr.rev)
                {
                    cur = tmp;
                    ++flips;
                }
            }
        }
        if (to_dp && !c->gssw_on)
        {
            rec.status = (uint8_t)ST_UNMAPPED;
            to_dp = false;
        }
        if (to_dp)
        {
            const int rc = shim_dp(sd, gb, gi, cur.data(), L, flags, rec, ops);
            if (rc)
                return shim_fail(c, PG_E_ARG, "emulator failure " + std::to_string(rc));
            rec.mapped_by = (uint8_t)(flips == 0 ? STAGE_GSSW : (flips == 1 ? STAGE_GSSW_REV : STAGE_GSSW_REV2));
        }
        rec.cigar_off = (uint32_t)c->arena.size();
        c->arena.insert(c->arena.end(), ops.begin(), ops.begin() + rec.cigar_len);
        c->recs[(size_t)i] = rec;
    }
    c->ran = true;
    if (used)
        *used = c->arena.size();
    if (c->arena.size() > cap || (!ops_out && !c->arena.empty()))
        return shim_fail(c, PG_E_CAPACITY, "cigar arena too small: need " + std::to_string(c->arena.size()) + " words");
    memcpy(records, c->recs.data(), (size_t)n_reads * sizeof(Record));
    if (!c->arena.empty())
        memcpy(ops_out, c->arena.data(), c->arena.size() * sizeof(uint32_t));
    return PG_OK;
}

// the three-step form: upload keeps a copy of the input, run notes the flags, download does the work
int pg_batch_upload(pg_ctx* c, int32_t n_reads, const char* bases, const int32_t* off, const int32_t* site)
{
    if (!c || n_reads < 0 || (n_reads > 0 && (!bases || !off)))
        return shim_fail(c, PG_E_ARG, "pg_batch_upload: bad arguments");
    c->up_n = n_reads;
    c->up_off.assign(off, off + n_reads + 1);
    c->up_bases.assign(bases, bases + (n_reads ? off[n_reads] : 0));
    c->up_have_site = site != nullptr;
    if (site)
        c->up_site.assign(site, site + n_reads);
    c->up_ran = false;
    return PG_OK;
}

int pg_batch_run(pg_ctx* c, uint32_t flags)
{
    if (!c || c->up_n < 0)
        return shim_fail(c, PG_E_STATE, "pg_batch_run: no batch uploaded");
    c->up_flags = flags;
    c->up_ran = true;
    return PG_OK;
}

int pg_batch_download(pg_ctx* c, pg_record* records, uint32_t* ops, uint64_t cap, uint64_t* used)
{
    if (!c || !c->up_ran)
        return shim_fail(c, PG_E_STATE, "pg_batch_download: no batch ran");
    return pg_align_batch(c, c->up_n, c->up_bases.data(), c->up_off.data(), c->up_have_site ? c->up_site.data() : nullptr,
                          c->up_flags, records, ops, cap, used);
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
    if (!out)
        return PG_E_ARG;
    *out = malloc(bytes ? (size_t)bytes : 1);
    return *out ? PG_OK : PG_E_CUDA;
}
void pg_host_free(void* p) { free(p); }

int pg_set_edge_labels(pg_ctx* c, int32_t site, const uint64_t* masks)
{
    if (!c)
        return PG_E_ARG;
    if (site < 0 || (size_t)site >= c->graphs.sites.size())
        return shim_fail(c, PG_E_ARG, "pg_set_edge_labels: unknown site " + std::to_string(site));
    const int64_t eb = c->graphs.edge_base[(size_t)site];
    const int n = c->graphs.sites[(size_t)site].n_edges;
    for (int e = 0; e < n; ++e)
        c->graphs.in_label[(size_t)(eb + e)] = masks ? masks[e] : 0ull;
    return PG_OK;
}

// pg_batch_count (pg_kernels.cu) with the three kernels as loops over their threads
int pg_batch_count(pg_ctx* c, const int32_t* fragment, const uint8_t* is_reverse_strand, const pg_count_params* params,
                   pg_read_support* support, uint32_t* path_words, uint64_t path_cap, uint64_t* path_used,
                   pg_count4* node_counts, uint64_t node_cap, pg_count4* edge_counts, uint64_t edge_cap,
                   uint32_t* family_words, uint64_t family_cap, uint64_t* family_used)
{
    static_assert(sizeof(pg_read_support) == sizeof(ReadSupport) && sizeof(pg_count4) == sizeof(Count4), "ABI structs");
    if (!c || !params)
        return shim_fail(c, PG_E_ARG, "pg_batch_count: bad arguments");
    if (!c->ran)
        return shim_fail(c, PG_E_STATE, "pg_batch_count before pg_batch_run");
    const host::GraphStore& gs = c->graphs;
    const size_t ns = gs.sites.size();
    if (ns == 0)
        return shim_fail(c, PG_E_STATE, "no graph registered (pg_add_graph)");
    const uint64_t total_nodes = (uint64_t)gs.node_base[ns], total_edges = (uint64_t)gs.edge_base[ns];
    if ((node_counts && node_cap < total_nodes) || (edge_counts && edge_cap < total_edges))
        return shim_fail(c, PG_E_CAPACITY, "pg_batch_count: need " + std::to_string(total_nodes) + " node rows and "
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
    const int n = c->n_reads;
    if (n == 0)
        return PG_OK;
    std::vector<int32_t> next;
    std::vector<uint8_t> head;
    {
        std::string err;
        if (!host::build_fragment_chains(fragment, c->have_sites ? c->site.data() : nullptr, n, next, head, err))
            return shim_fail(c, PG_E_ARG, "pg_batch_count: " + err);
    }
    const uint64_t n_ops = c->arena.size();
    if (path_words && path_cap < n_ops)
        return shim_fail(c, PG_E_CAPACITY, "pg_batch_count: path_words needs " + std::to_string(n_ops) + " words");
    host::CountHostTables ht;
    host::build_count_tables(gs, prm.family_slots, ht);
    CountTables t;
    t.sites = gs.sites.data();
    t.gints = gs.ints.data();
    t.csite = ht.csite.data();
    t.csr_input = ht.csr_input.data();
    t.lab_edge = ht.lab_edge.data();
    t.lab_out = ht.lab_out.data();
    t.lab_in = ht.lab_in.data();
    std::vector<ReadSupport> sup((size_t)n);
    std::vector<uint32_t> path((size_t)n_ops + 1, 0u);
    for (int i = 0; i < n; ++i) // pg_support_kernel
        support_read(c->recs[(size_t)i], c->arena.data(), c->read_off[(size_t)i + 1] - c->read_off[(size_t)i],
                     c->site[(size_t)i], is_reverse_strand ? is_reverse_strand[i] != 0 : false, t, prm, sup[(size_t)i],
                     path.data());
    std::vector<Count4> nc((size_t)total_nodes + 1), ec((size_t)total_edges + 1), fam((size_t)ht.fam_rows + 1);
    std::vector<unsigned long long> keys((size_t)ht.fam_keys + 1, 0ull);
    memset(nc.data(), 0, nc.size() * sizeof(Count4));
    memset(ec.data(), 0, ec.size() * sizeof(Count4));
    memset(fam.data(), 0, fam.size() * sizeof(Count4));
    long long overflow_site = -1;
    for (int i = 0; i < n; ++i) // pg_fragment_kernel
        if (head[(size_t)i]
            && !count_fragment(i, c->site[(size_t)i], next.data(), sup.data(), path.data(), t, prm, nc.data(), ec.data(),
                               keys.data(), fam.data()))
            overflow_site = std::max<long long>(overflow_site, c->site[(size_t)i]);
    std::vector<uint32_t> fam_out;
    for (size_t s = 0; s < ns; ++s) // pg_family_compact_kernel
    {
        const CountSite cs = ht.csite[s];
        const SiteDev& sd = gs.sites[s];
        const int rows = 1 + sd.n_nodes + sd.n_edges;
        for (int slot = 0; slot < cs.slots; ++slot)
        {
            const unsigned long long key = keys[(size_t)(cs.key_base + slot)];
            if (!key)
                continue;
            fam_out.push_back((uint32_t)s);
            fam_out.push_back((uint32_t)rows);
            fam_out.push_back((uint32_t)key);
            fam_out.push_back((uint32_t)(key >> 32));
            const uint32_t* src = reinterpret_cast<const uint32_t*>(fam.data() + cs.fam_base + (int64_t)slot * rows);
            fam_out.insert(fam_out.end(), src, src + 4 * rows);
        }
    }
    if (support)
        memcpy(support, sup.data(), (size_t)n * sizeof(ReadSupport));
    if (path_words)
        memcpy(path_words, path.data(), (size_t)n_ops * sizeof(uint32_t));
    if (node_counts)
        memcpy(node_counts, nc.data(), (size_t)total_nodes * sizeof(Count4));
    if (edge_counts)
        memcpy(edge_counts, ec.data(), (size_t)total_edges * sizeof(Count4));
    if (path_used)
        *path_used = n_ops;
    if (overflow_site >= 0)
        return shim_fail(c, PG_E_CAPACITY, "pg_batch_count: site " + std::to_string(overflow_site) + " has more than "
                                               + std::to_string(prm.family_slots) + " distinct path-family sets (raise family_slots)");
    if (family_used)
        *family_used = fam_out.size();
    if (!fam_out.empty())
    {
        if (!family_words || family_cap < fam_out.size())
            return shim_fail(c, PG_E_CAPACITY, "pg_batch_count: family_words needs " + std::to_string(fam_out.size()) + " words");
        memcpy(family_words, fam_out.data(), fam_out.size() * sizeof(uint32_t));
    }
    return PG_OK;
}

int pg_count_stats(const pg_ctx* c, uint64_t* launches, float* ms)
{
    if (!c)
        return PG_E_ARG;
    if (launches)
        *launches = 0;
    if (ms)
        *ms = 0.f;
    return PG_OK;
}
int pg_stats(const pg_ctx* c, uint64_t* launches, float* fill_ms, float* trace_ms)
{
    if (!c)
        return PG_E_ARG;
    if (launches)
        *launches = 0;
    if (fill_ms)
        *fill_ms = 0.f;
    if (trace_ms)
        *trace_ms = 0.f;
    return PG_OK;
}
}
