// paragraph_b200 -- host-side C++ mirror of the reference's aligner interface for the read->graph path,
// implemented on top of the C-ABI (include/pg_align.h).  Header only, no Boost / htslib / spdlog.
//
// Mirrors, with the same names, argument meaning and error behaviour (exceptions):
//   grm::GraphAligner      setGraph / alignRead / align / AF_* flags   (src/c++/include/grm/GraphAligner.hh:36-88)
//   grm::CompositeAligner  ctor flags, setGraph, alignRead(read, filter), counters
//                                                                     (src/c++/include/grm/CompositeAligner.hh:44-91)
//   grm::alignReads        batch entry: keeps MAPPED reads only        (src/c++/include/grm/Align.hh:49-52,
//                                                                      src/c++/lib/grm/Align.cpp:114-156)
// The classes are templates over the caller's Graph / Read types so that a paragraph build can instantiate
// them with graphtools::Graph and common::Read unchanged; pgb::Graph / pgb::Read below are
// minimal stand-ins with the same member names for stand-alone use and for this repo's tests.
//
// The GPU engine is batch-first: alignReads() makes ONE pg_align_batch call for the whole read vector
// (results are written back in input order, so the outcome is deterministic for any `threads`); the per-read
// alignRead() facades are batches of one.  The path stage (grm::PathAligner) and the gssw stage run on the GPU; asking
// for the kmer / klib stages throws (SURVEY.md 8f "next" rows, not silently skipped).  `threads` host threads gather
// the batch into page-locked staging and write the records back into the reads.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <climits>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <exception>
#include <functional>
#include <future>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <set>
#include <stdexcept>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../../include/pg_align.h"

namespace pgb
{

// ---- stand-in data types (same member names as graphtools::Graph / common::Read) ------------------------
class Graph
{
public:
    explicit Graph(size_t n = 0) : seq_(n), name_(n), pred_(n), succ_(n) {}
    size_t numNodes() const { return seq_.size(); }
    void setNodeSeq(uint32_t id, std::string const& s) { seq_.at(id) = s; }
    void setNodeName(uint32_t id, std::string const& s) { name_.at(id) = s; }
    std::string const& nodeSeq(uint32_t id) const { return seq_.at(id); }
    std::string const& nodeName(uint32_t id) const { return name_.at(id); }
    void addEdge(uint32_t from, uint32_t to)
    {
        if (from > to) // graphtools::Graph::addEdge, Graph.cpp:113-116
            throw std::logic_error("Edge (" + std::to_string(from) + " ," + std::to_string(to) + ") breaks topological order");
        succ_.at(from).insert(to);
        pred_.at(to).insert(from);
    }
    std::set<uint32_t> const& predecessors(uint32_t id) const { return pred_.at(id); }
    std::set<uint32_t> const& successors(uint32_t id) const { return succ_.at(id); }
    bool hasEdge(uint32_t from, uint32_t to) const { return succ_.at(from).count(to) != 0; }
    void addLabelToEdge(uint32_t from, uint32_t to, std::string const& label) // graphtools Graph.cpp:151-158
    {
        if (!hasEdge(from, to))
            throw std::logic_error("There is no edge between " + std::to_string(from) + " and " + std::to_string(to));
        labels_[std::make_pair(from, to)].insert(label);
    }
    std::set<std::string> edgeLabels(uint32_t from, uint32_t to) const
    {
        auto it = labels_.find(std::make_pair(from, to));
        return it == labels_.end() ? std::set<std::string>() : it->second;
    }

private:
    std::vector<std::string> seq_, name_;
    std::vector<std::set<uint32_t>> pred_, succ_;
    std::map<std::pair<uint32_t, uint32_t>, std::set<std::string>> labels_;
};

struct Path // the part of graphtools::Path the k-mer stage needs: whole nodes from the first base of the first one to the
{           // last base of the last one, as grm::pathsFromJson builds them (GraphInput.cpp:168-197)
    Path() = default;
    explicit Path(std::vector<uint32_t> nodes) : nodes_(std::move(nodes)) {}
    std::vector<uint32_t> const& nodeIds() const { return nodes_; }

private:
    std::vector<uint32_t> nodes_;
};

class Read // the fields of common::Read the alignment path touches (src/c++/include/common/Read.hh:97-131)
{
public:
    enum MappingStatus { UNMAPPED = 0, MAPPED = 1, BAD_ALIGN = 2 };
    Read() = default;
    Read(std::string const& id, std::string const& bases, std::string const& quals) { setCoreInfo(id, bases, quals); }
    void setCoreInfo(std::string const& id, std::string const& bases, std::string const& quals)
    {
        fragment_id_ = id;
        bases_ = bases;
        quals_ = quals;
    }
    std::string const& fragment_id() const { return fragment_id_; }
    std::string const& bases() const { return bases_; }
    void set_bases(std::string const& v) { bases_ = v; }
    std::string const& quals() const { return quals_; }
    void set_quals(std::string const& v) { quals_ = v; }
    bool is_reverse_strand() const { return is_reverse_strand_; }
    void set_is_reverse_strand(bool v) { is_reverse_strand_ = v; }
    int32_t graph_pos() const { return graph_pos_; }
    void set_graph_pos(int32_t v) { graph_pos_ = v; }
    std::string const& graph_cigar() const { return graph_cigar_; }
    void set_graph_cigar(std::string v) { graph_cigar_ = std::move(v); }
    int32_t graph_mapq() const { return graph_mapq_; }
    void set_graph_mapq(int32_t v) { graph_mapq_ = v; }
    int32_t graph_alignment_score() const { return graph_alignment_score_; }
    void set_graph_alignment_score(int32_t v) { graph_alignment_score_ = v; }
    bool is_graph_alignment_unique() const { return is_graph_alignment_unique_; }
    void set_is_graph_alignment_unique(bool v) { is_graph_alignment_unique_ = v; }
    bool is_graph_reverse_strand() const { return is_graph_reverse_strand_; }
    void set_is_graph_reverse_strand(bool v) { is_graph_reverse_strand_ = v; }
    MappingStatus graph_mapping_status() const { return graph_mapping_status_; }
    void set_graph_mapping_status(MappingStatus s) { graph_mapping_status_ = s; }
    // what disambiguateReads fills in (Read.hh:115-131)
    std::vector<std::string> const& graph_nodes_supported() const { return nodes_supported_; }
    void add_graph_nodes_supported(std::string const& v) { nodes_supported_.push_back(v); }
    void clear_graph_nodes_supported() { nodes_supported_.clear(); }
    std::vector<std::string> const& graph_edges_supported() const { return edges_supported_; }
    void add_graph_edges_supported(std::string const& v) { edges_supported_.push_back(v); }
    void clear_graph_edges_supported() { edges_supported_.clear(); }
    std::vector<std::string> const& graph_sequences_supported() const { return sequences_supported_; }
    void add_graph_sequences_supported(std::string const& v) { sequences_supported_.push_back(v); }
    void clear_graph_sequences_supported() { sequences_supported_.clear(); }

private:
    std::vector<std::string> nodes_supported_, edges_supported_, sequences_supported_;
    std::string fragment_id_, bases_, quals_, graph_cigar_;
    bool is_reverse_strand_ = false;
    int32_t graph_pos_ = 0, graph_mapq_ = 0, graph_alignment_score_ = 0;
    bool is_graph_alignment_unique_ = false, is_graph_reverse_strand_ = false;
    MappingStatus graph_mapping_status_ = UNMAPPED;
};

namespace paragraph
{
// The reference's default read-filter chain (createReadFilter, src/c++/lib/paragraph/ReadFilter.cpp:73-90: NonUniq,
// then BadAlign) evaluated from the engine's record instead of decoding the CIGAR string:
//   NonUniq  : !is_graph_alignment_unique                                   (readfilters/NonUniq.hh:48-52)
//   BadAlign : read_len - query_clipped < round(bad_align_frac * read_len)  (readfilters/BadAlign.hh:62-73)
// Returns the reason like ReadFilter::filterRead does ("" = keep).  (KmerFilter is not on the GPU path.)
struct DefaultReadFilter
{
    bool remove_nonuniq = true;
    double bad_align_frac = 0.8;
    const char* operator()(pg_record const& r, int read_len) const
    {
        if (remove_nonuniq && !r.unique)
            return "nonuniq";
        const double thr = (double)(long)(bad_align_frac * read_len + 0.5); // round(), non-negative argument
        if ((double)(read_len - (int)r.query_clipped) < thr)
            return "bad_align";
        return "";
    }
};

// Result of the counting stage for one site: the three tables paragraph::countReads writes
// (src/c++/lib/paragraph/ReadCounting.cpp:215-233) with the same keys -- node name, "<from>_<to>", and for
// read_counts_by_sequence the sorted label names joined by "," holding "total" plus node and edge rows.
struct Count4
{
    uint64_t fragments = 0, reads = 0, fwd = 0, rev = 0; // JSON "<key>", "<key>:READS", "<key>:FWD", "<key>:REV"
    bool operator==(Count4 const& o) const { return fragments == o.fragments && reads == o.reads && fwd == o.fwd && rev == o.rev; }
};
struct SiteCounts
{
    std::map<std::string, Count4> read_counts_by_node, read_counts_by_edge;
    std::map<std::string, std::map<std::string, Count4>> read_counts_by_sequence;
    size_t invalid_alignments = 0; // reads whose CIGAR the reference's decodeGraphAlignment would reject (it throws there)
};
struct CountOptions // paragraph::Parameters defaults (include/paragraph/Parameters.hh:47-141)
{
    bool remove_nonuniq_reads = true;
    double bad_align_frac = 0.8;
    bool use_support_filters = true; // false = disambiguateReads(graph, reads) with null filters (the unit tests)
    int family_slots = 0;
};
} // namespace paragraph

namespace grm
{

namespace detail
{
// Worker threads shared by every aligner of the process (created on first use, grown on demand, joined at exit):
// grm::alignReads is called once per (sample, target) with a few hundred to a few thousand reads, and starting
// `threads` std::threads twice per call (pack, write-back) costs more than the work they do.
class WorkerPool
{
public:
    static WorkerPool& instance()
    {
        static WorkerPool p;
        return p;
    }
    // run job(0) .. job(n_jobs - 1); job 0 on the calling thread, the others on pool threads.  Rethrows the first exception.
    template <typename F> void run(size_t n_jobs, F&& job)
    {
        if (n_jobs <= 1)
        {
            if (n_jobs == 1)
                job((size_t)0);
            return;
        }
        Batch batch;
        batch.left = n_jobs - 1;
        batch.fn = [&job](size_t k) { job(k); };
        {
            std::lock_guard<std::mutex> lock(m_);
            while (threads_.size() < std::min<size_t>(n_jobs - 1, 256))
                threads_.emplace_back([this] { loop(); });
            for (size_t k = 1; k < n_jobs; ++k)
                queue_.emplace_back(&batch, k);
        }
        cv_.notify_all();
        std::exception_ptr mine;
        try
        {
            job((size_t)0);
        }
        catch (...)
        {
            mine = std::current_exception();
        }
        {
            // wait for the others -- and work the queue meanwhile, so that jobs that themselves call run() (a pipelined
            // batch: two halves, each with its own parallel pack / write-back) can never starve each other
            std::unique_lock<std::mutex> lock(m_);
            while (batch.left != 0)
            {
                if (queue_.empty())
                {
                    batch.done.wait(lock);
                    continue;
                }
                execute_front(lock);
            }
        }
        if (mine)
            std::rethrow_exception(mine);
        if (batch.err)
            std::rethrow_exception(batch.err);
    }
    ~WorkerPool()
    {
        {
            std::lock_guard<std::mutex> lock(m_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto& t : threads_)
            t.join();
    }

private:
    struct Batch
    {
        std::function<void(size_t)> fn;
        size_t left = 0;
        std::exception_ptr err;
        std::condition_variable done;
    };
    void execute_front(std::unique_lock<std::mutex>& lock) // takes one queued job; called and returns with the lock held
    {
        const std::pair<Batch*, size_t> item = queue_.front();
        queue_.pop_front();
        lock.unlock();
        std::exception_ptr err;
        try
        {
            item.first->fn(item.second);
        }
        catch (...)
        {
            err = std::current_exception();
        }
        lock.lock();
        if (err && !item.first->err)
            item.first->err = err;
        if (--item.first->left == 0)
            item.first->done.notify_all();
    }
    void loop()
    {
        std::unique_lock<std::mutex> lock(m_);
        for (;;)
        {
            cv_.wait(lock, [this] { return stop_ || !queue_.empty(); });
            if (queue_.empty())
                return; // stop_
            execute_front(lock);
        }
    }
    std::mutex m_;
    std::condition_variable cv_;
    std::deque<std::pair<Batch*, size_t>> queue_;
    std::vector<std::thread> threads_;
    bool stop_ = false;
};

// Where the host time of a batch goes (seconds, summed over calls; read and reset by a caller such as
// tools/cpp/bench_mirror): a handful of clock reads per batch.
struct Phases
{
    enum { SET_GRAPH, PACK, LAUNCH, WAIT, APPLY, FILTER, KEEP, N };
    std::atomic<uint64_t> ns[N];
    static Phases& instance()
    {
        static Phases p;
        return p;
    }
    static const char* name(int i)
    {
        static const char* const names[N] = { "set_graph", "pack", "launch", "wait_download", "apply", "filter", "keep" };
        return names[i];
    }
    void reset()
    {
        for (auto& x : ns)
            x = 0;
    }
    struct Scope // adds the lifetime of the object to one phase
    {
        explicit Scope(int phase) : phase_(phase), t0_(std::chrono::steady_clock::now()) {}
        ~Scope() { next(-1); }
        void next(int phase) // closes the running phase and starts `phase` (-1: none)
        {
            const auto t1 = std::chrono::steady_clock::now();
            if (phase_ >= 0)
                instance().ns[phase_] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0_).count();
            phase_ = phase;
            t0_ = t1;
        }
        int phase_;
        std::chrono::steady_clock::time_point t0_;
    };
};

// fn(begin, end) over [0, n) on up to `threads` host threads (the `threads` argument of grm::alignReads, Align.cpp:119:
// the reference spends it on aligning, here it packs the batch and writes the results back -- at 5 M reads/s on the
// device the ~0.3 us a single thread needs per read for that is what a caller would otherwise wait for).  Every index
// belongs to exactly one thread, so the outcome does not depend on `threads`; the first exception is rethrown.
template <typename F> void parallelFor(size_t n, unsigned threads, F&& fn)
{
    const size_t min_chunk = 512;
    const size_t nt = std::min<size_t>(threads ? threads : 1, (n + min_chunk - 1) / min_chunk);
    if (nt <= 1)
    {
        fn((size_t)0, n);
        return;
    }
    WorkerPool::instance().run(nt, [&](size_t t) { fn(n * t / nt, n * (t + 1) / nt); });
}

// page-locked staging (pg_host_alloc), grown geometrically and reused between batches: the library copies such
// buffers to and from the device directly instead of staging pageable memory once more
template <typename T> class PinnedBuf
{
public:
    PinnedBuf() = default;
    ~PinnedBuf() { pg_host_free(p_); }
    PinnedBuf(PinnedBuf const&) = delete;
    PinnedBuf& operator=(PinnedBuf const&) = delete;
    T* reserve(size_t n) // contents are not kept
    {
        if (n > cap_)
        {
            const size_t want = std::max(n, cap_ * 2);
            pg_host_free(p_);
            p_ = nullptr;
            cap_ = 0;
            void* q = nullptr;
            if (pg_host_alloc((uint64_t)(want * sizeof(T)), &q) != PG_OK || !q)
                throw std::runtime_error("paragraph_b200: cannot allocate " + std::to_string(want * sizeof(T)) + " bytes of page-locked memory");
            p_ = static_cast<T*>(q);
            cap_ = want;
        }
        return p_;
    }
    T* data() const { return p_; }
    size_t capacity() const { return cap_; }

private:
    T* p_ = nullptr;
    size_t cap_ = 0;
};
} // namespace detail

// RAII around pg_ctx; errors become std::runtime_error like the reference's error() (common/Error.hh:55-167)
class Engine
{
public:
    explicit Engine(int device = 0) : device_(device)
    {
        if (pg_create(device, &ctx_) != PG_OK || !ctx_)
            throw std::runtime_error("paragraph_b200: no usable sm_100 CUDA device (there is no CPU fallback)");
    }
    ~Engine() { pg_destroy(ctx_); }
    Engine(Engine const&) = delete;
    Engine& operator=(Engine const&) = delete;
    pg_ctx* get() const { return ctx_; }
    int device() const { return device_; }
    void check(int rc) const
    {
        if (rc != PG_OK)
            throw std::runtime_error(std::string("paragraph_b200: ") + pg_last_error(ctx_));
    }

    // Engines of finished aligners are kept per host thread and handed to the next aligner on that thread: the
    // reference builds its aligners per call (Align.cpp:96-110) and grmpy calls alignReads once per (sample, target)
    // with a few hundred reads (Workflow.cpp:108-146) -- creating a context, its streams and its device buffers each
    // time costs milliseconds, more than aligning those reads.  A pooled engine comes back with the default stages
    // and, unless the taker asks for it (keep_graph: GraphAligner, which compares graph_key), without graphs.
    static std::unique_ptr<Engine> take(int device, bool keep_graph = false)
    {
        auto& p = pool();
        for (size_t i = 0; i < p.size(); ++i)
            if (p[i]->device() == device)
            {
                std::unique_ptr<Engine> e = std::move(p[i]);
                p.erase(p.begin() + (std::ptrdiff_t)i);
                if (!keep_graph && !e->graph_key.empty())
                {
                    e->graph_key.clear();
                    e->check(pg_clear_graphs(e->get()));
                }
                return e;
            }
        return std::unique_ptr<Engine>(new Engine(device));
    }
    static void give(std::unique_ptr<Engine> e)
    {
        if (!e || pool().size() >= 4)
            return; // (destroyed)
        if ((!e->graph_key.empty() || pg_clear_graphs(e->get()) == PG_OK) && pg_set_stages(e->get(), 0, 1, 0) == PG_OK
            && pg_set_kmer_stage(e->get(), 0) == PG_OK)
            pool().push_back(std::move(e));
    }
    // Non-empty: the engine holds exactly the one graph (+ paths) this key spells out, registered by a GraphAligner.
    // grmpy aligns every sample to the same graphs (Workflow.cpp:108-146): the next aligner on this thread that sets
    // the same graph finds it on the device already.
    std::string graph_key;

    // staging of the current batch (one batch at a time per engine, like one aligner per thread in the reference)
    detail::PinnedBuf<char> blob;
    detail::PinnedBuf<int32_t> off;
    detail::PinnedBuf<pg_record> rec;
    detail::PinnedBuf<uint32_t> ops;

    // gather the bases of which[0..n) into blob / off; returns the bytes.  Both looks at the reads (lengths, then bytes)
    // run on `threads` threads -- at this point the reads are cold in the caches and a serial pass over them costs more
    // than the copy; the prefix sum between them runs over the offsets only.
    template <typename GetBases> size_t pack(size_t n, unsigned threads, GetBases&& bases_of)
    {
        int32_t* o = off.reserve(n + 1);
        o[0] = 0;
        std::atomic<bool> too_long{ false };
        detail::parallelFor(n, threads, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
            {
                const size_t len = bases_of(i).size();
                if (len > (size_t)INT_MAX)
                    too_long = true;
                o[i + 1] = (int32_t)len;
            }
        });
        size_t total = 0;
        for (size_t i = 0; i < n && !too_long; ++i)
        {
            total += (size_t)o[i + 1];
            if (total > (size_t)INT_MAX)
                too_long = true;
            o[i + 1] = (int32_t)total;
        }
        if (too_long)
            throw std::runtime_error("paragraph_b200: more than 2 GB of read bases in one batch");
        char* b = blob.reserve(total + 1);
        detail::parallelFor(n, threads, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
                std::memcpy(b + o[i], bases_of(i).data(), (size_t)(o[i + 1] - o[i]));
        });
        return total;
    }

private:
    static std::vector<std::unique_ptr<Engine>>& pool()
    {
        static thread_local std::vector<std::unique_ptr<Engine>> p;
        return p;
    }
    pg_ctx* ctx_ = nullptr;
    int device_ = 0;
};

// graphtools::reverseComplement (SequenceOperations.cpp:66-89): case-sensitive, non-ACGT -> 'N'
inline std::string reverseComplement(std::string s)
{
    for (char& c : s)
        c = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    std::reverse(s.begin(), s.end());
    return s;
}

// the same into a caller-owned string (applyRecord keeps one per thread: no allocation per read)
inline void reverseComplementInto(std::string const& in, std::string& out)
{
    const size_t n = in.size();
    out.resize(n);
    for (size_t i = 0; i < n; ++i) // (a select chain: the compiler vectorises it, unlike a table lookup)
    {
        const char c = in[n - 1 - i];
        out[i] = c == 'A' ? 'T' : c == 'C' ? 'G' : c == 'G' ? 'C' : c == 'T' ? 'A' : 'N';
    }
}

class GraphAligner
{
public:
    static const unsigned int AF_CIGAR = 0x01;
    static const unsigned int AF_BOTH_STRANDS = 0x02;
    static const unsigned int AF_REVERSE_GRAPH = 0x04;
    static const unsigned int AF_ALL = (unsigned int)-1;

    explicit GraphAligner(int device = 0) : engine_(Engine::take(device, true)) {}
    ~GraphAligner()
    {
        Engine::give(std::move(engine_));
        Engine::give(std::move(engine2_));
    }
    GraphAligner(GraphAligner const&) = delete;
    GraphAligner& operator=(GraphAligner const&) = delete;

    // GraphAligner::setGraph (GraphAligner.cpp:277-285); the reversed graph is derived by the engine
    template <typename GraphT> void setGraph(GraphT const* g)
    {
        blob_.clear();
        off_.assign(1, 0);
        ef_.clear();
        et_.clear();
        const int32_t n = (int32_t)g->numNodes();
        for (int32_t i = 0; i < n; ++i)
        {
            blob_ += g->nodeSeq((uint32_t)i);
            off_.push_back((int32_t)blob_.size());
            for (auto p : g->predecessors((uint32_t)i))
            {
                ef_.push_back((int32_t)p);
                et_.push_back(i);
            }
        }
        path_ptr_.clear();
        path_nodes_.clear();
        key_.clear();
        engine1_fresh_ = engine2_fresh_ = false; // registered when the next batch (or context()) needs it
    }
    // the stages in front of gssw and the paths the k-mer stage aligns to (pg_set_stages / pg_set_kmer_stage / pg_set_paths):
    // kept here so that a second engine can be brought to the same configuration
    void setStages(int path_kmer_len, bool graph_matching, bool nonuniq_second_chance, int kmer_len)
    {
        path_k_ = path_kmer_len;
        graph_on_ = graph_matching;
        second_ = nonuniq_second_chance;
        kmer_k_ = kmer_len;
    }
    void setPaths(std::vector<int32_t> ptr, std::vector<int32_t> nodes)
    {
        path_ptr_ = std::move(ptr);
        path_nodes_ = std::move(nodes);
        key_.clear();
        engine1_fresh_ = engine2_fresh_ = false;
    }

    // the loop `for read: alignRead(read, flags)` as one batch; writes the fields GraphAligner::alignRead writes
    // (GraphAligner.cpp:358-401).  ReadIt iterates over (smart) pointers to reads; empty reads are skipped.
    // A batch of a few thousand reads or more is cut in two halves that run on two engines from two host threads: the
    // packing and the write-back of one half (strings: reverse complements, CIGAR text) overlap the kernels of the other.
    template <typename ReadIt>
    void alignBatch(ReadIt begin, ReadIt end, unsigned flags = AF_ALL, std::vector<pg_record>* records_out = nullptr,
                    bool tolerate_unmapped = false, bool all_nonempty = false,
                    std::function<void(size_t, size_t)> const* post = nullptr) const
    {
        // post(lo, hi): called on the thread that just wrote reads [lo, hi) of the batch back (indices into the non-empty
        // reads of [begin, end)), after records_out holds their records -- the caller's per-read work (filter, tallies)
        // joins the pipeline instead of following it
        std::vector<ReadIt> which;
        for (ReadIt it = begin; it != end; ++it)
            if (all_nonempty || !(*it)->bases().empty())
                which.push_back(it);
        if (which.empty())
            return;
        const size_t n = which.size();
        if (records_out)
            records_out->resize(n);
        anchored_ = 0;
        ensureRegistered();
        // one part of the batch in two steps: submit = pack + H2D + kernels (returns with the kernels in flight),
        // collect = D2H + write-back into the reads
        auto submit = [&](Engine& e, size_t lo, size_t hi) -> size_t {
            const size_t m = hi - lo;
            detail::Phases::Scope ph(detail::Phases::PACK);
            e.check(pg_set_kmer_stage(e.get(), kmer_k_));
            e.check(pg_set_stages(e.get(), path_k_, graph_on_ ? 1 : 0, second_ ? 1 : 0));
            const size_t bytes = e.pack(m, threads_, [&](size_t i) -> std::string const& { return (*which[lo + i])->bases(); });
            e.rec.reserve(m);
            const size_t ops_cap = bytes + 16 * m + 64;
            e.ops.reserve(ops_cap);
            ph.next(detail::Phases::LAUNCH);
            e.check(pg_batch_upload(e.get(), (int32_t)m, e.blob.data(), e.off.data(), nullptr));
            e.check(pg_batch_run(e.get(), flags));
            return ops_cap;
        };
        auto collect = [&](Engine& e, size_t lo, size_t hi, size_t ops_cap) {
            const size_t m = hi - lo;
            pg_record* rec = e.rec.data();
            uint32_t* ops = e.ops.data();
            uint64_t used = 0;
            detail::Phases::Scope ph(detail::Phases::WAIT);
            e.check(pg_batch_download(e.get(), rec, ops, ops_cap, &used));
            ph.next(detail::Phases::APPLY);
            if (path_k_ > 0) // PathAligner::anchored of this part
            {
                uint64_t cnt[4] = { 0, 0, 0, 0 };
                e.check(pg_path_stats(e.get(), cnt, nullptr));
                anchored_ += cnt[1];
            }
            if (records_out) // e.g. for paragraph::DefaultReadFilter, which needs query_clipped
                std::copy(rec, rec + m, records_out->begin() + (std::ptrdiff_t)lo);
            detail::parallelFor(m, threads_, [&](size_t a, size_t b) {
                for (size_t i = a; i < b; ++i)
                    if (!(tolerate_unmapped && rec[i].status == 3)) // 3: no enabled stage mapped the read -- it stays UNMAPPED
                        applyRecord(**which[lo + i], rec[i], ops, flags, lo + i);
                if (post)
                    (*post)(lo + a, lo + b);
            });
        };
        if (n < pipeline_min_reads_)
        {
            const size_t cap = submit(*engine_, 0, n);
            collect(*engine_, 0, n, cap);
            return;
        }
        // Software pipeline over two engines (two contexts, two streams): part k + 1 is packed and submitted before part
        // k is collected, so the write-back of one part (strings: reverse complements, CIGAR text -- on all host
        // threads) runs while the kernels of the next are in flight.  Parts of pipeline_min_reads_ / 2 reads or more,
        // at most pipeline_max_parts_ of them.
        if (!engine2_)
            engine2_ = Engine::take(engine_->device(), true);
        if (!engine2_fresh_)
        {
            registerGraph(*engine2_);
            engine2_fresh_ = true;
        }
        size_t parts = std::max<size_t>(2, std::min<size_t>(pipeline_max_parts_, n / std::max<size_t>(pipeline_min_reads_ / 2, 1)));
        // (two parts: the first one may be the larger -- what follows the last download, the write-back of the last part,
        // is not hidden behind any kernel.  PGB_PIPELINE_SPLIT = "15,70,15": explicit part sizes in per cent -- a small
        // first part gets the device going early, a small last one leaves little write-back uncovered)
        std::vector<size_t> cut;
        if (!pipeline_split_.empty())
        {
            size_t acc = 0;
            cut.push_back(0);
            for (size_t pct : pipeline_split_)
            {
                acc += pct;
                cut.push_back(std::min(n, n * acc / 100));
            }
            cut.back() = n;
            parts = cut.size() - 1;
        }
        auto bound = [&](size_t k) {
            if (!cut.empty())
                return cut[std::min(k, cut.size() - 1)];
            return parts == 2 && k == 1 ? n * pipeline_first_pct_ / 100 : n * k / parts;
        };
        Engine* eng[2] = { engine_.get(), engine2_.get() };
        size_t cap[2] = { 0, 0 };
        cap[0] = submit(*eng[0], bound(0), bound(1));
        for (size_t k = 0; k < parts; ++k)
        {
            if (k + 1 < parts) // (its engine's buffers are free: part k - 1 was collected in the last iteration)
                cap[(k + 1) & 1] = submit(*eng[(k + 1) & 1], bound(k + 1), bound(k + 2));
            collect(*eng[k & 1], bound(k), bound(k + 1), cap[k & 1]);
        }
    }
    uint64_t lastAnchored() const { return anchored_; } // reads of the last batch the exact-match stage anchored
    // reads from which a batch is cut in two pipelined halves (0 = never)
    void setPipelineMinReads(size_t n) { pipeline_min_reads_ = n ? n : (size_t)-1; }
    void setPipelineMaxParts(size_t n) { pipeline_max_parts_ = n < 2 ? 2 : n; }

    // host threads for packing a batch and writing its results back (the `threads` of grm::alignReads); results are the
    // same for any value
    void setThreads(unsigned threads) { threads_ = threads ? threads : 1; }

    // write one record back into a read: the fields GraphAligner::alignRead sets (GraphAligner.cpp:358-401)
    template <typename ReadT>
    static void applyRecord(ReadT& read, pg_record const& r, const uint32_t* ops, unsigned flags, size_t index)
    {
        static thread_local std::string tmp; // scratch for the rewritten bases / quals / CIGAR string
        if (r.status != 0)
            throw std::runtime_error("paragraph_b200: traceback failed for read " + std::to_string(index));
        if (r.mapped_by == PG_STAGE_PATH_ID) // what PathAligner::alignRead writes (PathAligner.cpp:124-161)
        {
            read.set_is_graph_reverse_strand(r.chose_reverse != 0);
            if (r.chose_reverse) // quals stay as they are
            {
                reverseComplementInto(read.bases(), tmp);
                read.set_bases(tmp);
            }
            read.set_graph_alignment_score(r.score);
            read.set_graph_pos(r.graph_pos);
            formatCigar(r, ops, tmp);
            read.set_graph_cigar(tmp);
            read.set_is_graph_alignment_unique(r.unique != 0);
            read.set_graph_mapq(r.unique ? 60 : 0);
            return;
        }
        // the stages in front had reverse-complemented the bases (each one graphtools::reverseComplement, quals untouched)
        const int flips = (r.mapped_by == PG_STAGE_GSSW_REV_ID || r.mapped_by == PG_STAGE_KMER_REV_ID) ? 1
            : (r.mapped_by == PG_STAGE_GSSW_REV2_ID ? 2 : 0);
        for (int f = 0; f < flips; ++f)
        {
            reverseComplementInto(read.bases(), tmp);
            read.set_bases(tmp);
        }
        if (r.mapped_by == PG_STAGE_KMER_ID || r.mapped_by == PG_STAGE_KMER_REV_ID) // KmerAligner.cpp:423-478
        {
            read.set_graph_pos(r.graph_pos);
            read.set_is_graph_reverse_strand(read.is_reverse_strand() != (r.chose_reverse != 0));
            if (r.chose_reverse) // quals stay as they are
            {
                reverseComplementInto(read.bases(), tmp);
                read.set_bases(tmp);
            }
            formatCigar(r, ops, tmp);
            read.set_graph_cigar(tmp);
            read.set_graph_alignment_score(r.score);
            read.set_graph_mapq(r.unique ? 60 : 0);
            read.set_is_graph_alignment_unique(r.unique != 0);
            return;
        }
        read.set_is_graph_reverse_strand(read.is_reverse_strand() != (r.chose_reverse != 0)); // :358-359
        if (r.chose_reverse) // :375-378
        {
            reverseComplementInto(read.bases(), tmp);
            read.set_bases(tmp);
            tmp.assign(read.quals().rbegin(), read.quals().rend());
            read.set_quals(tmp);
        }
        read.set_graph_pos(r.graph_pos);
        read.set_graph_alignment_score(r.score);
        read.set_is_graph_alignment_unique(r.unique != 0);
        read.set_graph_mapq(r.unique ? 60 : 0);
        if (flags & AF_CIGAR)
        {
            formatCigar(r, ops, tmp);
            read.set_graph_cigar(tmp);
        }
    }

    // extractCigar (GraphAligner.cpp:88-108) of one record into `out`: "<node>[<len><op>...]..." from the op words
    // (the layout of include/pg_align.h; same text as pg_format_cigar, without a round trip through the C-ABI)
    static void formatCigar(pg_record const& r, const uint32_t* ops, std::string& out)
    {
        out.resize((size_t)12 * r.cigar_len + 16);
        char* const base = &out[0];
        char* p = base;
        auto num = [&p](uint32_t v) {
            char tmp[12];
            int n = 0;
            do
            {
                tmp[n++] = (char)('0' + v % 10);
                v /= 10;
            } while (v);
            while (n)
                *p++ = tmp[--n];
        };
        static const char OPS[8] = { 'M', 'X', 'N', 'I', 'D', 'S', '?', '?' };
        uint32_t node = 0xFFFFFFFFu;
        for (uint32_t i = 0; i < r.cigar_len; ++i)
        {
            const uint32_t w = ops[r.cigar_off + i];
            if (PG_CIGAR_NODE(w) != node)
            {
                if (node != 0xFFFFFFFFu)
                    *p++ = ']';
                node = PG_CIGAR_NODE(w);
                num(node);
                *p++ = '[';
            }
            if (PG_CIGAR_OP(w) != PG_OP_NONE)
            {
                num(PG_CIGAR_LEN(w));
                *p++ = OPS[PG_CIGAR_OP(w)];
            }
        }
        if (node != 0xFFFFFFFFu)
            *p++ = ']';
        out.resize((size_t)(p - base));
    }

    pg_ctx* context() const
    {
        ensureRegistered();
        return engine_->get();
    }
    void check(int rc) const { engine_->check(rc); }

    template <typename ReadT> void alignRead(ReadT& read, unsigned flags = AF_ALL) const
    {
        ReadT* p = &read;
        alignBatch(&p, &p + 1, flags);
    }

    // GraphAligner::align (GraphAligner.cpp:287-296)
    std::string align(const std::string& read, int& mapq, int& position, int& score) const
    {
        Read tmp;
        tmp.set_bases(read);
        alignRead(tmp, AF_CIGAR);
        mapq = tmp.graph_mapq();
        position = tmp.graph_pos();
        score = tmp.graph_alignment_score();
        return tmp.graph_cigar();
    }

private:
    void ensureRegistered() const
    {
        if (!engine1_fresh_)
        {
            registerGraph(*engine_);
            engine1_fresh_ = true;
        }
    }
    // brings an engine to this aligner's graph and paths; nothing to do when it holds them already (Engine::graph_key)
    void registerGraph(Engine& e) const
    {
        if (off_.size() < 2)
            throw std::runtime_error("paragraph_b200: GraphAligner used before setGraph");
        if (key_.empty())
        {
            auto put = [this](const void* p, size_t bytes) {
                const uint64_t n = bytes;
                key_.append(reinterpret_cast<const char*>(&n), sizeof n);
                key_.append(static_cast<const char*>(p), bytes);
            };
            put(blob_.data(), blob_.size());
            put(off_.data(), off_.size() * sizeof(int32_t));
            put(ef_.data(), ef_.size() * sizeof(int32_t));
            put(et_.data(), et_.size() * sizeof(int32_t));
            put(path_ptr_.data(), path_ptr_.size() * sizeof(int32_t));
            put(path_nodes_.data(), path_nodes_.size() * sizeof(int32_t));
        }
        if (e.graph_key == key_)
            return;
        e.graph_key.clear();
        e.check(pg_clear_graphs(e.get()));
        int32_t sid = -1;
        e.check(pg_add_graph(e.get(), (int32_t)off_.size() - 1, blob_.data(), off_.data(), (int32_t)ef_.size(), ef_.data(), et_.data(),
                             &sid));
        if (!path_ptr_.empty())
            e.check(pg_set_paths(e.get(), 0, (int32_t)path_ptr_.size() - 1, path_ptr_.data(), path_nodes_.data()));
        e.graph_key = key_;
    }
    std::unique_ptr<Engine> engine_;
    mutable std::unique_ptr<Engine> engine2_; // second half of a pipelined batch
    mutable bool engine1_fresh_ = false, engine2_fresh_ = false; // the engine holds the current graph and paths
    mutable std::string key_;                                    // what Engine::graph_key is compared with
    mutable std::atomic<uint64_t> anchored_{ 0 };
    unsigned threads_ = 1;
    // tuning knobs, PGB_PIPELINE_MIN_READS / PGB_PIPELINE_PARTS in the environment override the defaults
    static size_t envKnob(const char* name, size_t dflt)
    {
        const char* e = std::getenv(name);
        return e && *e ? (size_t)std::strtoull(e, nullptr, 10) : dflt;
    }
    size_t pipeline_min_reads_ = envKnob("PGB_PIPELINE_MIN_READS", 4096);
    size_t pipeline_max_parts_ = std::max<size_t>(2, envKnob("PGB_PIPELINE_PARTS", 2));
    size_t pipeline_first_pct_ = std::min<size_t>(90, std::max<size_t>(10, envKnob("PGB_PIPELINE_FIRST_PCT", 50)));
    static std::vector<size_t> envSplit(const char* name)
    {
        std::vector<size_t> v;
        const char* e = std::getenv(name);
        if (!e || !*e)
            return v;
        size_t sum = 0;
        for (const char* p = e; *p;)
        {
            char* end = nullptr;
            const unsigned long x = std::strtoul(p, &end, 10);
            if (end == p)
                break;
            v.push_back((size_t)x);
            sum += x;
            p = *end ? end + 1 : end;
        }
        if (v.size() < 2 || sum != 100)
            v.clear();
        return v;
    }
    std::vector<size_t> pipeline_split_ = envSplit("PGB_PIPELINE_SPLIT");
    std::string blob_;
    std::vector<int32_t> off_{ 0 }, ef_, et_, path_ptr_, path_nodes_;
    int path_k_ = 0, kmer_k_ = 0;
    bool graph_on_ = true, second_ = false;
};

// Many sites in ONE launch sequence.  grmpy hands (sample, graph) pairs to threads one at a time
// (src/c++/lib/grmpy/Workflow.cpp:108-146) and a 30x site has only ~200 reads -- far too few to fill a B200 -- so the
// GPU caller collects sites and flushes them together; results are the same as calling alignReads per site.
template <typename ReadPtrT> class MultiSiteAligner
{
public:
    explicit MultiSiteAligner(int device = 0, unsigned flags = GraphAligner::AF_ALL) : engine_(Engine::take(device)), flags_(flags) {}
    ~MultiSiteAligner() { Engine::give(std::move(engine_)); }
    MultiSiteAligner(MultiSiteAligner const&) = delete;
    MultiSiteAligner& operator=(MultiSiteAligner const&) = delete;

    // path_sequence_matching of alignAndDisambiguate (`paragraph` switches it on by default, main/paragraph.cpp:60):
    // the exact-match stage (grm::PathAligner, k-mer size 32 in the reference) runs in front of gssw in alignAndCount().
    // A non-unique exact match is what the default filter chain rejects right after that stage; it then gets its second
    // chance in gssw on the device (pg_set_stages: nonuniq_second_chance) iff remove_nonuniq_reads is set.
    void setPathMatching(int kmer_len) { path_kmer_ = kmer_len; }
    // host threads for packing the batch and writing the results back (results do not depend on it)
    void setThreads(unsigned threads) { threads_ = threads ? threads : 1; }

    // register a site: its graph and its reads (the vector is updated in place by run(), like grm::alignReads does)
    template <typename GraphT> void addSite(GraphT const* g, std::vector<ReadPtrT>* reads)
    {
        std::string blob;
        std::vector<int32_t> off{ 0 }, ef, et;
        const int32_t n = (int32_t)g->numNodes();
        for (int32_t i = 0; i < n; ++i)
        {
            blob += g->nodeSeq((uint32_t)i);
            off.push_back((int32_t)blob.size());
            for (auto p : g->predecessors((uint32_t)i))
            {
                ef.push_back((int32_t)p);
                et.push_back(i);
            }
        }
        int32_t sid = -1;
        engine_->check(pg_add_graph(engine_->get(), n, blob.data(), off.data(), (int32_t)ef.size(), ef.data(), et.data(), &sid));
        Site site{ sid, reads, {}, {}, {} };
        // names and path-family labels for the counting stage (Graph::edgeLabels; at most 64 labels per site)
        std::vector<uint64_t> masks(ef.size(), 0);
        std::map<std::string, int> label_id;
        for (int32_t i = 0; i < n; ++i)
            site.node_names.push_back(g->nodeName((uint32_t)i));
        for (size_t e = 0; e < ef.size(); ++e)
        {
            site.edges.emplace_back(ef[e], et[e]);
            for (auto const& label : g->edgeLabels((uint32_t)ef[e], (uint32_t)et[e]))
            {
                auto it = label_id.find(label);
                if (it == label_id.end())
                {
                    if (site.labels.size() == 64)
                        throw std::runtime_error("paragraph_b200: more than 64 sequence labels on one graph");
                    it = label_id.emplace(label, (int)site.labels.size()).first;
                    site.labels.push_back(label);
                }
                masks[e] |= 1ull << it->second;
            }
        }
        engine_->check(pg_set_edge_labels(engine_->get(), sid, masks.data()));
        sites_.push_back(std::move(site));
    }

    // alignAndDisambiguate's core for every registered site in one go (Disambiguation.cpp:152-299 without the JSON
    // plumbing): align, apply the default filter chain, disambiguate and count -- filters, supports and counts are
    // computed on the device from the op arena (pg_batch_count).  Each site's read vector keeps its MAPPED reads, with
    // graph_{nodes,edges,sequences}_supported filled in like disambiguateReads does; returns one SiteCounts per site
    // in addSite order.
    std::vector<paragraph::SiteCounts> alignAndCount(paragraph::CountOptions const& opt = paragraph::CountOptions())
    {
        std::vector<int32_t> site, fragment;
        std::vector<uint8_t> is_rev;
        std::vector<ReadPtrT*> which;
        int32_t next_fragment = 0;
        size_t total_nodes = 0, total_edges = 0;
        for (auto& s : sites_)
        {
            std::unordered_map<std::string, int32_t> frag_id; // readsToFragments groups by fragment_id (Fragment.cpp:165-181)
            for (auto& r : *s.reads)
            {
                if (r->bases().empty())
                    continue;
                site.push_back(s.id);
                is_rev.push_back(r->is_reverse_strand() ? 1 : 0);
                auto it = frag_id.find(r->fragment_id());
                if (it == frag_id.end())
                    it = frag_id.emplace(r->fragment_id(), next_fragment++).first;
                fragment.push_back(it->second);
                which.push_back(&r);
            }
            total_nodes += s.node_names.size();
            total_edges += s.edges.size();
        }
        std::vector<paragraph::SiteCounts> out(sites_.size());
        std::vector<pg_read_support> sup(which.size());
        std::vector<pg_count4> nc(total_nodes + 1), ec(total_edges + 1);
        std::vector<uint32_t> path, fam(1 << 16);
        uint64_t fam_used = 0;
        if (!which.empty())
        {
            Engine& e = *engine_;
            const size_t bytes = e.pack(which.size(), threads_, [&](size_t i) -> std::string const& { return (**which[i]).bases(); });
            pg_record* rec = e.rec.reserve(which.size());
            const size_t ops_cap = bytes + 16 * which.size() + 64;
            uint32_t* ops = e.ops.reserve(ops_cap);
            uint64_t used = 0, path_used = 0;
            engine_->check(pg_set_stages(engine_->get(), path_kmer_, 1, opt.remove_nonuniq_reads ? 1 : 0));
            engine_->check(pg_align_batch(engine_->get(), (int32_t)which.size(), e.blob.data(), e.off.data(), site.data(), flags_,
                                          rec, ops, ops_cap, &used));
            path.resize(used + 1);
            pg_count_params prm{ opt.remove_nonuniq_reads ? 1 : 0, opt.use_support_filters ? 1 : 0, opt.bad_align_frac,
                                 opt.family_slots, 0 };
            for (int attempt = 0;; ++attempt)
            {
                const int rc = pg_batch_count(engine_->get(), fragment.data(), is_rev.data(), &prm, sup.data(), path.data(),
                                              path.size(), &path_used, nc.data(), nc.size(), ec.data(), ec.size(), fam.data(),
                                              fam.size(), &fam_used);
                if (rc == PG_E_CAPACITY && attempt == 0 && fam_used > fam.size())
                {
                    fam.resize(fam_used);
                    continue;
                }
                engine_->check(rc);
                break;
            }
            std::unordered_map<int32_t, size_t> site_index;
            for (size_t k = 0; k < sites_.size(); ++k)
                site_index[sites_[k].id] = k;
            for (size_t i = 0; i < which.size(); ++i)
                out[site_index.at(site[i])].invalid_alignments += sup[i].verdict == PG_V_INVALID;
            detail::parallelFor(which.size(), threads_, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
            {
                auto& read = **which[i];
                typedef typename std::remove_reference<decltype(read)>::type ReadT;
                const Site& s = sites_[site_index.at(site[i])];
                GraphAligner::applyRecord(read, rec[i], ops, flags_, i);
                read.clear_graph_nodes_supported();
                read.clear_graph_edges_supported();
                read.clear_graph_sequences_supported();
                if (sup[i].verdict != PG_V_MAPPED)
                {
                    read.set_graph_mapping_status(ReadT::BAD_ALIGN); // Disambiguation.cpp:184 / CompositeAligner.cpp:165-169
                    continue;
                }
                read.set_graph_mapping_status(ReadT::MAPPED);
                // std::set order of the reference: nodes by id, edges by (from name, to name), labels by name
                std::set<std::pair<std::string, std::string>> edges;
                for (uint32_t k = 0; k < sup[i].path_len; ++k)
                {
                    const uint32_t w = path[sup[i].path_off + k];
                    if (w & PG_SUP_NODE)
                        read.add_graph_nodes_supported(s.node_names[w & PG_SUP_NODE_MASK]);
                    if (k > 0 && (w & PG_SUP_EDGE))
                        edges.emplace(s.node_names[path[sup[i].path_off + k - 1] & PG_SUP_NODE_MASK], s.node_names[w & PG_SUP_NODE_MASK]);
                }
                for (auto const& e : edges)
                    read.add_graph_edges_supported(e.first + "_" + e.second);
                std::set<std::string> seqs;
                for (size_t k = 0; k < s.labels.size(); ++k)
                    if ((sup[i].sequences >> k) & 1)
                        seqs.insert(s.labels[k]);
                for (auto const& l : seqs)
                    read.add_graph_sequences_supported(l);
            }
            });
        }
        // tables: rows are site-major in addSite order (this object registers nothing else on the context)
        size_t nb = 0, eb = 0;
        std::vector<std::pair<size_t, size_t>> bases;
        for (size_t k = 0; k < sites_.size(); ++k)
        {
            const Site& s = sites_[k];
            bases.emplace_back(nb, eb);
            for (size_t v = 0; v < s.node_names.size(); ++v)
                if (nc[nb + v].fragments)
                    out[k].read_counts_by_node[s.node_names[v]] = conv(nc[nb + v]);
            for (size_t e = 0; e < s.edges.size(); ++e)
                if (ec[eb + e].fragments)
                    out[k].read_counts_by_edge[edgeName(s, e)] = conv(ec[eb + e]);
            nb += s.node_names.size();
            eb += s.edges.size();
        }
        for (uint64_t w = 0; w < fam_used;)
        {
            const uint32_t sid = fam[w], n = fam[w + 1];
            const uint64_t mask = (uint64_t)fam[w + 2] | ((uint64_t)fam[w + 3] << 32);
            size_t k = 0;
            while (k < sites_.size() && sites_[k].id != (int32_t)sid)
                ++k;
            const Site& s = sites_.at(k);
            std::set<std::string> names;
            for (size_t b = 0; b < s.labels.size(); ++b)
                if ((mask >> b) & 1)
                    names.insert(s.labels[b]);
            std::string key;
            for (auto const& nme : names)
                key += (key.empty() ? "" : ",") + nme;
            auto& dst = out[k].read_counts_by_sequence[key];
            const pg_count4* rows = reinterpret_cast<const pg_count4*>(&fam[w + 4]);
            dst["total"] = conv(rows[0]);
            for (size_t v = 0; v < s.node_names.size(); ++v)
                if (rows[1 + v].fragments)
                    dst[s.node_names[v]] = conv(rows[1 + v]);
            for (size_t e = 0; e < s.edges.size(); ++e)
                if (rows[1 + s.node_names.size() + e].fragments)
                    dst[edgeName(s, e)] = conv(rows[1 + s.node_names.size() + e]);
            w += 4 + 4ull * n;
        }
        keepMapped();
        return out;
    }

    // align every registered site in one batch, apply the filter, keep MAPPED reads per site (Align.cpp:72-84,155)
    template <typename FilterT> void run(FilterT filter)
    {
        std::vector<int32_t> site;
        std::vector<ReadPtrT*> which;
        for (auto& s : sites_)
            for (auto& r : *s.reads)
            {
                if (r->bases().empty())
                    continue;
                site.push_back(s.id);
                which.push_back(&r);
            }
        if (!which.empty())
        {
            Engine& e = *engine_;
            const size_t bytes = e.pack(which.size(), threads_, [&](size_t i) -> std::string const& { return (**which[i]).bases(); });
            pg_record* rec = e.rec.reserve(which.size());
            const size_t ops_cap = bytes + 16 * which.size() + 64;
            uint32_t* ops = e.ops.reserve(ops_cap);
            uint64_t used = 0;
            engine_->check(pg_set_stages(engine_->get(), 0, 1, 0)); // run(filter): gssw stage only
            engine_->check(pg_align_batch(engine_->get(), (int32_t)which.size(), e.blob.data(), e.off.data(), site.data(), flags_,
                                          rec, ops, ops_cap, &used));
            detail::parallelFor(which.size(), threads_, [&](size_t lo, size_t hi) {
                for (size_t i = lo; i < hi; ++i)
                    GraphAligner::applyRecord(**which[i], rec[i], ops, flags_, i);
            });
            for (size_t i = 0; i < which.size(); ++i) // the filter callback: calling thread, input order
            {
                auto& read = **which[i];
                typedef typename std::remove_reference<decltype(read)>::type ReadT;
                read.set_graph_mapping_status(ReadT::MAPPED); // CompositeAligner.cpp:156
                if (filter && filter(read))
                    read.set_graph_mapping_status(ReadT::BAD_ALIGN);
            }
        }
        keepMapped();
    }

private:
    struct Site
    {
        int32_t id;
        std::vector<ReadPtrT>* reads;
        std::vector<std::string> node_names, labels;
        std::vector<std::pair<int32_t, int32_t>> edges; // in the order given to pg_add_graph
    };
    static paragraph::Count4 conv(pg_count4 const& c)
    {
        paragraph::Count4 r;
        r.fragments = c.fragments;
        r.reads = c.reads;
        r.fwd = c.fwd;
        r.rev = c.rev;
        return r;
    }
    static std::string edgeName(Site const& s, size_t e)
    {
        return s.node_names[(size_t)s.edges[e].first] + "_" + s.node_names[(size_t)s.edges[e].second];
    }
    void keepMapped()
    {
        for (auto& s : sites_)
        {
            std::vector<ReadPtrT> kept;
            for (auto& r : *s.reads)
            {
                typedef typename std::remove_reference<decltype(*r)>::type ReadT;
                if (!r->bases().empty() && r->graph_mapping_status() == ReadT::MAPPED)
                    kept.emplace_back(std::move(r));
            }
            s.reads->swap(kept);
        }
        sites_.clear();
        engine_->check(pg_clear_graphs(engine_->get()));
    }
    std::unique_ptr<Engine> engine_;
    unsigned flags_;
    int path_kmer_ = 0;
    unsigned threads_ = 1;
    std::vector<Site> sites_;
};

// Site-level double buffering.  grmpy walks its targets one by one (Workflow.cpp:108-146):
// extract the reads of a site, align, count, write JSON.  Here the caller keeps handing over (graph, reads) pairs; they
// are collected into batches of about `batch_reads` reads, and a full batch runs MultiSiteAligner::alignAndCount on
// one of TWO engines (contexts, streams, page-locked staging) in a worker thread while the caller fills the batch of
// the other one -- read extraction, packing and result write-back of one batch overlap the device time of the other
// (the overlap bench.py's e2e_two_contexts leg measures).  Results are those of one MultiSiteAligner over all sites.
template <typename ReadPtrT> class SitePipeline
{
public:
    explicit SitePipeline(int device = 0, unsigned flags = GraphAligner::AF_ALL, size_t batch_reads = 1 << 16,
                          unsigned threads = 1, paragraph::CountOptions const& opt = paragraph::CountOptions())
        : batch_reads_(batch_reads ? batch_reads : 1), opt_(opt)
    {
        for (auto& s : slots_)
        {
            s.aligner.reset(new MultiSiteAligner<ReadPtrT>(device, flags));
            s.aligner->setThreads(threads);
        }
    }
    ~SitePipeline()
    {
        for (auto& s : slots_) // never leave a worker behind that still writes into the caller's reads
            if (s.done.valid())
                s.done.wait();
    }
    void setPathMatching(int kmer_len)
    {
        for (auto& s : slots_)
            s.aligner->setPathMatching(kmer_len);
    }
    // queue a site; `reads` must stay alive (and untouched) until finish() -- it is updated in place like
    // MultiSiteAligner does.  May block while the other batch is still running.
    template <typename GraphT> void addSite(GraphT const* g, std::vector<ReadPtrT>* reads)
    {
        Slot& s = slots_[cur_];
        if (s.n_sites == 0)
            s.first_site = n_sites_;
        s.aligner->addSite(g, reads);
        ++s.n_sites;
        ++n_sites_;
        s.n_reads += reads->size();
        if (s.n_reads >= batch_reads_)
            flush();
    }
    // run what is left and wait; one SiteCounts per site in addSite order
    std::vector<paragraph::SiteCounts> finish()
    {
        flush();
        collect(slots_[0]);
        collect(slots_[1]);
        std::vector<paragraph::SiteCounts> out;
        out.swap(results_);
        n_sites_ = 0;
        return out;
    }

private:
    struct Slot
    {
        std::unique_ptr<MultiSiteAligner<ReadPtrT>> aligner;
        std::future<std::vector<paragraph::SiteCounts>> done;
        size_t first_site = 0, n_sites = 0, n_reads = 0;
    };
    void flush()
    {
        Slot& s = slots_[cur_];
        if (s.n_sites == 0)
            return;
        MultiSiteAligner<ReadPtrT>* a = s.aligner.get();
        const paragraph::CountOptions opt = opt_;
        s.done = std::async(std::launch::async, [a, opt] { return a->alignAndCount(opt); });
        cur_ ^= 1;
        collect(slots_[cur_]); // the engine the next sites go to must be idle
    }
    void collect(Slot& s)
    {
        if (!s.done.valid())
            return;
        std::vector<paragraph::SiteCounts> r = s.done.get(); // rethrows what the worker threw
        if (results_.size() < s.first_site + r.size())
            results_.resize(s.first_site + r.size());
        for (size_t k = 0; k < r.size(); ++k)
            results_[s.first_site + k] = std::move(r[k]);
        s.n_sites = s.n_reads = 0;
    }
    Slot slots_[2];
    int cur_ = 0;
    size_t n_sites_ = 0;
    const size_t batch_reads_;
    const paragraph::CountOptions opt_;
    std::vector<paragraph::SiteCounts> results_;
};

// Sites sharded over the GPUs of one box.  grmpy's workflow lets `threads` host threads pull (sample, graph) pairs off a
// shared list until it is empty (src/c++/lib/grmpy/Workflow.cpp:108-146); sites are independent, so on a multi-GPU box
// the same list is cut into one shard per device: longest-processing-time-first by cost = read bases x graph columns
// (what a site costs the fill kernel), one host thread per device driving a SitePipeline (two engines, double
// buffered) over its shard.  No device talks to another; results come back in addSite order and are those of one
// MultiSiteAligner over all sites.  Naming a device twice gives it two shards (two more engines on it).
template <typename ReadPtrT> class ShardedAligner
{
public:
    explicit ShardedAligner(std::vector<int> devices, unsigned flags = GraphAligner::AF_ALL, size_t batch_reads = 1 << 16,
                            unsigned threads_per_device = 1, paragraph::CountOptions const& opt = paragraph::CountOptions())
        : devices_(std::move(devices)), flags_(flags), batch_reads_(batch_reads), threads_(threads_per_device), opt_(opt)
    {
        if (devices_.empty())
            throw std::runtime_error("paragraph_b200: ShardedAligner needs at least one device");
    }
    void setPathMatching(int kmer_len) { path_kmer_ = kmer_len; }

    // queue a site; graph and reads must stay alive until run() returns (reads are updated in place)
    template <typename GraphT> void addSite(GraphT const* g, std::vector<ReadPtrT>* reads)
    {
        uint64_t cols = 0, bases = 0;
        for (size_t i = 0; i < g->numNodes(); ++i)
            cols += g->nodeSeq((uint32_t)i).size();
        for (auto const& r : *reads)
            bases += r->bases().size();
        Item it;
        it.cost = cols * bases;
        it.feed = [g, reads](SitePipeline<ReadPtrT>& p) { p.addSite(g, reads); };
        items_.push_back(std::move(it));
    }

    // shard index of every queued site (LPT: sites by falling cost, each to the least loaded shard so far)
    std::vector<int> partition() const
    {
        std::vector<size_t> order(items_.size());
        for (size_t i = 0; i < order.size(); ++i)
            order[i] = i;
        std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return items_[a].cost > items_[b].cost; });
        std::vector<uint64_t> load(devices_.size(), 0);
        std::vector<int> shard(items_.size(), 0);
        for (size_t i : order)
        {
            const size_t d = (size_t)(std::min_element(load.begin(), load.end()) - load.begin());
            shard[i] = (int)d;
            load[d] += items_[i].cost + 1;
        }
        return shard;
    }

    // align + count every queued site; one SiteCounts per site in addSite order
    std::vector<paragraph::SiteCounts> run()
    {
        const std::vector<int> shard = partition();
        std::vector<paragraph::SiteCounts> out(items_.size());
        std::vector<std::thread> pool;
        std::exception_ptr err;
        std::mutex m;
        for (size_t d = 0; d < devices_.size(); ++d)
            pool.emplace_back([&, d] {
                try
                {
                    std::vector<size_t> mine;
                    for (size_t i = 0; i < items_.size(); ++i)
                        if (shard[i] == (int)d)
                            mine.push_back(i);
                    if (mine.empty())
                        return;
                    SitePipeline<ReadPtrT> pipe(devices_[d], flags_, batch_reads_, threads_, opt_);
                    pipe.setPathMatching(path_kmer_);
                    for (size_t i : mine)
                        items_[i].feed(pipe);
                    std::vector<paragraph::SiteCounts> r = pipe.finish();
                    for (size_t k = 0; k < mine.size(); ++k)
                        out[mine[k]] = std::move(r[k]); // disjoint slots per shard
                }
                catch (...)
                {
                    std::lock_guard<std::mutex> lock(m);
                    if (!err)
                        err = std::current_exception();
                }
            });
        for (auto& th : pool)
            th.join();
        items_.clear();
        if (err)
            std::rethrow_exception(err);
        return out;
    }

private:
    struct Item
    {
        uint64_t cost;
        std::function<void(SitePipeline<ReadPtrT>&)> feed;
    };
    std::vector<int> devices_;
    unsigned flags_;
    size_t batch_reads_;
    unsigned threads_;
    paragraph::CountOptions opt_;
    int path_kmer_ = 0;
    std::vector<Item> items_;
};

template <typename ReadT> using ReadFilterT = std::function<bool(ReadT&)>; // include/grm/Filter.hh:36

class CompositeAligner
{
public:
    CompositeAligner(bool pathMatching, bool graphMatching, bool klibMatching, bool kmerMatching,
                     unsigned graphAlignmentFlags = GraphAligner::AF_ALL, int device = 0, int pathKmerSize = 32,
                     int kmerSize = 16)
        : pathMatching_(pathMatching), graphMatching_(graphMatching), kmerMatching_(kmerMatching), flags_(graphAlignmentFlags),
          pathKmerSize_(pathKmerSize), kmerSize_(kmerSize), graphAligner_(device)
    {
        if (klibMatching)
            throw std::runtime_error("paragraph_b200: the path, kmer and gssw stages run on the GPU; "
                                     "klib matching must stay on the reference's CPU aligner");
    }
    // CompositeAligner::setGraph (CompositeAligner.cpp:52-76): the k-mer stage aligns to the graph's paths
    template <typename GraphT, typename PathListT> void setGraph(GraphT const* graph, PathListT const& paths)
    {
        graphAligner_.setGraph(graph);
        if (kmerMatching_)
        {
            std::vector<int32_t> ptr{ 0 }, nodes;
            for (auto const& p : paths)
            {
                for (auto v : p.nodeIds())
                    nodes.push_back((int32_t)v);
                ptr.push_back((int32_t)nodes.size());
            }
            if (nodes.empty())
                nodes.push_back(0);
            graphAligner_.setPaths(std::move(ptr), std::move(nodes));
        }
    }
    void setThreads(unsigned threads)
    {
        threads_ = threads ? threads : 1;
        graphAligner_.setThreads(threads);
    }

    // CompositeAligner::alignRead for a whole range (CompositeAligner.cpp:78-176).  The reference runs, per read: the
    // exact-match stage (:82-95), the filter on its result with a second chance in the later stages for a rejected
    // read (:97-103), the k-mer stage (:105-126; a read it maps but not uniquely leaves it BAD_ALIGN and goes on), the
    // gssw stage (:146-175: every read it sees becomes MAPPED, then the filter may turn it into BAD_ALIGN).  Here every
    // stage is one launch over the batch, so the cascade runs in rounds: all enabled stages over all reads; then the
    // later stages over the reads whose exact match the filter rejected; then gssw over the reads whose k-mer
    // alignment it rejected.  Counters as in the reference.
    // reset_status: every non-empty read starts UNMAPPED (what grm::alignReads does first, Align.cpp:72-78)
    template <typename ReadIt, typename FilterT>
    void alignReads(ReadIt begin, ReadIt end, FilterT filter, bool reset_status = false)
    {
        typedef typename std::remove_reference<decltype(**begin)>::type ReadT;
        // the reads that take part: the look at each read (cold in the caches at this point) on the host threads, then
        // one pass over the flags
        std::vector<ReadT*> all;
        for (ReadIt it = begin; it != end; ++it)
            all.push_back(&**it);
        std::vector<uint8_t> takes_part(all.size());
        detail::parallelFor(all.size(), threads_, [&](size_t lo, size_t hi) {
            for (size_t i = lo; i < hi; ++i)
                if ((takes_part[i] = !all[i]->bases().empty()) && reset_status)
                    all[i]->set_graph_mapping_status(ReadT::UNMAPPED);
        });
        std::vector<ReadT*> todo;
        todo.reserve(all.size());
        for (size_t i = 0; i < all.size(); ++i)
            if (takes_part[i])
                todo.push_back(all[i]);
        attempted_ += (unsigned)todo.size();
        // stage sets of the rounds: {path?, kmer?, gssw?}
        bool path_on = pathMatching_, kmer_on = kmerMatching_;
        while (!todo.empty() && (path_on || kmer_on || graphMatching_))
        {
            graphAligner_.setStages(path_on ? pathKmerSize_ : 0, graphMatching_, false, kmer_on ? kmerSize_ : 0);
            std::vector<pg_record> rec;
            std::vector<ReadT*> after_path, after_kmer; // rejected by the filter right after that stage
            // statuses, filter and counters: on the `threads` host threads in contiguous chunks of the batch, right after a
            // chunk's records were written into its reads (alignBatch's `post`), as the reference's alignReads runs its
            // filter callback from its worker threads (Align.cpp:119-153); with one thread the callback sees the reads in
            // input order.  The chunks' tallies are merged in chunk order.
            struct Tally
            {
                size_t lo = 0;
                unsigned mappedPath = 0, mappedKmers = 0, mappedSw = 0, filtered = 0;
                std::vector<ReadT*> after_path, after_kmer;
            };
            std::vector<Tally> tallies;
            std::mutex tallies_mutex;
            const std::function<void(size_t, size_t)> post = [&](size_t lo, size_t hi) {
            Tally t;
            t.lo = lo;
            for (size_t i = lo; i < hi; ++i)
            {
                ReadT& read = *todo[i];
                const pg_record& r = rec[i];
                if (r.status == 3) // no enabled stage mapped it
                {
                    if (kmer_on) // KmerAlignerImpl::alignRead starts by resetting the status (KmerAligner.cpp:521)
                        read.set_graph_mapping_status(ReadT::UNMAPPED);
                    continue;
                }
                const bool by_path = r.mapped_by == PG_STAGE_PATH_ID;
                const bool by_kmer = r.mapped_by == PG_STAGE_KMER_ID || r.mapped_by == PG_STAGE_KMER_REV_ID;
                if (by_kmer && !r.unique) // KmerAligner itself: an equally good second candidate -> BAD_ALIGN (gssw is off,
                {                         // else the device had passed the read on)
                    read.set_graph_mapping_status(ReadT::BAD_ALIGN);
                    continue;
                }
                read.set_graph_mapping_status(ReadT::MAPPED);
                const bool rejected = filter && filter(read);
                if (by_path)
                {
                    ++t.mappedPath;
                    if (rejected)
                    {
                        read.set_graph_mapping_status(ReadT::BAD_ALIGN);
                        t.filtered += !kmerMatching_ && !graphMatching_;
                        if (kmerMatching_ || graphMatching_)
                            t.after_path.push_back(&read);
                    }
                }
                else if (by_kmer)
                {
                    if (rejected)
                    {
                        read.set_graph_mapping_status(ReadT::BAD_ALIGN);
                        t.filtered += !graphMatching_;
                        if (graphMatching_)
                            t.after_kmer.push_back(&read);
                    }
                    else
                        ++t.mappedKmers;
                }
                else if (rejected)
                {
                    read.set_graph_mapping_status(ReadT::BAD_ALIGN);
                    ++t.filtered;
                }
                else
                    ++t.mappedSw;
            }
            std::lock_guard<std::mutex> lock(tallies_mutex);
            tallies.push_back(std::move(t));
            };
            graphAligner_.alignBatch(todo.begin(), todo.end(), flags_, &rec, /*tolerate_unmapped=*/true, /*all_nonempty=*/true, &post);
            detail::Phases::Scope ph(detail::Phases::FILTER);
            std::sort(tallies.begin(), tallies.end(), [](Tally const& a, Tally const& b) { return a.lo < b.lo; });
            for (Tally const& t : tallies)
            {
                mappedPath_ += t.mappedPath;
                mappedKmers_ += t.mappedKmers;
                mappedSw_ += t.mappedSw;
                filtered_ += t.filtered;
                after_path.insert(after_path.end(), t.after_path.begin(), t.after_path.end());
                after_kmer.insert(after_kmer.end(), t.after_kmer.begin(), t.after_kmer.end());
            }
            ph.next(-1);
            if (path_on)
                anchoredPath_ += (unsigned)graphAligner_.lastAnchored();
            // next round: what the filter sent on.  (After the exact-match stage: k-mer stage and gssw; the reads the
            // k-mer stage of THIS round lost to the filter join the gssw-only round that follows.)
            if (path_on)
            {
                path_on = false;
                pending_gssw_.insert(pending_gssw_.end(), after_kmer.begin(), after_kmer.end());
                todo.swap(after_path);
                if (todo.empty())
                {
                    kmer_on = false;
                    todo.clear();
                    for (void* q : pending_gssw_)
                        todo.push_back(static_cast<ReadT*>(q));
                    pending_gssw_.clear();
                }
            }
            else if (kmer_on)
            {
                kmer_on = false;
                todo.clear();
                for (void* q : pending_gssw_)
                    todo.push_back(static_cast<ReadT*>(q));
                pending_gssw_.clear();
                todo.insert(todo.end(), after_kmer.begin(), after_kmer.end());
            }
            else
                todo.clear();
            if (!graphMatching_ && !kmer_on && !path_on)
                break;
        }
        pending_gssw_.clear();
    }
    template <typename ReadT, typename FilterT> void alignRead(ReadT& read, FilterT filter)
    {
        ReadT* p = &read;
        alignReads(&p, &p + 1, filter);
    }

    unsigned attempted() const { return attempted_; }
    unsigned filtered() const { return filtered_; }
    unsigned mappedKlib() const { return 0; }
    unsigned mappedPath() const { return mappedPath_; }
    unsigned anchoredPath() const { return anchoredPath_; }
    unsigned mappedKmers() const { return mappedKmers_; }
    unsigned mappedSw() const { return mappedSw_; }

private:
    const bool pathMatching_, graphMatching_, kmerMatching_;
    const unsigned flags_;
    const int pathKmerSize_, kmerSize_;
    GraphAligner graphAligner_;
    unsigned threads_ = 1;
    unsigned attempted_ = 0, filtered_ = 0, mappedSw_ = 0, mappedPath_ = 0, anchoredPath_ = 0, mappedKmers_ = 0;
    std::vector<void*> pending_gssw_; // reads the filter rejected after the k-mer stage, waiting for the gssw-only round
};

// grm::alignReads (Align.hh:49-52; Align.cpp:114-156): aligns, then keeps only MAPPED reads (input order).
// The batch is one GPU launch sequence; `threads` host threads pack it, write the results back and run the filter
// callback (detail::parallelFor: contiguous chunks of the input, so one thread means input order) -- the reference calls
// the filter from its `threads` worker threads too (Align.cpp:119-153), a filter that is safe there is safe here.
template <typename GraphT, typename PathListT, typename ReadPtrT, typename FilterT>
void alignReads(GraphT const* graph, PathListT const& paths, std::vector<ReadPtrT>& reads, FilterT const& filter,
                bool path_sequence_matching, bool graph_sequence_matching, bool klib_sequence_matching,
                bool kmer_sequence_matching, bool validate_alignments, uint32_t threads = 1, int device = 0)
{
    if (validate_alignments)
        throw std::runtime_error("paragraph_b200: ValidationAligner is diagnostics-only and not provided");
    detail::Phases::Scope ph(detail::Phases::SET_GRAPH);
    CompositeAligner aligner(path_sequence_matching, graph_sequence_matching, klib_sequence_matching,
                             kmer_sequence_matching, GraphAligner::AF_ALL, device);
    aligner.setGraph(graph, paths);
    aligner.setThreads(threads);
    typedef typename std::remove_reference<decltype(*reads.front())>::type ReadT;
    ph.next(-1);
    aligner.alignReads(reads.begin(), reads.end(), filter, /*reset_status=*/true); // (Align.cpp:72-78 inside)
    ph.next(detail::Phases::KEEP);
    // MAPPED reads only, in input order: the flags on the host threads (each touches the reads it wrote), then one pass
    // over the pointers
    std::vector<uint8_t> keep(reads.size());
    detail::parallelFor(reads.size(), threads, [&](size_t lo, size_t hi) {
        for (size_t i = lo; i < hi; ++i)
            keep[i] = !reads[i]->bases().empty() && reads[i]->graph_mapping_status() == ReadT::MAPPED;
    });
    std::vector<ReadPtrT> kept;
    kept.reserve(reads.size());
    for (size_t i = 0; i < reads.size(); ++i)
        if (keep[i])
            kept.emplace_back(std::move(reads[i]));
    reads.swap(kept); // Align.cpp:155
}

} // namespace grm
} // namespace pgb
