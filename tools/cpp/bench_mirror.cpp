// Benchmark of the drop-in surface itself: the C++ host mirror (paragraph_b200/csrc/host/pg_grm.hh) over
// std::vector<std::unique_ptr<Read>>, the way paragraph::alignAndDisambiguate calls grm::alignReads
// (src/c++/lib/paragraph/Disambiguation.cpp:207-210) and grmpy's workflow walks its sites (Workflow.cpp:108-146).
// Built by __graft_entry__.build(); driven by bench.py (legs e2e_mirror, e2e_pipeline, sweep_config4_cpp).
//
//   bench_mirror <workload file> <mode> <steps> <warmup> <host threads> [device[,device...]]
//     alignReads : per step one grm::alignReads call per site: pack + H2D + kernels + D2H + applyRecord (bases,
//                  quals, CIGAR string) + filter callback + MAPPED-only swap.  The reads are fresh copies per step
//                  (made outside the timed region: alignReads consumes its input).
//     pipeline   : SitePipeline fed by a per-site producer that builds the site's Read objects when the site is
//                  handed over (the stand-in for common::extractReads, ReadExtraction.cpp:38-122, whose htslib side
//                  is not in this image); align + filters + counts on the device, two engines double buffered.
//                  The producer's work is inside the timed region.
//     sharded    : ShardedAligner over the listed devices (one host thread + SitePipeline per device, LPT shards).
// Workload file (written by bench.py): "SITES n" then per site "SITE nodes edges reads", the node sequences, the
// edges "from to", the reads, one per line.
// Output: one JSON object on stdout.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <list>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "../../paragraph_b200/csrc/host/pg_grm.hh"

using namespace pgb;
typedef std::vector<std::unique_ptr<Read>> ReadVec;

struct SiteIn
{
    Graph graph;
    std::vector<std::string> reads;
    explicit SiteIn(size_t n) : graph(n) {}
};

static std::vector<std::unique_ptr<SiteIn>> load(const char* path)
{
    std::ifstream in(path);
    if (!in)
        throw std::runtime_error(std::string("cannot open ") + path);
    std::string tag;
    size_t n_sites = 0;
    in >> tag >> n_sites;
    std::vector<std::unique_ptr<SiteIn>> sites;
    for (size_t s = 0; s < n_sites; ++s)
    {
        size_t nn = 0, ne = 0, nr = 0;
        in >> tag >> nn >> ne >> nr;
        std::unique_ptr<SiteIn> si(new SiteIn(nn));
        std::string seq;
        for (size_t i = 0; i < nn; ++i)
        {
            in >> seq;
            si->graph.setNodeName((uint32_t)i, "n" + std::to_string(i));
            si->graph.setNodeSeq((uint32_t)i, seq);
        }
        for (size_t e = 0; e < ne; ++e)
        {
            uint32_t f = 0, t = 0;
            in >> f >> t;
            si->graph.addEdge(f, t);
        }
        si->reads.resize(nr);
        for (size_t r = 0; r < nr; ++r)
            in >> si->reads[r];
        sites.push_back(std::move(si));
    }
    if (!in)
        throw std::runtime_error("workload file is truncated");
    return sites;
}

// what the BAM side hands over per read: id, bases, quals, strand, mate flag (BamReader.cpp:50-105)
static void produce(SiteIn const& s, ReadVec& out)
{
    out.clear();
    out.reserve(s.reads.size());
    for (size_t i = 0; i < s.reads.size(); ++i)
    {
        std::unique_ptr<Read> r(new Read());
        r->setCoreInfo("frag" + std::to_string(i / 2), s.reads[i], std::string(s.reads[i].size(), '#'));
        r->set_is_reverse_strand((i & 1) != 0);
        out.push_back(std::move(r));
    }
}

static double now()
{
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

int main(int argc, char** argv)
{
    if (argc < 6)
    {
        fprintf(stderr, "usage: bench_mirror <workload> <alignReads|pipeline|sharded> <steps> <warmup> <threads> [devices]\n");
        return 2;
    }
    try
    {
        const std::string mode = argv[2];
        const int steps = atoi(argv[3]), warmup = atoi(argv[4]);
        const unsigned threads = (unsigned)atoi(argv[5]);
        std::vector<int> devices;
        {
            std::stringstream ss(argc > 6 ? argv[6] : "0");
            std::string tok;
            while (std::getline(ss, tok, ','))
                devices.push_back(atoi(tok.c_str()));
        }
        const std::vector<std::unique_ptr<SiteIn>> sites = load(argv[1]);
        size_t n_reads = 0;
        for (auto const& s : sites)
            n_reads += s->reads.size();
        // the default read filter of alignAndDisambiguate that needs no CIGAR decode on the host: NonUniq
        // (readfilters/NonUniq.hh:48-52); BadAlign runs on the device in the counting legs
        grm::ReadFilterT<Read> filter = [](Read& r) { return !r.is_graph_alignment_unique(); };
        std::list<Path> paths;
        double timed = 0, producer = 0;
        size_t kept = 0, counted_nodes = 0;
        for (int it = 0; it < warmup + steps; ++it)
        {
            const bool on = it >= warmup;
            if (it == warmup)
                grm::detail::Phases::instance().reset();
            if (mode == "alignReads")
            {
                std::vector<ReadVec> rv(sites.size());
                for (size_t s = 0; s < sites.size(); ++s)
                    produce(*sites[s], rv[s]); // outside the timed region
                const double t0 = now();
                for (size_t s = 0; s < sites.size(); ++s)
                    grm::alignReads(&sites[s]->graph, paths, rv[s], filter, false, true, false, false, false, threads, devices[0]);
                if (on)
                    timed += now() - t0;
                kept = 0;
                for (auto const& v : rv)
                    kept += v.size();
            }
            else if (mode == "pipeline" || mode == "sharded")
            {
                std::vector<ReadVec> rv(sites.size());
                std::vector<paragraph::SiteCounts> counts;
                const double t0 = now();
                double tp = 0;
                if (mode == "pipeline")
                {
                    grm::SitePipeline<std::unique_ptr<Read>> pipe(devices[0], grm::GraphAligner::AF_ALL, 1 << 16, threads);
                    for (size_t s = 0; s < sites.size(); ++s)
                    {
                        const double p0 = now();
                        produce(*sites[s], rv[s]);
                        tp += now() - p0;
                        pipe.addSite(&sites[s]->graph, &rv[s]);
                    }
                    counts = pipe.finish();
                }
                else
                {
                    grm::ShardedAligner<std::unique_ptr<Read>> sh(devices, grm::GraphAligner::AF_ALL, 1 << 16, threads);
                    for (size_t s = 0; s < sites.size(); ++s)
                    {
                        const double p0 = now();
                        produce(*sites[s], rv[s]);
                        tp += now() - p0;
                        sh.addSite(&sites[s]->graph, &rv[s]);
                    }
                    counts = sh.run();
                }
                if (on)
                {
                    timed += now() - t0;
                    producer += tp;
                }
                kept = counted_nodes = 0;
                for (auto const& v : rv)
                    kept += v.size();
                for (auto const& c : counts)
                    counted_nodes += c.read_counts_by_node.size();
            }
            else
                throw std::runtime_error("unknown mode " + mode);
        }
        {
            // host phases of the mirror, mean per timed step
            auto& ph = pgb::grm::detail::Phases::instance();
            fprintf(stderr, "phases (ms per step, %d timed steps):", steps);
            for (int i = 0; i < pgb::grm::detail::Phases::N; ++i)
                fprintf(stderr, " %s %.3f", pgb::grm::detail::Phases::name(i), ph.ns[i] * 1e-6 / (steps > 0 ? steps : 1));
            fprintf(stderr, "\n");
        }
        printf("{\"mode\": \"%s\", \"sites\": %zu, \"reads\": %zu, \"steps\": %d, \"threads\": %u, \"devices\": %zu, "
               "\"seconds\": %.6f, \"reads_per_s\": %.1f, \"producer_seconds\": %.6f, \"kept\": %zu, \"node_rows\": %zu}\n",
               mode.c_str(), sites.size(), n_reads, steps, threads, devices.size(), timed,
               timed > 0 ? (double)n_reads * steps / timed : 0.0, producer, kept, counted_nodes);
    }
    catch (std::exception const& e)
    {
        fprintf(stderr, "bench_mirror: %s\n", e.what());
        return 1;
    }
    return 0;
}
