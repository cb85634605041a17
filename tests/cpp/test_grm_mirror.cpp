// Drives the host-side mirror (paragraph_b200/csrc/host/pg_grm.hh) the way the reference's own unit test drives
// grm::alignReads (src/c++/test/test_paragraph_parts.cpp:52-104): same graph, same reads, gssw stage only.
// Prints one line per surviving read: id pos cigar score mapq reverse bases status
#include <cstdio>
#include <cstring>
#include <string>
#include <list>
#include <memory>
#include <vector>

#include "../../paragraph_b200/csrc/host/pg_grm.hh"

int main()
{
    using namespace pgb;
    std::vector<Read> reads(7);
    reads[0].setCoreInfo("f1", "AAAAAAAATTTTCTTTAAAAAAAA", "########################");
    reads[1].setCoreInfo("f2", "TTTTTTAAAGAAAATTTTTTT", "#####################");
    reads[2].setCoreInfo("f3", "AAAAAGCGGGGGGAAAAAA", "###################");
    reads[3].setCoreInfo("f4", "AAAAGCGGGGGGAAAAAA", "##################");
    reads[4].setCoreInfo("f5", "TTTTTTCCCCCCGCTTTTT", "###################");
    reads[5].setCoreInfo("f6", "AAAAAAAAAAAAAAAAAAA", "###################");
    reads[6].setCoreInfo("f7", "ATATATAT", "########"); // best score ends in several nodes on both strands -> not unique -> filtered
    Graph graph(4);
    graph.setNodeSeq(0, "AAAAAAAAAAA");
    graph.setNodeSeq(1, "TTTTTTTT");
    graph.setNodeSeq(2, "GGGGGGGG");
    graph.setNodeSeq(3, "AAAAAAAAAAA");
    graph.addEdge(0, 1);
    graph.addEdge(0, 2);
    graph.addEdge(0, 3);
    graph.addEdge(1, 3);
    graph.addEdge(2, 3);
    std::vector<std::unique_ptr<Read>> rb;
    for (auto const& r : reads)
        rb.emplace_back(new Read(r));
    std::list<Path> paths;
    // NonUniq read filter (src/c++/lib/paragraph/readfilters/NonUniq.hh:48-52)
    grm::ReadFilterT<Read> filter = [](Read& r) { return !r.is_graph_alignment_unique(); };
    try
    {
        grm::alignReads(&graph, paths, rb, filter, false, true, false, false, false, 4);
        for (auto const& r : rb)
            printf("%s %d %s %d %d %d %s %d\n", r->fragment_id().c_str(), r->graph_pos(), r->graph_cigar().c_str(),
                   r->graph_alignment_score(), r->graph_mapq(), (int)r->is_graph_reverse_strand(), r->bases().c_str(),
                   (int)r->graph_mapping_status());
        bool threw = false;
        try
        {
            grm::CompositeAligner bad(true, true, true, false); // klib stage: not on the GPU
        }
        catch (std::runtime_error const&)
        {
            threw = true;
        }
        printf("klib-stage-throws %d\n", (int)threw);

        // the k-mer stage (grm::KmerAligner): the reference's own unit test (src/c++/test/test_kmeraligner.cpp:58-191:
        // K = 10, paths P / Q / D), first the stage alone, then with gssw behind it and the NonUniq filter
        {
            const char* kb[6] = { "AAAAAAAATTTTTTTTAAAAAAAA", "TTTTTTAAAAAAAATTTTTTT", "AAAAAGGGGGGGGAAAAAA", "AAAAGGGGGGGGAAAAAA",
                                  "TTTTTTCCCCCCCCTTTTT", "AAAAAAAAAAAAAAAAAAA" };
            std::list<Path> kpaths;
            kpaths.emplace_back(std::vector<uint32_t>{ 0, 1, 3 });
            kpaths.emplace_back(std::vector<uint32_t>{ 0, 2, 3 });
            kpaths.emplace_back(std::vector<uint32_t>{ 0, 3 });
            for (int with_gssw = 0; with_gssw < 2; ++with_gssw)
            {
                std::vector<std::unique_ptr<Read>> kr;
                for (int i = 0; i < 6; ++i)
                {
                    kr.emplace_back(new Read());
                    kr.back()->setCoreInfo("q" + std::to_string(i + 1), kb[i], std::string(strlen(kb[i]), '#'));
                }
                grm::CompositeAligner ka(false, with_gssw != 0, false, true, grm::GraphAligner::AF_ALL, 0, 32, 10);
                ka.setGraph(&graph, kpaths);
                ka.alignReads(kr.begin(), kr.end(), filter);
                for (auto const& r : kr)
                    printf("%s %d %s %d %d %d %s %d\n", r->fragment_id().c_str(), r->graph_pos(), r->graph_cigar().c_str(),
                           r->graph_alignment_score(), r->graph_mapq(), (int)r->is_graph_reverse_strand(), r->bases().c_str(),
                           (int)r->graph_mapping_status());
                printf("kmer-stage %u %u %u %u\n", ka.attempted(), ka.mappedKmers(), ka.mappedSw(), ka.filtered());
            }
        }

        // alignAndDisambiguate's core on the device: ParagraphTest.Aligns expects these supports
        // (test_paragraph_parts.cpp:113-144; the test calls disambiguateReads with null filters)
        {
            Graph lg = graph;
            lg.addLabelToEdge(0, 1, "P");
            lg.addLabelToEdge(1, 3, "P");
            lg.addLabelToEdge(0, 2, "Q");
            lg.addLabelToEdge(2, 3, "Q");
            lg.addLabelToEdge(0, 3, "D");
            std::vector<std::unique_ptr<Read>> rd;
            for (auto const& r : reads)
                rd.emplace_back(new Read(r));
            const char* names[] = { "LF", "P1", "Q1", "RF" };
            for (uint32_t v = 0; v < 4; ++v)
                lg.setNodeName(v, names[v]);
            grm::MultiSiteAligner<std::unique_ptr<Read>> counter;
            counter.addSite(&lg, &rd);
            pgb::paragraph::CountOptions opt;
            opt.use_support_filters = false;
            auto counts = counter.alignAndCount(opt);
            printf("count-sites %zu reads %zu\n", counts.size(), rd.size());
            for (auto const& r : rd)
            {
                printf("s %s n", r->fragment_id().c_str());
                for (auto const& x : r->graph_nodes_supported())
                    printf(" %s", x.c_str());
                printf(" e");
                for (auto const& x : r->graph_edges_supported())
                    printf(" %s", x.c_str());
                printf(" q");
                for (auto const& x : r->graph_sequences_supported())
                    printf(" %s", x.c_str());
                printf("\n");
            }
            for (auto const& kv : counts[0].read_counts_by_node)
                printf("cn %s %llu %llu %llu %llu\n", kv.first.c_str(), (unsigned long long)kv.second.fragments,
                       (unsigned long long)kv.second.reads, (unsigned long long)kv.second.fwd, (unsigned long long)kv.second.rev);
            for (auto const& kv : counts[0].read_counts_by_edge)
                printf("ce %s %llu\n", kv.first.c_str(), (unsigned long long)kv.second.fragments);
            for (auto const& kv : counts[0].read_counts_by_sequence)
                printf("cs %s total %llu keys %zu\n", kv.first.c_str(), (unsigned long long)kv.second.at("total").fragments,
                       kv.second.size());
        }

        // the same site twice plus a second graph, all in ONE launch through MultiSiteAligner
        Graph g2(2);
        g2.setNodeSeq(0, "ACGTACGTAC");
        g2.setNodeSeq(1, "GGGTTTGGGA");
        g2.addEdge(0, 1);
        std::vector<std::unique_ptr<Read>> a, b, c;
        for (auto const& r : reads)
        {
            a.emplace_back(new Read(r));
            b.emplace_back(new Read(r));
        }
        c.emplace_back(new Read("g1", "GTACGGGTTT", "##########"));
        c.emplace_back(new Read("g2", "AAACCCGTAC", "##########")); // reverse complement of GTACGGGTTT
        grm::MultiSiteAligner<std::unique_ptr<Read>> multi;
        multi.addSite(&graph, &a);
        multi.addSite(&g2, &c);
        multi.addSite(&graph, &b);
        multi.run(filter);
        printf("multi %zu %zu %zu\n", a.size(), c.size(), b.size());
        for (auto const& r : b)
            printf("m %s %d %s %d\n", r->fragment_id().c_str(), r->graph_pos(), r->graph_cigar().c_str(),
                   r->graph_alignment_score());
        for (auto const& r : c)
            printf("m %s %d %s %d %d\n", r->fragment_id().c_str(), r->graph_pos(), r->graph_cigar().c_str(),
                   r->graph_alignment_score(), (int)r->is_graph_reverse_strand());
        // the cascade: exact-match stage (PathAligner, k = 8 so that it bites on this tiny graph) in front of gssw, with the
        // NonUniq filter -- c10 matches exactly on both strands (non-unique for PathAligner), is rejected by the filter
        // and gets its second chance in the gssw stage (CompositeAligner.cpp:97-103)
        {
            const char* cr[] = { "AAAAAAAATTTTTTTTAAAAAAAA", "TTTTTTTTAAAAAAAATTTTTTTT", "AAAAAAAATTTTCTTTAAAAAAAA", "AAAAAAAAAAAAAAAAAAA",
                                 "GGGGGGGG", "ATATATAT", "AAAATTTTTTTTAAAA", "AAAGGGGGGGGAAA", "TTTCCCCCCCCTTT", "AAAAAATTTTTT" };
            std::vector<std::unique_ptr<Read>> cs;
            for (int i = 0; i < 10; ++i)
                cs.emplace_back(new Read("c" + std::to_string(i + 1), cr[i], std::string(strlen(cr[i]), '#')));
            grm::CompositeAligner cascade(true, true, false, false, grm::GraphAligner::AF_ALL, 0, 8);
            cascade.setGraph(&graph, paths);
            for (auto& r : cs)
                r->set_graph_mapping_status(Read::UNMAPPED);
            cascade.alignReads(cs.begin(), cs.end(), filter);
            for (auto const& r : cs)
                printf("%s %d %s %d %d %d %s %d\n", r->fragment_id().c_str(), r->graph_pos(), r->graph_cigar().c_str(),
                       r->graph_alignment_score(), r->graph_mapq(), (int)r->is_graph_reverse_strand(), r->bases().c_str(),
                       (int)r->graph_mapping_status());
            printf("cascade %u %u %u %u %u\n", cascade.attempted(), cascade.mappedPath(), cascade.anchoredPath(),
                   cascade.mappedSw(), cascade.filtered());

            // ... and the same cascade under alignAndCount: exact-match stage + second chance on the device, supports
            // and counts from the device's counting stage (path-mapped reads keep PathAligner's strand convention)
            Graph lg = graph;
            lg.addLabelToEdge(0, 1, "P");
            lg.addLabelToEdge(1, 3, "P");
            lg.addLabelToEdge(0, 2, "Q");
            lg.addLabelToEdge(2, 3, "Q");
            lg.addLabelToEdge(0, 3, "D");
            const char* names[] = { "LF", "P1", "Q1", "RF" };
            for (uint32_t v = 0; v < 4; ++v)
                lg.setNodeName(v, names[v]);
            std::vector<std::unique_ptr<Read>> cd;
            for (int i = 0; i < 10; ++i)
                cd.emplace_back(new Read("c" + std::to_string(i + 1), cr[i], std::string(strlen(cr[i]), '#')));
            grm::MultiSiteAligner<std::unique_ptr<Read>> counter;
            counter.setPathMatching(8);
            counter.addSite(&lg, &cd);
            pgb::paragraph::CountOptions opt;
            opt.use_support_filters = false;
            auto counts = counter.alignAndCount(opt);
            for (auto const& r : cd)
            {
                printf("k %s %s n", r->fragment_id().c_str(), r->graph_cigar().c_str());
                for (auto const& x : r->graph_nodes_supported())
                    printf(" %s", x.c_str());
                printf(" q");
                for (auto const& x : r->graph_sequences_supported())
                    printf(" %s", x.c_str());
                printf("\n");
            }
            for (auto const& kv : counts[0].read_counts_by_node)
                printf("kn %s %llu %llu %llu %llu\n", kv.first.c_str(), (unsigned long long)kv.second.fragments,
                       (unsigned long long)kv.second.reads, (unsigned long long)kv.second.fwd, (unsigned long long)kv.second.rev);
            for (auto const& kv : counts[0].read_counts_by_edge)
                printf("ke %s %llu\n", kv.first.c_str(), (unsigned long long)kv.second.fragments);
        }
        // edges of the contract: a read without bases is skipped and dropped (Align.cpp:74-77), no reads at all is
        // fine, a read beyond PG_MAX_READ_LEN is an error (thrown like the reference's error()), not a silent skip
        {
            std::vector<std::unique_ptr<Read>> e1;
            e1.emplace_back(new Read(reads[0]));
            e1.emplace_back(new Read("empty", "", ""));
            e1.emplace_back(new Read(reads[5]));
            grm::alignReads(&graph, paths, e1, filter, false, true, false, false, false, 2);
            const bool dropped = e1.size() == 2 && e1[0]->fragment_id() == "f1" && e1[1]->fragment_id() == "f6";
            std::vector<std::unique_ptr<Read>> e2;
            grm::alignReads(&graph, paths, e2, filter, true, true, false, false, false, 2);
            std::vector<std::unique_ptr<Read>> e3;
            e3.emplace_back(new Read("long", std::string(1100, 'A'), std::string(1100, '#')));
            bool long_throws = false;
            try
            {
                grm::alignReads(&graph, paths, e3, filter, false, true, false, false, false, 1);
            }
            catch (std::runtime_error const&)
            {
                long_throws = true;
            }
            printf("edge empty-dropped %d none-ok %d long-throws %d\n", (int)dropped, (int)e2.empty(), (int)long_throws);
        }
        // `threads`: packing and writing back on several host threads gives the same reads in the same order
        {
            std::vector<std::unique_ptr<Read>> t1, t5;
            for (int i = 0; i < 4200; ++i)
            {
                Read r(reads[(size_t)(i % 7)]);
                r.set_is_reverse_strand((i / 7) % 2 == 1);
                t1.emplace_back(new Read(r));
                t5.emplace_back(new Read(r));
            }
            grm::alignReads(&graph, paths, t1, filter, true, true, false, false, false, 1);
            grm::alignReads(&graph, paths, t5, filter, true, true, false, false, false, 5);
            bool equal = t1.size() == t5.size();
            for (size_t i = 0; equal && i < t1.size(); ++i)
                equal = t1[i]->fragment_id() == t5[i]->fragment_id() && t1[i]->bases() == t5[i]->bases()
                    && t1[i]->quals() == t5[i]->quals() && t1[i]->graph_cigar() == t5[i]->graph_cigar()
                    && t1[i]->graph_pos() == t5[i]->graph_pos() && t1[i]->graph_mapq() == t5[i]->graph_mapq()
                    && t1[i]->graph_alignment_score() == t5[i]->graph_alignment_score()
                    && t1[i]->is_graph_reverse_strand() == t5[i]->is_graph_reverse_strand()
                    && t1[i]->graph_mapping_status() == t5[i]->graph_mapping_status();
            // GraphAligner::alignBatch as one submit + collect and as a software pipeline of five parts over two engines
            {
                std::vector<std::unique_ptr<Read>> whole, parts;
                for (int i = 0; i < 4200; ++i)
                {
                    Read r(reads[(size_t)(i % 7)]);
                    r.set_is_reverse_strand((i / 7) % 2 == 1);
                    whole.emplace_back(new Read(r));
                    parts.emplace_back(new Read(r));
                }
                grm::GraphAligner a, b;
                a.setGraph(&graph);
                a.setPipelineMinReads(0);
                a.alignBatch(whole.begin(), whole.end());
                b.setGraph(&graph);
                b.setThreads(3);
                b.setPipelineMinReads(512);
                b.setPipelineMaxParts(5);
                b.alignBatch(parts.begin(), parts.end());
                for (size_t i = 0; equal && i < whole.size(); ++i)
                    equal = whole[i]->bases() == parts[i]->bases() && whole[i]->quals() == parts[i]->quals()
                        && whole[i]->graph_cigar() == parts[i]->graph_cigar() && whole[i]->graph_pos() == parts[i]->graph_pos()
                        && whole[i]->graph_mapq() == parts[i]->graph_mapq()
                        && whole[i]->is_graph_reverse_strand() == parts[i]->is_graph_reverse_strand();
            }
            // the same for MultiSiteAligner::alignAndCount (two sites, exact-match stage in front)
            Graph lg = graph;
            lg.addLabelToEdge(0, 1, "P");
            lg.addLabelToEdge(1, 3, "P");
            lg.addLabelToEdge(0, 2, "Q");
            lg.addLabelToEdge(2, 3, "Q");
            lg.addLabelToEdge(0, 3, "D");
            const char* names[] = { "LF", "P1", "Q1", "RF" };
            for (uint32_t v = 0; v < 4; ++v)
                lg.setNodeName(v, names[v]);
            std::vector<std::unique_ptr<Read>> m[2][2];
            std::vector<paragraph::SiteCounts> counts[2];
            for (int variant = 0; variant < 2; ++variant)
            {
                for (int i = 0; i < 4200; ++i)
                {
                    Read r(reads[(size_t)(i % 7)]);
                    r.setCoreInfo("frag" + std::to_string(i / 2), r.bases(), r.quals());
                    r.set_is_reverse_strand(i % 2 == 1);
                    m[variant][i % 2].emplace_back(new Read(r));
                }
                grm::MultiSiteAligner<std::unique_ptr<Read>> ms;
                ms.setPathMatching(8);
                ms.setThreads(variant ? 5 : 1);
                ms.addSite(&lg, &m[variant][0]);
                ms.addSite(&lg, &m[variant][1]);
                counts[variant] = ms.alignAndCount();
            }
            for (int k = 0; k < 2; ++k)
            {
                equal = equal && m[0][k].size() == m[1][k].size()
                    && counts[0][(size_t)k].read_counts_by_node.size() == counts[1][(size_t)k].read_counts_by_node.size();
                for (size_t i = 0; equal && i < m[0][k].size(); ++i)
                    equal = m[0][k][i]->bases() == m[1][k][i]->bases() && m[0][k][i]->graph_cigar() == m[1][k][i]->graph_cigar()
                        && m[0][k][i]->graph_nodes_supported() == m[1][k][i]->graph_nodes_supported()
                        && m[0][k][i]->graph_sequences_supported() == m[1][k][i]->graph_sequences_supported();
                for (auto const& kv : counts[0][(size_t)k].read_counts_by_node)
                    equal = equal && counts[1][(size_t)k].read_counts_by_node.count(kv.first)
                        && counts[1][(size_t)k].read_counts_by_node.at(kv.first).reads == kv.second.reads;
            }
            // SitePipeline: six sites in batches of ~1 000 reads on two engines = one MultiSiteAligner over all six
            std::vector<std::unique_ptr<Read>> p[2][6];
            std::vector<paragraph::SiteCounts> pc[2];
            for (int variant = 0; variant < 2; ++variant)
                for (int k = 0; k < 6; ++k)
                    for (int i = 0; i < 700; ++i)
                    {
                        Read r(reads[(size_t)((i + k) % 7)]);
                        r.setCoreInfo("frag" + std::to_string(i / 2), r.bases(), r.quals());
                        r.set_is_reverse_strand((i + k) % 3 == 1);
                        p[variant][k].emplace_back(new Read(r));
                    }
            {
                grm::MultiSiteAligner<std::unique_ptr<Read>> ms;
                ms.setPathMatching(8);
                for (int k = 0; k < 6; ++k)
                    ms.addSite(&lg, &p[0][k]);
                pc[0] = ms.alignAndCount();
                grm::SitePipeline<std::unique_ptr<Read>> pipe(0, grm::GraphAligner::AF_ALL, 1000, 3);
                pipe.setPathMatching(8);
                for (int k = 0; k < 6; ++k)
                    pipe.addSite(&lg, &p[1][k]);
                pc[1] = pipe.finish();
            }
            bool pipe_equal = pc[0].size() == 6 && pc[1].size() == 6;
            size_t pipe_kept = 0;
            for (int k = 0; pipe_equal && k < 6; ++k)
            {
                pipe_equal = p[0][k].size() == p[1][k].size()
                    && pc[0][(size_t)k].read_counts_by_node.size() == pc[1][(size_t)k].read_counts_by_node.size()
                    && pc[0][(size_t)k].read_counts_by_sequence.size() == pc[1][(size_t)k].read_counts_by_sequence.size();
                for (size_t i = 0; pipe_equal && i < p[0][k].size(); ++i)
                    pipe_equal = p[0][k][i]->bases() == p[1][k][i]->bases() && p[0][k][i]->graph_cigar() == p[1][k][i]->graph_cigar()
                        && p[0][k][i]->graph_nodes_supported() == p[1][k][i]->graph_nodes_supported();
                for (auto const& kv : pc[0][(size_t)k].read_counts_by_node)
                    pipe_equal = pipe_equal && pc[1][(size_t)k].read_counts_by_node.count(kv.first)
                        && pc[1][(size_t)k].read_counts_by_node.at(kv.first).fragments == kv.second.fragments
                        && pc[1][(size_t)k].read_counts_by_node.at(kv.first).reads == kv.second.reads;
                for (auto const& kv : pc[0][(size_t)k].read_counts_by_edge)
                    pipe_equal = pipe_equal && pc[1][(size_t)k].read_counts_by_edge.count(kv.first)
                        && pc[1][(size_t)k].read_counts_by_edge.at(kv.first).fragments == kv.second.fragments;
                pipe_kept += p[1][k].size();
            }
            // ShardedAligner: the same six sites (different sizes) over three shards = one MultiSiteAligner over all six
            {
                std::vector<std::unique_ptr<Read>> q[6];
                for (int k = 0; k < 6; ++k)
                    for (int i = 0; i < 700 - 100 * k; ++i)
                    {
                        Read r(reads[(size_t)((i + k) % 7)]);
                        r.setCoreInfo("frag" + std::to_string(i / 2), r.bases(), r.quals());
                        r.set_is_reverse_strand((i + k) % 3 == 1);
                        q[k].emplace_back(new Read(r));
                    }
                grm::ShardedAligner<std::unique_ptr<Read>> sh(std::vector<int>{ 0, 0, 0 }, grm::GraphAligner::AF_ALL, 500, 2);
                sh.setPathMatching(8);
                for (int k = 0; k < 6; ++k)
                    sh.addSite(&lg, &q[k]);
                const std::vector<int> part = sh.partition();
                const std::vector<paragraph::SiteCounts> sc = sh.run();
                bool sh_equal = sc.size() == 6 && part.size() == 6;
                // LPT over costs 700, 600, ... 200 (x the same graph): shards {0, 5}, {1, 4}, {2, 3}
                sh_equal = sh_equal && part[0] == 0 && part[1] == 1 && part[2] == 2 && part[3] == 2 && part[4] == 1 && part[5] == 0;
                for (int k = 0; sh_equal && k < 6; ++k)
                {
                    // site k holds the first 700 - 100 k reads of pipeline site k: same reads survive, same fields
                    size_t kept_ref = 0;
                    std::vector<std::unique_ptr<Read>> again;
                    for (int i = 0; i < 700 - 100 * k; ++i)
                    {
                        Read r(reads[(size_t)((i + k) % 7)]);
                        r.setCoreInfo("frag" + std::to_string(i / 2), r.bases(), r.quals());
                        r.set_is_reverse_strand((i + k) % 3 == 1);
                        again.emplace_back(new Read(r));
                    }
                    grm::MultiSiteAligner<std::unique_ptr<Read>> one;
                    one.setPathMatching(8);
                    one.addSite(&lg, &again);
                    const std::vector<paragraph::SiteCounts> oc = one.alignAndCount();
                    kept_ref = again.size();
                    sh_equal = q[k].size() == kept_ref && oc[0].read_counts_by_node.size() == sc[(size_t)k].read_counts_by_node.size()
                        && oc[0].read_counts_by_edge.size() == sc[(size_t)k].read_counts_by_edge.size();
                    for (size_t i = 0; sh_equal && i < kept_ref; ++i)
                        sh_equal = q[k][i]->graph_cigar() == again[i]->graph_cigar() && q[k][i]->bases() == again[i]->bases()
                            && q[k][i]->graph_nodes_supported() == again[i]->graph_nodes_supported();
                    for (auto const& kv : oc[0].read_counts_by_node)
                        sh_equal = sh_equal && sc[(size_t)k].read_counts_by_node.count(kv.first)
                            && sc[(size_t)k].read_counts_by_node.at(kv.first).fragments == kv.second.fragments;
                }
                printf("sharded-equal %d\n", (int)sh_equal);
            }
            printf("threads-equal %d kept %zu of 4200, multi-site kept %zu + %zu, pipeline-equal %d kept %zu of 4200\n", (int)equal,
                   t1.size(), m[1][0].size(), m[1][1].size(), (int)pipe_equal, pipe_kept);
        }
    }
    catch (std::exception const& e)
    {
        fprintf(stderr, "ERROR %s\n", e.what());
        return 2;
    }
    return 0;
}
