// TEST INFRASTRUCTURE ONLY: Boost is not in this image.  ReadCounting.cpp::alignmentStats (fragment-length
// mean/median/variance) needs boost::accumulators; this stand-in lets the file compile and yields 0 for those
// statistics.  "fragment_statistics" is therefore NOT part of any parity claim; the node/edge/sequence count
// code in the same file (countNodes/countEdges/countPathFamilies/countReads) is the unmodified reference.
#pragma once
#include <utility>
#include <vector>
namespace boost
{
template <typename It> struct iterator_range
{
    It b, e;
    It begin() const { return b; }
    It end() const { return e; }
};
namespace accumulators
{
    namespace tag
    {
        struct named_arg
        {
            int v;
            named_arg operator=(int x) const { return named_arg{ x }; }
        };
        struct mean {};
        struct median {};
        struct variance {};
        struct density
        {
            static constexpr named_arg num_bins{ 0 };
            static constexpr named_arg cache_size{ 0 };
        };
    }
    template <typename... T> struct features {};
    template <typename V, typename F> struct accumulator_set
    {
        template <typename... A> accumulator_set(A...) {}
        void operator()(V) {}
    };
    template <typename A> double mean(A const&) { return 0; }
    template <typename A> double median(A const&) { return 0; }
    template <typename A> double variance(A const&) { return 0; }
    template <typename A> iterator_range<std::vector<std::pair<double, double>>::iterator> density(A const&) { return {}; }
}
}
